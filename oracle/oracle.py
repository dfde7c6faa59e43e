"""ctypes loaders for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

  port()  -> oracle/liboracle.so        the C restatement (oracle/snap_oracle.c), built by `make -C oracle`
  ref()   -> oracle/_ref/libsnapref.so  the reference itself (built from /root/reference by build_ref.py)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from snap_rnaseq_b200 import _abi as A  # noqa: E402
from snap_rnaseq_b200._binding import BatchLib  # noqa: E402

PORT_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libsnapref.so")
REF_BIN = os.path.join(HERE, "_ref", "snap-rna")


def build_port():
    subprocess.run(["make", "-C", HERE, "-s"], check=True)


def have_ref():
    return os.path.exists(REF_SO) and os.path.exists(REF_BIN)


class _Cpu(BatchLib):
    def load_index(self, d):
        f = getattr(self.lib, self.prefix + "index_load")
        f.restype = C.c_void_p
        h = f(str(d).encode())
        if not h:
            raise RuntimeError(f"{self.prefix}index_load({d}) failed")
        return C.c_void_p(h)

    def index_info(self, h):
        info = A.IndexInfo()
        self.fn("index_info")(h, C.byref(info))
        return info


_port = None
_ref = {}


def port():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(HERE, "snap_oracle.c")):
            build_port()
        _port = _Cpu(C.CDLL(PORT_SO), "oracle_")
    return _port


def ref(threads=1):
    if threads not in _ref:
        if not have_ref():
            raise RuntimeError("oracle/_ref not built (needs /root/reference; run python oracle/build_ref.py)")
        lib = C.CDLL(REF_SO)
        lib.ref_init()
        _ref[threads] = _Cpu(lib, "ref_", threads=threads)
    return _ref[threads]


def ref_genome_bytes(h, start, n):
    import numpy as np
    out = np.zeros(n, np.uint8)
    rc = ref().lib.ref_genome_bytes(h, C.c_uint(start), C.c_uint(n), A.p8(out))
    assert rc == 0
    return out


def ref_build_index(fasta, outdir, seed_len=20, threads=1, extra=()):
    """`snap-rna index <fa> <dir> -s N -tN` with the compiled reference (SNAPLib/GenomeIndex.cpp:197-292)."""
    if not have_ref():
        raise RuntimeError("oracle/_ref/snap-rna not built")
    cmd = [REF_BIN, "index", str(fasta), str(outdir), "-s", str(seed_len), f"-t{threads}", *extra]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 or not os.path.exists(os.path.join(outdir, "GenomeIndexHash")):
        raise RuntimeError("reference indexer failed:\n" + r.stdout[-2000:])
    return outdir
