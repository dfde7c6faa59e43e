#!/usr/bin/env python3
"""snapb200_filter_paired_batch on the GPU against the reference's AlignmentFilter (run as a script, in its own process, by
tests/test_filter_oracle.py::test_cuda_filter_first_run): 1500 simulated spliced / chimeric pairs, every input produced by the CUDA
library itself (transcriptome multi-hits, genome pairs, CharacterizeSeeds tuples), records compared field by field and the event
records replayed into the reference's GTFReader to compare the statistics files."""
import ctypes as C
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import filter_cases as F  # noqa: E402
import snap_rnaseq_b200 as S  # noqa: E402
from oracle import oracle as O  # noqa: E402
from snap_rnaseq_b200 import _abi as A  # noqa: E402


def main():
    ref = O.ref()
    L = S.lib(0)
    with tempfile.TemporaryDirectory() as d:
        contigs = F.build_workspace(d, O.REF_BIN)
        (b0, b1), sam_reads = F.reads(contigs, d)
        gdir, tdir, gtf = os.path.join(d, "gidx"), os.path.join(d, "tidx"), os.path.join(d, "a.gtf")
        hg, ht = L.load_index(gdir), L.load_index(tdir)
        ann = L.annotation_open(hg, ht, gtf)
        hits, genome_res, pp = F.alignments(L, hg, ht, b0, b1)
        cp = A.single_defaults(max_hits=300, num_seeds=12)
        ch = [L.characterize(hg, cp, b) for b in (b0, b1)]
        prm = A.FilterParams(pp.max_spacing, pp.force_spacing, 2, 15, F.MAX_HITS_TO_GET)
        res, ev, needs_host = L.filter_paired(ann, prm, np.diff(b0.offsets), np.diff(b1.offsets), hits[0], hits[1], genome_res, ch[0], ch[1])
        assert not needs_host.any(), f"{int(needs_host.sum())} pairs overflowed the device scratch"
        rg, rt = ref.load_index(gdir), ref.load_index(tdir)
        want = F.run_reference_filter(ref, rg, rt, gtf, os.path.join(d, "want"), sam_reads, hits, genome_res, pp)
        bad = [i for i in range(b0.n) if any(not np.array_equal(want[f][i], res[f][i]) for f in want.dtype.names if f != "pad")]
        assert not bad, (len(bad), bad[:10], [(want[i], res[i]) for i in bad[:3]])
        lib = ref.lib
        lib.ref_gtf_load.restype = C.c_void_p
        g2 = C.c_void_p(lib.ref_gtf_load(gtf.encode(), os.path.join(d, "replay").encode()))
        lib.ref_gtf_export(g2, os.path.join(d, "gtf.tsv").encode())
        t_ids = [ln.split("\t")[1] for ln in open(os.path.join(d, "gtf.tsv")) if ln.startswith("T")]
        with open(os.path.join(gdir, "Genome"), "rb") as f:
            k = int(f.readline().split()[1])
            chr_names = [f.readline().decode().rstrip("\n").split(" ", 1)[1] for _ in range(k)]
        arr = lambda names: (C.c_char_p * len(names))(*[x.encode() for x in names])
        evc = np.ascontiguousarray(ev)
        assert lib.ref_filter_replay_events(rg, rt, g2, sam_reads[0].byref(), sam_reads[1].byref(), C.c_uint(15), C.c_void_p(evc.ctypes.data), arr(t_ids),
                                            arr(chr_names)) == 0
        lib.ref_gtf_finish(g2)
        for f in sorted(x for x in os.listdir(d) if x.startswith("replay")):
            assert open(os.path.join(d, f), "rb").read() == open(os.path.join(d, "want" + f[len("replay"):]), "rb").read(), f
        L.annotation_close(ann)
        print("FILTER_GPU_OK", b0.n, "pairs;", int(res["is_transcriptome"].sum()), "ends placed by a transcriptome alignment;", int((ev["kind"] > 0).sum()), "events")


if __name__ == "__main__":
    main()
