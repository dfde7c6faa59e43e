#!/usr/bin/env python3
"""A/B harness for kernel experiments: several builds of the library, one workload, one gpurun call.

  python scripts/ab_variants.py build name1=-DFLAG_A name2=-DFLAG_B,-DX=3 ...   (here, no GPU: nvcc cross-compiles)
  python scripts/ab_variants.py run [pairs] [c3|c2]                             (on the GPU box)

`build` writes snap_rnaseq_b200/variants/libsnapb200_<name>.so (git-ignored, travels with gpurun).  `run` generates the genome and
the pairs once, then for the production library and every variant: builds the index, runs the paired path four times on a resident
batch, and prints the best step time, the time of the dominant kernel and whether every result record is bit-identical to the
production library's.  One JSON line at the end.  GPU-minutes are the scarce resource: this turns N experiments into one call."""
import ctypes as C
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from snap_rnaseq_b200 import build as B
    if len(sys.argv) >= 2 and sys.argv[1] == "build":
        for spec in sys.argv[2:]:
            name, _, flags = spec.partition("=")
            print(B.build_variant(name, [f for f in flags.split(",") if f]))
        return
    import numpy as np
    import bench
    import snap_rnaseq_b200 as S
    from snap_rnaseq_b200 import _abi as A, synth
    args = sys.argv[2:] if len(sys.argv) >= 2 and sys.argv[1] == "run" else sys.argv[1:]
    pairs = int(args[0]) if args else 1_000_000
    if len(args) > 1 and args[1] == "c3":
        bench.GENOME_CONTIGS, bench.READ_LEN, bench.ERR_RATE = [25_000_000] * 124, 150, 0.01
    contigs = bench.make_genome()
    bases, offs = synth.snap_layout(contigs, 500)
    b0, b1 = bench.make_pairs(contigs, pairs, 1000)
    p = A.paired_defaults()
    libs = [("production", S.SO_PATH)] + [(os.path.basename(f)[len("libsnapb200_"):-3], f) for f in sorted(glob.glob(os.path.join(B.VARIANT_DIR, "libsnapb200_*.so")))]
    base, rows = None, []
    for name, path in libs:
        L = S.Snapb200(C.CDLL(path), device=0)
        h = L.build_index(bases, offs, list(contigs), seed_len=20)
        sess = S.Session(L, h, pairs, 256)
        sess.upload(0, b0)
        sess.upload(1, b1)
        times = []
        for _ in range(4):
            sess.run_paired(p)
            times.append((sess.last_run()[0], sess.main_kernel_ms()))
        out = np.zeros(pairs, A.PAIRED_RESULT)
        sess.download_paired(out)
        sess.close()
        L.close_index(h)
        if base is None:
            base = out
        same = all(np.array_equal(out[f], base[f]) for f in out.dtype.names if f != "pad")
        best = min(times)
        rows.append({"variant": name, "step_ms": best[0], "main_kernel_ms": best[1], "bit_identical_to_production": bool(same)})
        print(f"{name:24s} step {best[0]:8.2f} ms   main kernel {best[1]:8.2f} ms   identical {same}", flush=True)
    print(json.dumps({"pairs": pairs, "genome_mbp": sum(bench.GENOME_CONTIGS) // 1_000_000, "read_len": bench.READ_LEN, "variants": rows}))


if __name__ == "__main__":
    main()
