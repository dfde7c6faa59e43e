#!/usr/bin/env python3
"""Golden vectors for AlignmentFilter (row f3, next): the COMPILED REFERENCE's filter run over the inputs of tests/filter_cases.py.
Run in the build container:  python tests/golden/make_golden_filter.py  -> tests/golden/filter_cases.npz"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
import filter_cases as F  # noqa: E402


def main():
    ref = O.ref()
    with tempfile.TemporaryDirectory() as d:
        contigs = F.build_workspace(d, O.REF_BIN)
        (b0, b1), sam_reads = F.reads(contigs, d)
        hg, ht = ref.load_index(os.path.join(d, "gidx")), ref.load_index(os.path.join(d, "tidx"))
        hits, genome_res, pp = F.alignments(ref, hg, ht, b0, b1)
        out = F.run_reference_filter(ref, hg, ht, os.path.join(d, "a.gtf"), os.path.join(d, "golden_out"), sam_reads, hits, genome_res, pp)
        side = {}
        for f in sorted(os.listdir(d)):
            if f.startswith("golden_out"):
                side[f[len("golden_out"):].strip("._") or "main"] = np.frombuffer(open(os.path.join(d, f), "rb").read(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "filter_cases.npz"), result=out, hit_counts0=hits[0][0], hit_counts1=hits[1][0],
                        genome_status=genome_res["status"], genome_location=genome_res["location"],
                        **{"file_" + k: v for k, v in side.items()})
    st = out["status"]
    print("wrote filter_cases.npz:", len(out), "pairs; transcriptome alignments chosen:", int(out["is_transcriptome"].sum()),
          "; not found:", int((st == 0).sum()), "; changed by the filter:", int((out["location"] != genome_res["location"]).any(axis=1).sum()),
          "; side files:", {k: v.size for k, v in side.items()})


if __name__ == "__main__":
    main()
