#!/usr/bin/env python3
"""Golden vectors for BaseAligner::CharacterizeSeeds, generated from the COMPILED REFERENCE (oracle/_ref).

Run in the build container (needs oracle/_ref, i.e. /root/reference):
    python tests/golden/make_golden_characterize.py
Reads tests/golden/small_index.tar.gz (the reference-built index of make_golden.py) and writes
characterize_cases.npz next to this file: two read sets plus, per parameter set, the reference's two seed maps per read
flattened in iteration order (segment 2*i+dir: ascending location, then ascending seed offset).
"""
import os
import sys
import tarfile
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
from snap_rnaseq_b200 import _abi as A  # noqa: E402
from snap_rnaseq_b200 import synth  # noqa: E402
from tests_genome import small_genome  # noqa: E402

# name -> (read set, snapb200_single_params overrides); "partial" is the partialAligner of PairedAligner.cpp:518-527
CASES = {
    "partial": ("r100", dict(max_hits=300, num_seeds=12, max_k=15)),
    "popular": ("r100", dict(max_hits=4, num_seeds=12, max_k=15)),
    "explore": ("r150", dict(max_hits=3, num_seeds=20, max_k=15, explore_popular_seeds=1)),
    "coverage": ("r150", dict(max_hits=300, num_seeds=0, seed_coverage=2.5, max_k=8)),
}


def main():
    if not O.have_ref():
        raise SystemExit("oracle/_ref missing: run python oracle/build_ref.py first")
    ref = O.ref()
    contigs = small_genome()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        with tarfile.open(os.path.join(HERE, "small_index.tar.gz")) as tf:
            tf.extractall(tmp)
        h = ref.load_index(os.path.join(tmp, "small_index"))
        sets = {
            "r100": synth.simulate(contigs, 600, 100, paired=False, err=0.03, seed=301, junk_frac=0.05, n_rate=0.03)["batches"][0],
            "r150": synth.simulate(contigs, 400, 150, paired=False, err=0.02, seed=302, junk_frac=0.05, n_rate=0.03)["batches"][0],
        }
        for name, b in sets.items():
            out[name + "_bases"], out[name + "_quals"], out[name + "_offsets"] = b.bases, b.quals, b.offsets
        for name, (rs, kw) in CASES.items():
            seg, locs, offs = ref.characterize(h, A.single_defaults(**kw), sets[rs])
            out[name + "_seg"], out[name + "_locs"], out[name + "_offs"] = seg, locs, offs
            print(name, rs, kw, int(seg[-1]), "tuples")
    np.savez_compressed(os.path.join(HERE, "characterize_cases.npz"), **out)


if __name__ == "__main__":
    main()
