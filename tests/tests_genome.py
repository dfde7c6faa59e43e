"""The genome behind tests/golden/small_index.tar.gz, regenerated from its seeds (must match make_golden.py)."""
from snap_rnaseq_b200 import synth


def small_genome():
    contigs = synth.random_contigs([16000, 12000, 8000], seed=20)
    synth.inject_repeats(contigs, frac=0.10, seed=21, min_len=100, max_len=600, max_copies=40)
    return contigs
