"""AlignmentFilter (SURVEY.md section 8 row f3 -- the next row; there is no CUDA version yet): the oracle entry point
ref_filter_paired_batch (oracle/ref_driver.cpp: the reference's own AlignmentFilter driven as PairedAligner.cpp:575-663 drives it)
is pinned by tests/golden/filter_cases.npz, and gives the same records whether its inputs come from the compiled reference's
aligners or from the C restatement's -- so the device version will be checked against a fixed target from its first line."""
import os

import numpy as np
import pytest

import filter_cases as F
from conftest import assert_records_equal

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden_filter():
    return np.load(os.path.join(HERE, "golden", "filter_cases.npz"))


def test_reference_filter_reproduces_golden(ref, port, golden_filter, tmp_path):
    from oracle import oracle as O
    d = str(tmp_path)
    contigs = F.build_workspace(d, O.REF_BIN)
    (b0, b1), sam_reads = F.reads(contigs, d)
    hg, ht = ref.load_index(os.path.join(d, "gidx")), ref.load_index(os.path.join(d, "tidx"))
    hits, genome_res, pp = F.alignments(ref, hg, ht, b0, b1)
    assert np.array_equal(hits[0][0], golden_filter["hit_counts0"]) and np.array_equal(hits[1][0], golden_filter["hit_counts1"])
    assert np.array_equal(genome_res["location"], golden_filter["genome_location"])
    out = F.run_reference_filter(ref, hg, ht, os.path.join(d, "a.gtf"), os.path.join(d, "run1"), sam_reads, hits, genome_res, pp)
    assert_records_equal(golden_filter["result"], out, what="AlignmentFilter vs golden")
    for f in sorted(os.listdir(d)):
        if f.startswith("run1"):
            key = "file_" + (f[len("run1"):].strip("._") or "main")
            assert open(os.path.join(d, f), "rb").read() == golden_filter[key].tobytes(), f
    # the filter's inputs from the C restatement of the aligners are the same, hence its outputs
    hp, htp = port.load_index(os.path.join(d, "gidx")), port.load_index(os.path.join(d, "tidx"))
    hits_p, genome_p, _ = F.alignments(port, hp, htp, b0, b1)
    for e in range(2):
        n = hits[e][0]
        assert np.array_equal(n, hits_p[e][0])
        m = np.arange(F.MAX_HITS_TO_GET)[None, :] < n[:, None]
        for a, b in zip(hits[e][1:], hits_p[e][1:]):
            assert np.array_equal(a[m], b[m])
    assert_records_equal(genome_res, genome_p, fields=("location", "score", "mapq", "status", "direction"), what="genome pair, port vs reference")
    # what the cases exercise
    r = golden_filter["result"]
    assert r["is_transcriptome"].sum() > 100 and (r["status"] == 0).sum() > 50
    assert (r["location"] != golden_filter["genome_location"]).any(axis=1).sum() > 100
