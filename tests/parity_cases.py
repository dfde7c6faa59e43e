"""Parity checks shared by every implementation (oracle port, compiled reference, CUDA library).

Each check takes `impl` (a BatchLib-like object with load_index) and compares against the committed golden
outputs of the compiled reference (tests/golden/small_cases.npz) or the reference's own KATs.
"""
import numpy as np

from conftest import assert_records_equal, batch_from
from golden.lv_kats import CIGAR_KATS, SCORE_KATS
from snap_rnaseq_b200 import _abi as A


def check_score_kats(impl):
    # tests/LandauVishkinTest.cpp:11-32 -- the 5-argument form: no quality string, no probability
    texts = [t.encode() for t, _, _, _ in SCORE_KATS]
    pats = [p.encode() for _, p, _, _ in SCORE_KATS]
    ks = [k for _, _, k, _ in SCORE_KATS]
    score, _, _ = impl.lv(1, texts, pats, None, ks)
    assert list(score) == [e for _, _, _, e in SCORE_KATS]
    # the backward instance on reversed text must agree (LandauVishkin<-1> walks the text from its end)
    score_r, _, _ = impl.lv(-1, [t[::-1] for t in texts], pats, None, ks)
    assert list(score_r) == list(score)


def check_cigar_kats(impl):
    # tests/LandauVishkinTest.cpp:34-130
    for use_m in (False, True):
        rows = [r for r in CIGAR_KATS if r[3] == use_m]
        cg, _ = impl.lv_cigar([r[0].encode() for r in rows], [r[1].encode() for r in rows], [r[2] for r in rows], use_m)
        assert cg == [r[4] for r in rows]


def check_golden_lv(impl, g):
    off_t, off_p = g["lv_text_off"], g["lv_pat_off"]
    texts = [g["lv_texts"][off_t[i]:off_t[i + 1]].tobytes() for i in range(len(off_t) - 1)]
    pats = [g["lv_pats"][off_p[i]:off_p[i + 1]].tobytes() for i in range(len(off_p) - 1)]
    quals = [g["lv_quals"][off_p[i]:off_p[i + 1]].tobytes() for i in range(len(off_p) - 1)]
    for d, tag in ((1, "f"), (-1, "r")):
        s, p, ni = impl.lv(d, texts, pats, quals, g["lv_k"])
        np.testing.assert_array_equal(s, g[f"lv{tag}_score"])
        ok = s >= 0
        np.testing.assert_array_equal(p[ok], g[f"lv{tag}_prob"][ok])  # bit-exact doubles
        np.testing.assert_array_equal(ni[ok], g[f"lv{tag}_indel"][ok])


def check_golden_mapq(impl, g):
    out = impl.mapq(g["mapq_pall"], g["mapq_pbest"], g["mapq_score"], g["mapq_pop"])
    np.testing.assert_array_equal(out, g["mapq_out"])


def check_golden_lookup(impl, h, g):
    seeds = [bytes(r) for r in g["lookup_seeds"]]
    nh, hits = impl.lookup(h, seeds, max_out=64)
    np.testing.assert_array_equal(nh, g["lookup_nhits"])
    for i in range(len(seeds)):
        for d in range(2):
            n = min(int(nh[i, d]), 64)
            np.testing.assert_array_equal(hits[i, d, :n], g["lookup_hits"][i, d, :n])


def check_golden_single(impl, h, g):
    b = batch_from(g, "single")
    assert_records_equal(g["single_res"], impl.single(h, A.single_defaults(), b), what="single")
    b = batch_from(g, "long")
    assert_records_equal(g["long_res"], impl.single(h, A.single_defaults(max_k=20), b), what="long reads -d 20")


def check_golden_multihit(impl, h, g):
    b = batch_from(g, "single")
    pm = A.single_defaults(max_hits_to_get=1000, max_hits=16000, num_seeds=8, max_k=15)
    r, cnt, locs, rcs, scores = impl.single_multihit(h, pm, b)
    assert_records_equal(g["multihit_res"], r, what="multihit")
    np.testing.assert_array_equal(cnt, g["multihit_cnt"])
    for i in range(b.n):
        n = int(cnt[i])
        np.testing.assert_array_equal(locs[i, :n], g["multihit_locs"][i, :n])
        np.testing.assert_array_equal(rcs[i, :n], g["multihit_rcs"][i, :n])
        np.testing.assert_array_equal(scores[i, :n], g["multihit_scores"][i, :n])


def check_golden_paired(impl, h, g):
    b0, b1 = batch_from(g, "pair0"), batch_from(g, "pair1")
    assert_records_equal(g["paired_res"], impl.paired(h, A.paired_defaults(), b0, b1), what="paired")


def check_golden_cigar(impl, h, g):
    b = batch_from(g, "single")
    res = g["single_res"]
    for use_m in (0, 1):
        cg, ed = impl.cigar(h, b, res["location"], res["direction"], use_m)
        np.testing.assert_array_equal(ed, g[f"cigar{use_m}_ed"])
        assert cg == [str(s) for s in g[f"cigar{use_m}_str"]]


def check_empty(impl, h):
    e = A.Batch.from_strings([])
    assert len(impl.single(h, A.single_defaults(), e)) == 0
    assert len(impl.paired(h, A.paired_defaults(), e, e)) == 0
    seg, locs, offs = impl.characterize(h, A.single_defaults(), e)
    assert list(seg) == [0] and len(locs) == 0 and len(offs) == 0


# BaseAligner::CharacterizeSeeds (BaseAligner.cpp:206-508): golden = the compiled reference's seed maps, flattened
CHARACTERIZE_CASES = {
    "partial": ("r100", dict(max_hits=300, num_seeds=12, max_k=15)),
    "popular": ("r100", dict(max_hits=4, num_seeds=12, max_k=15)),
    "explore": ("r150", dict(max_hits=3, num_seeds=20, max_k=15, explore_popular_seeds=1)),
    "coverage": ("r150", dict(max_hits=300, num_seeds=0, seed_coverage=2.5, max_k=8)),
}


def check_golden_characterize(impl, h, gc):
    for name, (rs, kw) in CHARACTERIZE_CASES.items():
        b = batch_from(gc, rs)
        seg, locs, offs = impl.characterize(h, A.single_defaults(**kw), b)
        np.testing.assert_array_equal(seg, gc[name + "_seg"], err_msg=name)
        np.testing.assert_array_equal(locs, gc[name + "_locs"], err_msg=name)
        np.testing.assert_array_equal(offs, gc[name + "_offs"], err_msg=name)
        # properties of a flattened std::map<location, std::set<offset>>: strictly ascending (location, offset) per segment
        key = locs.astype(np.uint64) * 512 + offs
        starts = np.zeros(key.size, bool)
        starts[seg[:-1][seg[:-1] < key.size].astype(np.int64)] = True
        assert np.all((np.diff(key.astype(np.int64)) > 0) | starts[1:]), name
