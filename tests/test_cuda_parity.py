"""Parity tests proper: the CUDA library, through its C ABI, against the reference's KATs, the committed golden
outputs of the compiled reference, and the oracle (port, and the reference itself when oracle/_ref is present) on
fresh seeded inputs.  Everything here needs a GPU."""
import os

import numpy as np
import pytest

import parity_cases as P
from conftest import assert_records_equal
from snap_rnaseq_b200 import _abi as A
from snap_rnaseq_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle(cuda, small_index_dir):
    h = cuda.load_index(small_index_dir)
    yield h
    cuda.close_index(h)


def test_score_kats(cuda):
    P.check_score_kats(cuda)


def test_cigar_kats(cuda):
    P.check_cigar_kats(cuda)


def test_golden_lv(cuda, golden):
    P.check_golden_lv(cuda, golden)


def test_golden_mapq(cuda, golden):
    P.check_golden_mapq(cuda, golden)


def test_mapq_near_integers_force_the_host_reevaluation(cuda, ref):
    """computeMAPQ truncates -10 * log10(1 - pBest / pAll) (mapq.h:51): CUDA's log10 and glibc's may differ in the last ulp, which
    matters exactly when the value is within ulps of an integer.  The kernels flag such values and the library re-evaluates them
    with libm; these vectors sit on and around every integer 1..68, so the path is taken (and the results are the reference's)."""
    import ctypes as C
    from test_oracle import near_integer_mapq_vectors
    pa, pb, sc, po = near_integer_mapq_vectors()
    want = ref.mapq(pa, pb, sc, po)
    np.testing.assert_array_equal(cuda.mapq(pa, pb, sc, po), want)
    out, flags = np.zeros(len(pa), np.int32), np.zeros(len(pa), np.uint8)
    rc = cuda.lib.snapb200_mapq_batch_ex(C.c_int(cuda.device), C.c_uint32(len(pa)), A.pf64(pa), A.pf64(pb), A.p32i(sc), A.p32i(po), A.p32i(out), A.p8(flags))
    assert rc == 0
    np.testing.assert_array_equal(out, want)
    assert flags.sum() > 300, int(flags.sum())  # the near-integer vectors were re-evaluated on the host


def test_golden_lookup(cuda, handle, golden):
    P.check_golden_lookup(cuda, handle, golden)


def test_golden_single(cuda, handle, golden):
    P.check_golden_single(cuda, handle, golden)


def test_golden_multihit(cuda, handle, golden):
    P.check_golden_multihit(cuda, handle, golden)


def test_golden_paired(cuda, handle, golden):
    P.check_golden_paired(cuda, handle, golden)


def test_golden_cigar(cuda, handle, golden):
    P.check_golden_cigar(cuda, handle, golden)


def test_golden_characterize(cuda, handle, golden_characterize):
    P.check_golden_characterize(cuda, handle, golden_characterize)


def test_characterize_fresh_and_edges(cuda, handle, port, small_index_dir, request):
    """CharacterizeSeeds on fresh reads (incl. repeat-family reads with thousands of tuples), ragged/edge reads, the
    multi-chunk path, the count-only call and the capacity check."""
    from tests_genome import small_genome
    contigs = small_genome()
    c1 = contigs["chr1"].tobytes().decode()
    edge = A.Batch.from_strings(["", "A", c1[:19], c1[:20], c1[:49], c1[-100:], "N" * 80, c1[5:65] + "N" * 10 + c1[75:125],
                                 c1[300:400][::-1], c1[1000:1100]])
    sim = synth.simulate(contigs, 3000, 100, paired=True, err=0.03, seed=55, junk_frac=0.05, n_rate=0.02)
    b0, b1 = sim["batches"]
    for name, chk in _checkers(port, request):
        hc = chk.load_index(small_index_dir)
        for b, kw in ((edge, dict(max_hits=300, num_seeds=12)), (b0, dict(max_hits=300, num_seeds=12)),
                      (b1, dict(max_hits=16000, num_seeds=25, max_k=14)), (b0, dict(max_hits=2, num_seeds=40, explore_popular_seeds=1))):
            ps = A.single_defaults(**kw)
            want, got = chk.characterize(hc, ps, b), cuda.characterize(handle, ps, b)
            for w, g, what in zip(want, got, ("seg_offsets", "locations", "seed_offsets")):
                np.testing.assert_array_equal(w, g, err_msg=f"{what} vs {name} {kw}")
    ps = A.single_defaults(max_hits=300, num_seeds=12)
    full = cuda.characterize(handle, ps, b0)
    os.environ["SNAPB200_CHUNK"] = "1024"  # three internal launch groups
    try:
        chunked = cuda.characterize(handle, ps, b0)
    finally:
        del os.environ["SNAPB200_CHUNK"]
    for w, g in zip(full, chunked):
        np.testing.assert_array_equal(w, g)
    # capacity too small is an argument error, not a truncated answer
    import ctypes as C
    seg = np.zeros(2 * b0.n + 1, np.uint64)
    locs = np.zeros(4, np.uint32)
    offs = np.zeros(4, np.uint16)
    rc = cuda.lib.snapb200_characterize_batch(handle, C.byref(ps), b0.byref(), seg.ctypes.data_as(C.c_void_p), A.p32u(locs),
                                              offs.ctypes.data_as(C.c_void_p), C.c_uint64(4))
    assert rc == -1 and int(seg[-1]) == int(full[0][-1])


def test_empty(cuda, handle):
    P.check_empty(cuda, handle)


def test_index_info(cuda, handle, port, small_index_dir):
    a, b = cuda.index_info(handle), port.index_info(port.load_index(small_index_dir))
    for f in ("n_bases", "n_pieces", "seed_len", "n_hash_tables", "overflow_table_size", "chromosome_padding", "hash_table_entries"):
        assert getattr(a, f) == getattr(b, f), f
    assert a.device_bytes > a.n_bases


# ---- differential tests on fresh inputs (port always; compiled reference too when it is on the box) ----
def _checkers(port, request):
    out = [("port", port)]
    from oracle import oracle as O
    if O.have_ref():
        out.append(("ref", O.ref()))
    return out


@pytest.mark.parametrize("rlen,err,max_k,seed", [(100, 0.02, 15, 31), (150, 0.01, 15, 32), (250, 0.04, 20, 33), (62, 0.05, 15, 34)])
def test_fresh_paired_and_single(cuda, handle, port, small_index_dir, request, rlen, err, max_k, seed):
    from tests_genome import small_genome
    contigs = small_genome()
    sim = synth.simulate(contigs, 3000, rlen, paired=True, err=err, seed=seed, junk_frac=0.03, n_rate=0.02,
                         frag=(max(250, rlen + 20), max(450, rlen + 200)))
    b0, b1 = sim["batches"]
    pp = A.paired_defaults(max_k=max_k)
    ps = A.single_defaults(max_k=min(max_k, 20))
    got_p = cuda.paired(handle, pp, b0, b1)
    got_s = cuda.single(handle, ps, b1)
    for name, chk in _checkers(port, request):
        hc = chk.load_index(small_index_dir)
        assert_records_equal(chk.paired(hc, pp, b0, b1), got_p, what=f"paired vs {name}")
        assert_records_equal(chk.single(hc, ps, b1), got_s, what=f"single vs {name}")
        cg, ed = cuda.cigar(handle, b1, got_s["location"], got_s["direction"], False)
        cw, ew = chk.cigar(hc, b1, got_s["location"], got_s["direction"], False)
        assert cg == cw and np.array_equal(ed, ew)


def test_parameter_variants(cuda, handle, port, small_index_dir):
    """-x, -f, seed coverage instead of seed count, small -h (popular seeds skipped), -fs."""
    from tests_genome import small_genome
    contigs = small_genome()
    sim = synth.simulate(contigs, 2000, 100, paired=True, err=0.03, seed=41, junk_frac=0.05)
    b0, b1 = sim["batches"]
    hp = port.load_index(small_index_dir)
    for kw in (dict(explore_popular_seeds=1, max_hits=4), dict(stop_on_first_hit=1), dict(num_seeds=0, seed_coverage=2.0),
               dict(max_hits=2), dict(max_k=8, extra_search_depth=1), dict(num_seeds=60)):
        ps = A.single_defaults(**kw)
        assert_records_equal(port.single(hp, ps, b0), cuda.single(handle, ps, b0), what=f"single {kw}")
    for kw in (dict(force_spacing=1), dict(min_spacing=0, max_spacing=300), dict(max_big_hits=3), dict(num_seeds=0, seed_coverage=1.5),
               dict(max_hits=2), dict(num_seeds=25), dict(max_candidate_pool_size=64)):
        pp = A.paired_defaults(**kw)
        try:
            want = port.paired(hp, pp, b0, b1)
        except RuntimeError:
            want = None  # the reference would exit: pool exhausted
        if want is None:
            with pytest.raises(RuntimeError):
                cuda.paired(handle, pp, b0, b1)
        else:
            assert_records_equal(want, cuda.paired(handle, pp, b0, b1), what=f"paired {kw}")


def test_ragged_and_edge_reads(cuda, handle, port, small_index_dir):
    """Ragged lengths incl. empty, < seedLen, < 50, all-N, and reads at the very start/end of contigs."""
    from tests_genome import small_genome
    contigs = small_genome()
    c1 = contigs["chr1"].tobytes().decode()
    c3 = contigs["chr3"].tobytes().decode()
    seqs = ["", "A", c1[:19], c1[:20], c1[:49], c1[:50], c1[:100], c1[-100:], c3[-75:], c3[-60:] + "ACGTACGTAC", "N" * 80,
            c1[5:65] + "N" * 10 + c1[75:125], c1[200:300][::-1], c1[1000:1100], c1[1000:1040] + c1[1043:1103], c1[2000:2050] + "GG" + c1[2050:2098]]
    b0 = A.Batch.from_strings(seqs)
    b1 = A.Batch.from_strings(list(reversed(seqs)))
    hp = port.load_index(small_index_dir)
    assert_records_equal(port.single(hp, A.single_defaults(), b0), cuda.single(handle, A.single_defaults(), b0), what="ragged single")
    assert_records_equal(port.paired(hp, A.paired_defaults(), b0, b1), cuda.paired(handle, A.paired_defaults(), b0, b1), what="ragged paired")


def test_scratch_tier_overflow_is_invisible(cuda, handle, port, small_index_dir):
    """Reads drawn from the repeat families overflow the small scratch tier and are rerun in the large one; results
    must not depend on that."""
    from tests_genome import small_genome
    contigs = small_genome()
    sim = synth.simulate(contigs, 1500, 100, paired=True, err=0.01, seed=77)
    b0, b1 = sim["batches"]
    hp = port.load_index(small_index_dir)
    ps = A.single_defaults(max_hits=16000, num_seeds=8, max_k=15)
    assert_records_equal(port.single(hp, ps, b0), cuda.single(handle, ps, b0), what="large maxHits single")


def test_stats_and_session(cuda, handle, golden):
    import snap_rnaseq_b200 as S
    from conftest import batch_from
    b0, b1 = batch_from(golden, "pair0"), batch_from(golden, "pair1")
    cuda.stats_reset(handle)
    sess = S.Session(cuda, handle, b0.n, 128)
    sess.upload(0, b0)
    sess.upload(1, b1)
    sess.run_paired(A.paired_defaults())
    out = np.zeros(b0.n, A.PAIRED_RESULT)
    sess.download_paired(out)
    assert_records_equal(golden["paired_res"], out, what="session paired")
    ms, launches, total = sess.last_run()
    assert ms > 0 and launches >= 2 and total >= launches
    w = cuda.stats(handle)
    assert w[0] == 2 * b0.n                      # total_reads
    assert w[2] + w[3] + w[4] == 2 * b0.n        # single + multi + not found
    st = golden["paired_res"]["status"]
    assert w[2] == int((st == 1).sum()) and w[3] == int((st == 2).sum()) and w[4] == int((st == 0).sum())
    sess.close()


def test_large_batch_properties(cuda, handle):
    """At a size the oracle would take too long for: chunking/order independence and determinism."""
    from tests_genome import small_genome
    contigs = small_genome()
    n = 300_000
    sim = synth.simulate(contigs, n, 100, paired=True, err=0.02, seed=5)
    b0, b1 = sim["batches"]
    pp = A.paired_defaults()
    os.environ["SNAPB200_CHUNK"] = "262144"  # two internal launch groups
    try:
        full = cuda.paired(handle, pp, b0, b1)
    finally:
        del os.environ["SNAPB200_CHUNK"]
    again = cuda.paired(handle, pp, b0, b1)
    assert_records_equal(full, again, what="determinism")
    lo, hi = 262000, 262400  # straddles the chunk boundary
    part = cuda.paired(handle, pp, b0.slice(lo, hi), b1.slice(lo, hi))
    assert_records_equal(full[lo:hi], part, what="chunk independence")
    # swapping the mates swaps the per-end fields
    sw = cuda.paired(handle, pp, b1.slice(0, 2000), b0.slice(0, 2000))
    for f in ("location", "score", "mapq", "status", "direction"):
        assert np.array_equal(sw[f][:, ::-1], full[f][:2000]), f
    ok = full["status"][:, 0] != 0
    assert ok.mean() > 0.9


# ---- medium-sized differential run: large enough for the work-ordering pass (>= 4096 items), lane-mode batches of 32, the
# order array and the scratch tiers to be exercised against the reference itself ----
@pytest.fixture(scope="module")
def medium(cuda, tmp_path_factory):
    contigs = synth.random_contigs([900_000, 700_000, 400_000], seed=120)
    synth.inject_repeats(contigs, frac=0.08, seed=121, min_len=150, max_len=2000, max_copies=300)
    bases, offs = synth.snap_layout(contigs, 500)
    h = cuda.build_index(bases, offs, list(contigs), seed_len=20)
    d = tmp_path_factory.mktemp("medium") / "idx"
    cuda.save_index(h, d)
    yield contigs, h, str(d)
    cuda.close_index(h)


def _best_checker(port):
    from oracle import oracle as O
    return ("ref", O.ref(threads=8)) if O.have_ref() else ("port", port)


@pytest.mark.parametrize("rlen,err,max_k,n", [(150, 0.01, 15, 24000), (100, 0.03, 15, 16000), (250, 0.04, 20, 6000)])
def test_medium_genome_paired(cuda, port, medium, rlen, err, max_k, n):
    contigs, h, d = medium
    sim = synth.simulate(contigs, n, rlen, paired=True, err=err, indel_frac=0.15, seed=rlen, junk_frac=0.01, n_rate=0.01,
                         frag=(max(250, rlen + 20), max(450, rlen + 200)))
    b0, b1 = sim["batches"]
    pp = A.paired_defaults(max_k=max_k)
    got = cuda.paired(h, pp, b0, b1)
    name, chk = _best_checker(port)
    want = chk.paired(chk.load_index(d), pp, b0, b1)
    assert_records_equal(want, got, what=f"medium paired {rlen}bp vs {name}")
    assert int(got["n_lv_calls"].max()) > 200  # repeat-family pairs are in the set


def test_medium_genome_single_multihit_characterize(cuda, port, medium):
    contigs, h, d = medium
    sim = synth.simulate(contigs, 12000, 100, paired=True, err=0.02, indel_frac=0.15, seed=77, junk_frac=0.02, n_rate=0.01)
    b0, b1 = sim["batches"]
    name, chk = _best_checker(port)
    hc = chk.load_index(d)
    ps = A.single_defaults()
    assert_records_equal(chk.single(hc, ps, b0), cuda.single(h, ps, b0), what=f"medium single vs {name}")
    pm = A.single_defaults(max_hits_to_get=1000, max_hits=16000, num_seeds=8, max_k=15)  # transcriptomeAligner, PairedAligner.cpp:512
    want, got = chk.single_multihit(hc, pm, b1), cuda.single_multihit(h, pm, b1)
    assert_records_equal(want[0], got[0], what=f"medium multihit vs {name}")
    np.testing.assert_array_equal(want[1], got[1])
    for i in range(b1.n):
        k = int(want[1][i])
        assert np.array_equal(want[2][i, :k], got[2][i, :k]) and np.array_equal(want[3][i, :k], got[3][i, :k]) and np.array_equal(want[4][i, :k], got[4][i, :k]), i
    pc = A.single_defaults(max_hits=300, num_seeds=12, max_k=15)  # partialAligner, PairedAligner.cpp:518-527
    for w, g in zip(chk.characterize(hc, pc, b0), cuda.characterize(h, pc, b0)):
        np.testing.assert_array_equal(w, g)


# ---- device-side index construction (lookup-equivalent to GenomeIndex::BuildIndexToDirectory) ----
@pytest.mark.parametrize("seed_len", [20, 16, 23])
def test_index_build_equivalence(cuda, port, tmp_path, seed_len, golden, small_index_dir):
    from tests_genome import small_genome
    contigs = small_genome()
    bases, offs = synth.snap_layout(contigs, 500)
    h = cuda.build_index(bases, offs, list(contigs), seed_len=seed_len)
    try:
        info = cuda.index_info(h)
        assert info.n_bases == bases.size and info.seed_len == seed_len
        d = tmp_path / f"built_{seed_len}"
        cuda.save_index(h, d)
        # the saved directory must be loadable by the CPU checkers, and lookups must agree with the device
        rng = np.random.default_rng(seed_len)
        flat = np.concatenate(list(contigs.values()))
        seeds = [flat[o:o + seed_len].tobytes() for o in rng.integers(0, flat.size - seed_len, 400)]
        seeds += [synth.BASES[rng.integers(0, 4, size=seed_len)].tobytes() for _ in range(100)]
        nh_c, hits_c = cuda.lookup(h, seeds, 64)
        checkers = [("port", port)]
        from oracle import oracle as O
        if O.have_ref():
            checkers.append(("ref", O.ref()))
        sim = synth.simulate(contigs, 1500, 100, paired=True, err=0.03, seed=91, junk_frac=0.03)
        b0, b1 = sim["batches"]
        got = cuda.paired(h, A.paired_defaults(), b0, b1)
        for name, chk in checkers:
            hc = chk.load_index(str(d))
            nh, hits = chk.lookup(hc, seeds, 64)
            np.testing.assert_array_equal(nh, nh_c)
            np.testing.assert_array_equal(hits, hits_c)
            assert_records_equal(chk.paired(hc, A.paired_defaults(), b0, b1), got, what=f"paired on device-built index vs {name}")
        if seed_len == 20:
            # same genome as the reference-built golden index: identical lookup answers and alignments
            P.check_golden_lookup(cuda, h, golden)
            P.check_golden_paired(cuda, h, golden)
            P.check_golden_single(cuda, h, golden)
    finally:
        cuda.close_index(h)


def test_concurrent_callers(cuda, handle, golden):
    """Two host threads in the C ABI at once (as the reference's -t N worker threads would be): each gets its own
    internal session; results must be those of the serial calls."""
    import threading

    from conftest import batch_from
    b0, b1 = batch_from(golden, "pair0"), batch_from(golden, "pair1")
    bs = batch_from(golden, "single")
    out = {}

    def work(tag):
        for i in range(3):
            out[(tag, i, "p")] = cuda.paired(handle, A.paired_defaults(), b0, b1)
            out[(tag, i, "s")] = cuda.single(handle, A.single_defaults(), bs)

    th = [threading.Thread(target=work, args=(t,)) for t in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert len(out) == 18
    for k, v in out.items():
        assert_records_equal(golden["paired_res" if k[2] == "p" else "single_res"], v, what=f"concurrent {k}")
