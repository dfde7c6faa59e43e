"""Pins the oracle: the C restatement (port) and the compiled reference (ref) against the reference's KATs
and the committed golden vectors.  CPU only."""
import pytest

import parity_cases as P


@pytest.fixture(params=["port", "ref"])
def impl(request):
    return request.getfixturevalue(request.param)


@pytest.fixture
def handle(impl, small_index_dir):
    return impl.load_index(small_index_dir)


def test_score_kats(impl):
    P.check_score_kats(impl)


def test_cigar_kats(impl):
    P.check_cigar_kats(impl)


def test_golden_lv(impl, golden):
    P.check_golden_lv(impl, golden)


def test_golden_mapq(impl, golden):
    P.check_golden_mapq(impl, golden)


def test_golden_lookup(impl, handle, golden):
    P.check_golden_lookup(impl, handle, golden)


def test_golden_single(impl, handle, golden):
    P.check_golden_single(impl, handle, golden)


def test_golden_multihit(impl, handle, golden):
    P.check_golden_multihit(impl, handle, golden)


def test_golden_paired(impl, handle, golden):
    P.check_golden_paired(impl, handle, golden)


def test_golden_cigar(impl, handle, golden):
    P.check_golden_cigar(impl, handle, golden)


def test_golden_characterize(impl, handle, golden_characterize):
    P.check_golden_characterize(impl, handle, golden_characterize)


def test_empty(impl, handle):
    P.check_empty(impl, handle)


# ---- ProbabilityDistance: dead code on the alignment path, pinned all the same ------------------------
def _pd(lib, prefix, reference, read, quality, start_shift, max_shift, params=(0.1, 0.01, 0.2), pad=24):
    """The reference's own tests read `reference` outside the literal when shifts are allowed; here it sits in a buffer padded with
    a byte that matches nothing, for the port and the compiled reference alike."""
    import ctypes as C
    buf = C.create_string_buffer(b"\x01" * pad + reference.encode() + b"\x01" * (pad + len(read)))
    ptr = C.cast(C.addressof(buf) + pad, C.c_char_p)
    prob = C.c_double()
    f = getattr(lib, prefix + "probability_distance")
    f.restype = C.c_int
    rc = f(C.c_double(params[0]), C.c_double(params[1]), C.c_double(params[2]), ptr, read.encode(), quality.encode(), C.c_int(len(read)), C.c_int(start_shift),
           C.c_int(max_shift), C.byref(prob))
    assert rc == 5
    return prob.value


def test_probability_distance_kats(impl):
    """The reference's 16 KATs at the reference's own tolerance (ASSERT_NEAR = 1 %), for the restatement and the compiled reference."""
    from golden.lv_kats import PROBABILITY_DISTANCE_KATS
    prefix = "oracle_" if impl.prefix == "oracle_" else "ref_"
    for ref_s, read, qual, ss, ms, want in PROBABILITY_DISTANCE_KATS:
        got = _pd(impl.lib, prefix, ref_s, read, qual, ss, ms)
        assert abs(got - want) <= 0.01 * want, (ref_s, read, ss, ms, got, want)


def test_probability_distance_restatement_agrees_with_the_reference(port, ref):
    """north_star: an exposed ProbabilityDistance score agrees to 1e-6 relative -- restatement against the compiled reference on the
    KATs and on 400 random cases (reads up to 60 bases with substitutions and indels, qualities over the Phred+33 range, shifts up
    to 8, several parameter sets)."""
    import numpy as np
    from golden.lv_kats import PROBABILITY_DISTANCE_KATS
    cases = [(a, b, q, ss, ms, (0.1, 0.01, 0.2)) for a, b, q, ss, ms, _ in PROBABILITY_DISTANCE_KATS]
    rng = np.random.default_rng(5)
    for _ in range(400):
        n = int(rng.integers(1, 60))
        ref_s = "".join(rng.choice(list("ACGT"), size=n + 12))
        read = list(ref_s[:n])
        for _ in range(int(rng.integers(0, 4))):
            p = int(rng.integers(0, len(read)))
            op = rng.random()
            if op < 0.5:
                read[p] = str(rng.choice(list("ACGT")))
            elif op < 0.75:
                read.insert(p, str(rng.choice(list("ACGT"))))
            elif len(read) > 1:
                del read[p]
        read = "".join(read)
        qual = "".join(chr(int(q)) for q in rng.integers(33, 75, size=len(read)))
        ms = int(rng.integers(0, 9))
        ss = int(rng.integers(0, ms + 1))
        prm = [(0.1, 0.01, 0.2), (0.001, 0.001, 0.5), (0.02, 0.005, 0.3)][int(rng.integers(0, 3))]
        cases.append((ref_s, read, qual, ss, ms, prm))
    for ref_s, read, qual, ss, ms, prm in cases:
        a = _pd(port.lib, "oracle_", ref_s, read, qual, ss, ms, prm)
        b = _pd(ref.lib, "ref_", ref_s, read, qual, ss, ms, prm)
        assert abs(a - b) <= 1e-6 * abs(b), (ref_s, read, ss, ms, prm, a, b)


# ---- computeMAPQ where it is last-ulp sensitive (SNAPLib/mapq.h:51: (int)(-10 * log10(1 - pBest / pAll))) --------------------------
def near_integer_mapq_vectors():
    """pBest / pAll such that -10 * log10(1 - ratio) is an integer k = 1..68 to within a few ulps: 1 - ratio = 10^(-k/10) exactly as
    doubles go, and its neighbours up to 3 ulps either side -- where truncation flips between k - 1 and k; with scores either side
    of the `score < 5` rule and popular-seed penalties."""
    import numpy as np
    pa, pb, sc, po = [], [], [], []
    for k in range(1, 69):
        x = 10.0 ** (-k / 10.0)
        for scale in (1.0, 0.37, 1e-9):
            for step in range(-3, 4):
                r = 1.0 - x
                for _ in range(abs(step)):
                    r = np.nextafter(r, 2.0 if step > 0 else -2.0)
                pa.append(scale); pb.append(scale * r); sc.append(3 if (k + step) % 2 else 7); po.append(0 if step else 12)
    for r in (0.9, 0.99, 0.999, 0.9999, 0.5, 0.75, 1.0 - 2.0 ** -52, 1.0, 0.0):
        pa.append(1.0); pb.append(r); sc.append(2); po.append(0)
    return np.array(pa), np.array(pb), np.array(sc, np.int32), np.array(po, np.int32)


def test_mapq_near_integers_restatement_agrees_with_the_reference(port, ref):
    import numpy as np
    pa, pb, sc, po = near_integer_mapq_vectors()
    want = ref.mapq(pa, pb, sc, po)
    np.testing.assert_array_equal(port.mapq(pa, pb, sc, po), want)
    assert len(set(want.tolist())) > 60  # every MAPQ value from 0 to 69 region is hit
