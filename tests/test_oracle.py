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
