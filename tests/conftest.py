import os
import sys
import tarfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "small_cases.npz"))


@pytest.fixture(scope="session")
def golden_characterize():
    return np.load(os.path.join(GOLDEN, "characterize_cases.npz"))


@pytest.fixture(scope="session")
def small_index_dir(tmp_path_factory):
    d = tmp_path_factory.mktemp("golden_index")
    with tarfile.open(os.path.join(GOLDEN, "small_index.tar.gz")) as tf:
        tf.extractall(d)
    return str(d / "small_index")


@pytest.fixture(scope="session")
def port():
    from oracle import oracle as O
    return O.port()


@pytest.fixture(scope="session")
def ref():
    from oracle import oracle as O
    if not O.have_ref():
        pytest.skip("oracle/_ref not built in this environment")
    return O.ref()


@pytest.fixture(scope="session")
def cuda():
    """The product: libsnapb200.so through its C ABI.  Fails (not skips) if the library is missing."""
    import snap_rnaseq_b200 as S
    return S.lib()


def batch_from(g, prefix):
    from snap_rnaseq_b200._abi import Batch
    return Batch(g[prefix + "_bases"], g[prefix + "_quals"], g[prefix + "_offsets"])


def assert_records_equal(a, b, fields=None, what=""):
    """Bit-exact comparison of result records; NaN in `a` (reference has no value) matches anything."""
    names = fields or a.dtype.names
    assert len(a) == len(b)
    bad = np.zeros(len(a), bool)
    for f in names:
        x, y = a[f], b[f]
        if x.dtype.kind == "f":
            neq = ~((x == y) | np.isnan(x))
        else:
            neq = x != y
        if neq.ndim > 1:
            neq = neq.any(axis=1)
        bad |= neq
    idx = np.nonzero(bad)[0]
    if idx.size:
        lines = [f"{what}: {idx.size} of {len(a)} records differ"]
        for i in idx[:5]:
            lines.append(f"  [{i}] expected {a[i]}\n       got      {b[i]}")
        raise AssertionError("\n".join(lines))
