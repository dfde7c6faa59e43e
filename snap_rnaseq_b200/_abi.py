"""ctypes mirror of include/snapb200.h (struct layouts and helpers).

Shared by the product binding (snap_rnaseq_b200/__init__.py) and by the test-only oracle loaders
(oracle/oracle.py) so that the three implementations fill byte-identical result records.
"""
import ctypes as C

import numpy as np

INVALID_LOCATION = 0xFFFFFFFF
MAX_K = 31
MAX_READ_LENGTH = 500
UNUSED_SCORE = 0xFFFF
NOT_FOUND, SINGLE_HIT, MULTIPLE_HITS = 0, 1, 2
FORWARD, RC = 0, 1


class IndexInfo(C.Structure):
    _fields_ = [
        ("n_bases", C.c_uint32), ("n_pieces", C.c_uint32), ("seed_len", C.c_uint32),
        ("n_hash_tables", C.c_uint32), ("overflow_table_size", C.c_uint32),
        ("chromosome_padding", C.c_uint32), ("hash_table_entries", C.c_uint64),
        ("device_bytes", C.c_uint64), ("device", C.c_int32),
    ]


class ReadBatch(C.Structure):
    _fields_ = [
        ("n", C.c_uint32), ("offsets", C.POINTER(C.c_uint32)),
        ("bases", C.POINTER(C.c_uint8)), ("quals", C.POINTER(C.c_uint8)),
    ]


class SingleParams(C.Structure):
    _fields_ = [
        ("max_hits", C.c_uint32), ("max_k", C.c_uint32), ("max_read_size", C.c_uint32),
        ("num_seeds", C.c_uint32), ("seed_coverage", C.c_double), ("extra_search_depth", C.c_uint32),
        ("explore_popular_seeds", C.c_uint32), ("stop_on_first_hit", C.c_uint32),
        ("max_hits_to_get", C.c_uint32),
    ]


class PairedParams(C.Structure):
    _fields_ = [
        ("max_hits", C.c_uint32), ("max_k", C.c_uint32), ("max_read_size", C.c_uint32),
        ("num_seeds", C.c_uint32), ("seed_coverage", C.c_double), ("min_spacing", C.c_uint32),
        ("max_spacing", C.c_uint32), ("force_spacing", C.c_uint32), ("max_big_hits", C.c_uint32),
        ("extra_search_depth", C.c_uint32), ("max_candidate_pool_size", C.c_uint32),
    ]


# numpy record dtypes laid out exactly like the C structs (checked against ctypes sizes below)
SINGLE_RESULT = np.dtype([
    ("location", "<u4"), ("score", "<i4"), ("mapq", "<i4"), ("status", "u1"), ("direction", "u1"),
    ("popular_seeds_skipped", "<u2"), ("n_lookups", "<u4"), ("n_scored", "<u4"),
    ("p_all", "<f8"), ("p_best", "<f8"),
], align=True)

PAIRED_RESULT = np.dtype([
    ("location", "<u4", (2,)), ("score", "<i4", (2,)), ("mapq", "<i4", (2,)), ("status", "u1", (2,)),
    ("direction", "u1", (2,)), ("from_align_together", "u1"), ("aligned_as_pair", "u1"), ("pad", "<u2"),
    ("n_lv_calls", "<u4"), ("n_lookups", "<u4"), ("p_all", "<f8"), ("p_best", "<f8"),
], align=True)

STATS_WORDS = 14 + 71
STATS_FIELDS = ["total_reads", "useful_reads", "single_hits", "multi_hits", "not_found", "errors",
                "aligned_as_pairs", "lv_calls", "n_hash_table_lookups", "n_locations_scored",
                "n_hits_ignored_popularity", "n_reads_ignored_ns", "n_table_probes", "n_hit_words_read"]

assert SINGLE_RESULT.itemsize == 40, SINGLE_RESULT.itemsize
assert PAIRED_RESULT.itemsize == 56, PAIRED_RESULT.itemsize


def single_defaults(**kw):
    """`snap-rna single` defaults: -h 300 -d 14 -n 25 -D 2 (SNAPLib/AlignerOptions.cpp:48-81)."""
    p = SingleParams(max_hits=300, max_k=14, max_read_size=MAX_READ_LENGTH, num_seeds=25, seed_coverage=0.0,
                     extra_search_depth=2, explore_popular_seeds=0, stop_on_first_hit=0, max_hits_to_get=0)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def paired_defaults(**kw):
    """`snap-rna paired` defaults: -h 16000 -H 16000 -d 15 -n 8 -s 50 1000 -mcp 1000000 -D 2
    (SNAPLib/AlignerOptions.cpp:64-81, SNAPLib/PairedAligner.cpp:57-58,229-237)."""
    p = PairedParams(max_hits=16000, max_k=15, max_read_size=MAX_READ_LENGTH, num_seeds=8, seed_coverage=0.0,
                     min_spacing=50, max_spacing=1000, force_spacing=0, max_big_hits=16000,
                     extra_search_depth=2, max_candidate_pool_size=1000000)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Batch:
    """Owns the numpy arrays behind a snapb200_read_batch."""

    def __init__(self, bases, quals, offsets):
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self.quals = np.ascontiguousarray(quals, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        assert self.offsets.ndim == 1 and self.offsets.size >= 1 and self.offsets[0] == 0
        assert self.bases.size == self.quals.size == int(self.offsets[-1])
        self.n = self.offsets.size - 1
        # one spare byte so that a zero-length buffer still has a valid address
        if self.bases.size == 0:
            self._b = np.zeros(1, np.uint8)
            self._q = np.zeros(1, np.uint8)
        else:
            self._b, self._q = self.bases, self.quals
        self.c = ReadBatch(self.n, self.offsets.ctypes.data_as(C.POINTER(C.c_uint32)),
                           self._b.ctypes.data_as(C.POINTER(C.c_uint8)),
                           self._q.ctypes.data_as(C.POINTER(C.c_uint8)))

    @classmethod
    def from_strings(cls, seqs, quals=None):
        if quals is None:
            quals = ["I" * len(s) for s in seqs]
        lens = np.array([len(s) for s in seqs], dtype=np.uint32)
        offsets = np.zeros(len(seqs) + 1, np.uint32)
        np.cumsum(lens, out=offsets[1:])
        b = np.frombuffer("".join(seqs).encode(), dtype=np.uint8) if len(seqs) else np.zeros(0, np.uint8)
        q = np.frombuffer("".join(quals).encode(), dtype=np.uint8) if len(seqs) else np.zeros(0, np.uint8)
        return cls(b.copy(), q.copy(), offsets)

    def read(self, i):
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        return self.bases[a:b].tobytes().decode(), self.quals[a:b].tobytes().decode()

    def slice(self, lo, hi):
        a, b = int(self.offsets[lo]), int(self.offsets[hi])
        return Batch(self.bases[a:b].copy(), self.quals[a:b].copy(), (self.offsets[lo:hi + 1] - self.offsets[lo]).copy())

    def byref(self):
        return C.byref(self.c)


def strings_to_offsets(strs):
    """Concatenate byte strings -> (uint8 array, uint32 offsets[n+1])."""
    lens = np.array([len(s) for s in strs], dtype=np.uint32)
    off = np.zeros(len(strs) + 1, np.uint32)
    np.cumsum(lens, out=off[1:])
    data = np.frombuffer(b"".join(strs), dtype=np.uint8).copy() if len(strs) and off[-1] else np.zeros(1, np.uint8)
    return data, off


def p8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def p32u(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


def p32i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def pf64(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))
