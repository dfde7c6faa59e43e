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


# ---- row f2: FASTQ parse / SAM text ---------------------------------------------------------------------------
class SamReadsStruct(C.Structure):
    _fields_ = [
        ("n", C.c_uint32), ("offsets", C.POINTER(C.c_uint32)), ("bases", C.POINTER(C.c_uint8)),
        ("quals", C.POINTER(C.c_uint8)), ("front_clip", C.POINTER(C.c_uint16)), ("clipped_len", C.POINTER(C.c_uint16)),
        ("id_offsets", C.POINTER(C.c_uint32)), ("ids", C.POINTER(C.c_uint8)),
    ]


SAM_ALIGNMENT = np.dtype([("location", "<u4"), ("mapq", "<i4"), ("status", "u1"), ("direction", "u1"), ("skip", "u1"),
                          ("is_transcriptome", "u1"), ("tlocation", "<u4")], align=True)
assert SAM_ALIGNMENT.itemsize == 16


def p16u(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint16))


class SamReads:
    """Owns the numpy arrays behind a snapb200_sam_reads: unclipped reads + clipping + ids."""

    def __init__(self, offsets, bases, quals, front_clip, clipped_len, id_offsets, ids):
        self.offsets = np.ascontiguousarray(offsets, np.uint32)
        self.n = self.offsets.size - 1
        pad = lambda a: a if a.size else np.zeros(1, a.dtype)
        self.bases = pad(np.ascontiguousarray(bases, np.uint8))
        self.quals = pad(np.ascontiguousarray(quals, np.uint8))
        self.front_clip = pad(np.ascontiguousarray(front_clip, np.uint16))
        self.clipped_len = pad(np.ascontiguousarray(clipped_len, np.uint16))
        self.id_offsets = np.ascontiguousarray(id_offsets, np.uint32)
        self.ids = pad(np.ascontiguousarray(ids, np.uint8))
        self.c = SamReadsStruct(self.n, p32u(self.offsets), p8(self.bases), p8(self.quals), p16u(self.front_clip),
                                p16u(self.clipped_len), p32u(self.id_offsets), p8(self.ids))

    @classmethod
    def from_lists(cls, ids, seqs, quals, front_clip=None, clipped_len=None):
        b, off = strings_to_offsets([s.encode() if isinstance(s, str) else s for s in seqs])
        q, _ = strings_to_offsets([s.encode() if isinstance(s, str) else s for s in quals])
        i, ioff = strings_to_offsets([s.encode() if isinstance(s, str) else s for s in ids])
        lens = np.diff(off).astype(np.uint16)
        fc = np.zeros(len(seqs), np.uint16) if front_clip is None else np.asarray(front_clip, np.uint16)
        cl = (lens - fc) if clipped_len is None else np.asarray(clipped_len, np.uint16)
        return cls(off, b[:off[-1]], q[:off[-1]], fc, cl, ioff, i[:ioff[-1]])

    def byref(self):
        return C.byref(self.c)

    def clipped_batch(self):
        """The reads as the aligners take them (Read::getData/getQuality/getDataLength)."""
        lens = self.clipped_len[:self.n].astype(np.uint32)
        if self.n and not self.front_clip[:self.n].any() and np.array_equal(lens, np.diff(self.offsets)) and self.offsets[0] == 0:
            return Batch(self.bases[:self.offsets[-1]], self.quals[:self.offsets[-1]], self.offsets)
        off = np.zeros(self.n + 1, np.uint32)
        np.cumsum(lens, out=off[1:])
        src = (self.offsets[:-1] + self.front_clip[:self.n]).astype(np.int64)
        idx = np.repeat(src - off[:-1].astype(np.int64), lens) + np.arange(int(off[-1]), dtype=np.int64)
        return Batch(self.bases[idx], self.quals[idx], off)

    def same_as(self, o):
        n = self.n
        return (n == o.n and np.array_equal(self.offsets, o.offsets) and np.array_equal(self.id_offsets, o.id_offsets)
                and np.array_equal(self.bases[:self.offsets[-1]], o.bases[:o.offsets[-1]])
                and np.array_equal(self.quals[:self.offsets[-1]], o.quals[:o.offsets[-1]])
                and np.array_equal(self.ids[:self.id_offsets[-1]], o.ids[:o.id_offsets[-1]])
                and np.array_equal(self.front_clip[:n], o.front_clip[:n]) and np.array_equal(self.clipped_len[:n], o.clipped_len[:n]))


# ---- row f3: AlignmentFilter -----------------------------------------------------------------------------------
class FilterParams(C.Structure):
    _fields_ = [("max_spacing", C.c_uint32), ("force_spacing", C.c_uint32), ("conf_diff", C.c_uint32), ("max_dist", C.c_uint32),
                ("max_hits_to_get", C.c_uint32)]


FILTER_RESULT = np.dtype([("location", "<u4", (2,)), ("tlocation", "<u4", (2,)), ("score", "<i4", (2,)), ("mapq", "<i4", (2,)),
                          ("status", "u1", (2,)), ("direction", "u1", (2,)), ("is_transcriptome", "u1", (2,)), ("aligned_as_pair", "u1"), ("pad", "u1")])
FILTER_EVENT = np.dtype([("kind", "<i4"), ("unaligned", "<i4"), ("transcript", "<i4", (2,)), ("chr", "<i4", (2,)), ("pos_original", "<u4", (2,)),
                         ("pos", "<u4", (2,)), ("pos_end", "<u4", (2,))])
assert FILTER_RESULT.itemsize == 40 and FILTER_EVENT.itemsize == 48


class RnaParams(C.Structure):
    """snapb200_rna_params: the aligner objects of one worker thread of the paired RNA loop (SNAPLib/PairedAligner.cpp:459-527)."""
    _fields_ = [("paired", PairedParams), ("transcriptome", SingleParams), ("partial", SingleParams), ("filter", FilterParams)]


class RnaView(C.Structure):
    """snapb200_rna_view: pointers into the batch object's pinned host memory."""
    _fields_ = [("n", C.c_uint32), ("results", C.c_void_p), ("events", C.c_void_p), ("needs_host", C.c_void_p), ("genome_pairs", C.c_void_p),
                ("hit_offsets", C.c_void_p * 2), ("hit_locations", C.c_void_p * 2), ("hit_rcs", C.c_void_p * 2), ("hit_scores", C.c_void_p * 2),
                ("seg_offsets", C.c_void_p * 2), ("ch_locations", C.c_void_p * 2), ("ch_seed_offsets", C.c_void_p * 2),
                ("splice_offsets", C.c_void_p), ("splices", C.c_void_p), ("splice_overflow", C.c_void_p), ("device_ms", C.c_float),
                ("sam_text", C.c_void_p), ("sam_line_offsets", C.c_void_p)]


SPLICE = np.dtype([("pair", "<u4"), ("kind", "<i4"), ("chr", "<i4", (2,)), ("pos", "<u4", (2,)), ("pos_end", "<u4", (2,))])
assert SPLICE.itemsize == 32


def rna_defaults(conf_diff=2):
    """`snap-rna paired` defaults for the three aligners of the RNA loop and the filter (SNAPLib/PairedAligner.cpp:470-527, 582-584)."""
    pp = paired_defaults()
    tp = SingleParams(max_hits=pp.max_hits, max_k=pp.max_k, max_read_size=MAX_READ_LENGTH, num_seeds=pp.num_seeds, seed_coverage=pp.seed_coverage,
                      extra_search_depth=pp.extra_search_depth, explore_popular_seeds=0, stop_on_first_hit=0, max_hits_to_get=1000)
    cp = SingleParams(max_hits=300, max_k=pp.max_k, max_read_size=MAX_READ_LENGTH, num_seeds=12, seed_coverage=pp.seed_coverage,
                      extra_search_depth=pp.extra_search_depth, explore_popular_seeds=0, stop_on_first_hit=0, max_hits_to_get=0)
    fp = FilterParams(pp.max_spacing, pp.force_spacing, conf_diff, pp.max_k, 1000)
    return RnaParams(pp, tp, cp, fp)
