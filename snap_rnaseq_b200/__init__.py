"""snap_rnaseq_b200 -- B200-native alignment core for SNAP-RNA behind a C ABI (include/snapb200.h).

This Python layer is only a binding: it loads the in-tree ``libsnapb200.so`` (hand-written sm_100a CUDA) and
mirrors the reference's class names for the path (GenomeIndex, BaseAligner, ChimericPairedEndAligner,
LandauVishkin, SAMFormat.computeCigarString) in batch form.  There is no CPU implementation here: if the
library is missing or no CUDA device is present, calls fail.
"""
import ctypes as C
import os

import numpy as np

from . import _abi as A
from ._abi import (Batch, IndexInfo, PairedParams, SingleParams, paired_defaults, single_defaults,  # noqa: F401
                   PAIRED_RESULT, SINGLE_RESULT)
from ._binding import BatchLib

_HERE = os.path.dirname(os.path.abspath(__file__))
# SNAPB200_SO: load another build of the same library (scripts/profrun.py uses the -DSNAPB200_PROFILE build)
SO_PATH = os.environ.get("SNAPB200_SO") or os.path.join(_HERE, "libsnapb200.so")
_lib = None


class Snapb200(BatchLib):
    """The C ABI of libsnapb200.so."""

    def __init__(self, cdll, device=0):
        super().__init__(cdll, "snapb200_", device=device)
        cdll.snapb200_last_error.restype = C.c_char_p

    def load_index(self, d, device=None):
        h = C.c_void_p()
        dev = self.device if device is None else device
        self._check(self.lib.snapb200_index_open(str(d).encode(), C.c_int(dev), C.byref(h)), "index_open")
        return h

    def index_from_memory(self, seed_len, padding, table_sizes, tables, overflow, bases, piece_offsets, device=None):
        h = C.c_void_p()
        dev = self.device if device is None else device
        ts = np.ascontiguousarray(table_sizes, np.uint64)
        tb = np.ascontiguousarray(tables, np.uint32)
        ov = np.ascontiguousarray(overflow, np.uint32)
        bs = np.ascontiguousarray(bases, np.uint8)
        po = np.ascontiguousarray(piece_offsets, np.uint32)
        self._check(self.lib.snapb200_index_from_memory(
            C.c_int(dev), C.c_uint32(seed_len), C.c_uint32(padding), C.c_uint32(len(ts)),
            ts.ctypes.data_as(C.c_void_p), tb.ctypes.data_as(C.c_void_p), ov.ctypes.data_as(C.c_void_p),
            C.c_uint32(ov.size), bs.ctypes.data_as(C.c_void_p), C.c_uint32(bs.size), po.ctypes.data_as(C.c_void_p),
            C.c_uint32(po.size), C.byref(h)), "index_from_memory")
        return h

    def build_index(self, bases, piece_offsets, piece_names=None, seed_len=20, padding=500, slack=0.3, device=None):
        """snapb200_index_build: sort-based index construction on the device (lookup-equivalent to the reference's)."""
        h = C.c_void_p()
        dev = self.device if device is None else device
        bs = np.ascontiguousarray(bases, np.uint8)
        po = np.ascontiguousarray(piece_offsets, np.uint32)
        names = None
        if piece_names is not None:
            names = (C.c_char_p * len(piece_names))(*[n.encode() for n in piece_names])
        self._check(self.lib.snapb200_index_build(
            C.c_int(dev), bs.ctypes.data_as(C.c_void_p), C.c_uint32(bs.size), po.ctypes.data_as(C.c_void_p), names,
            C.c_uint32(po.size), C.c_uint32(seed_len), C.c_uint32(padding), C.c_double(slack), C.byref(h)), "index_build")
        return h

    def save_index(self, h, directory):
        self._check(self.lib.snapb200_index_save(h, str(directory).encode()), "index_save")

    def index_info(self, h):
        info = IndexInfo()
        self._check(self.lib.snapb200_index_info_get(h, C.byref(info)), "index_info_get")
        return info

    def close_index(self, h):
        self.lib.snapb200_index_close.restype = None
        self.lib.snapb200_index_close(h)

    def stats(self, h):
        w = np.zeros(A.STATS_WORDS, np.int64)
        self._check(self.lib.snapb200_stats_get(h, w.ctypes.data_as(C.c_void_p)), "stats_get")
        return w

    def stats_reset(self, h):
        self._check(self.lib.snapb200_stats_reset(h), "stats_reset")

    # -- row f3: AlignmentFilter (first version) --------------------------------------------------------
    def annotation_open(self, genome, transcriptome, gtf_path):
        h = C.c_void_p()
        self._check(self.lib.snapb200_annotation_open(genome, transcriptome, str(gtf_path).encode(), C.byref(h)), "annotation_open")
        return h

    def annotation_close(self, a):
        self.lib.snapb200_annotation_close.restype = None
        self.lib.snapb200_annotation_close(a)

    def filter_paired(self, annotation, params, len0, len1, hits0, hits1, genome_pairs, ch0, ch1):
        """snapb200_filter_paired_batch.  hits = (counts, locations[n, mh], rcs, scores) of snapb200_single_multihit_batch; ch =
        (seg_offsets, locations, seed_offsets) of snapb200_characterize_batch -> (results, events, needs_host)."""
        n = len(len0)
        l0, l1 = np.ascontiguousarray(len0, np.uint32), np.ascontiguousarray(len1, np.uint32)
        g = np.ascontiguousarray(genome_pairs, A.PAIRED_RESULT)
        res, ev, nh = np.zeros(n, A.FILTER_RESULT), np.zeros(n, A.FILTER_EVENT), np.zeros(max(n, 1), np.uint8)
        hs = []
        for (cnt, loc, rcs, sc) in (hits0, hits1):
            hs += [A.p32i(np.ascontiguousarray(cnt, np.int32)), A.p32u(np.ascontiguousarray(loc, np.uint32)), A.p8(np.ascontiguousarray(rcs, np.uint8)),
                   A.p32i(np.ascontiguousarray(sc, np.int32))]
        cs = []
        keep = []
        for (seg, locs, offs) in (ch0, ch1):
            seg, locs, offs = np.ascontiguousarray(seg, np.uint64), np.ascontiguousarray(locs, np.uint32), np.ascontiguousarray(offs, np.uint16)
            keep += [seg, locs, offs]
            cs += [seg.ctypes.data_as(C.POINTER(C.c_uint64)), A.p32u(locs), offs.ctypes.data_as(C.POINTER(C.c_uint16))]
        self._check(self.lib.snapb200_filter_paired_batch(annotation, C.byref(params), C.c_uint32(n), A.p32u(l0), A.p32u(l1), *hs,
                                                          g.ctypes.data_as(C.c_void_p), *cs, res.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p),
                                                          A.p8(nh)), "filter_paired_batch")
        return res, ev, nh[:n]

    def annotation_names(self, a):
        """(transcript ids in the reference's map order, chromosome names) behind the indices of the filter events."""
        self.lib.snapb200_annotation_transcript_id.restype = C.c_char_p
        self.lib.snapb200_annotation_chromosome.restype = C.c_char_p
        self.lib.snapb200_annotation_transcript_count.restype = C.c_uint32
        nt = self.lib.snapb200_annotation_transcript_count(a)
        t = [self.lib.snapb200_annotation_transcript_id(a, C.c_uint32(i)).decode() for i in range(nt)]
        c, i = [], 0
        while True:
            s = self.lib.snapb200_annotation_chromosome(a, C.c_uint32(i))
            if s is None:
                break
            c.append(s.decode())
            i += 1
        return t, c

    # -- the whole RNA pair loop of a batch, intermediates resident in HBM (snapb200_rna_batch_*) ---------------------------------
    def rna_batch_create(self, annotation, genome, transcriptome):
        h = C.c_void_p()
        self._check(self.lib.snapb200_rna_batch_create(annotation, genome, transcriptome, C.byref(h)), "rna_batch_create")
        return h

    def rna_batch_destroy(self, b):
        self.lib.snapb200_rna_batch_destroy.restype = None
        self.lib.snapb200_rna_batch_destroy(b)

    def rna_batch_submit(self, b, params, b0, b1, sam=None):
        """sam = (SamReads of mate 0, of mate 1, use_m [or SNAPB200_SAM_* flags: 2 = BAM records], read group or None): also format the
        SAM lines (snapb200_rna_batch_submit_sam)."""
        if sam is None:
            self._check(self.lib.snapb200_rna_batch_submit(b, C.byref(params), b0.byref(), b1.byref()), "rna_batch_submit")
        else:
            rg = sam[3].encode() if sam[3] else None
            self._check(self.lib.snapb200_rna_batch_submit_sam(b, C.byref(params), b0.byref(), b1.byref(), sam[0].byref(), sam[1].byref(), C.c_int(int(sam[2])), rg),
                        "rna_batch_submit_sam")

    def rna_batch_wait(self, b):
        """-> dict of numpy COPIES of the batch object's pinned outputs (the views die with the next submit)."""
        v = A.RnaView()
        self._check(self.lib.snapb200_rna_batch_wait(b, C.byref(v)), "rna_batch_wait")
        n = v.n

        def arr(ptr, count, dtype):
            if not count or not ptr:
                return np.zeros(0, dtype)
            dt = np.dtype(dtype)
            return np.frombuffer((C.c_uint8 * (count * dt.itemsize)).from_address(ptr), dtype=dt, count=count).copy()

        out = {"n": n, "device_ms": v.device_ms, "results": arr(v.results, n, A.FILTER_RESULT), "events": arr(v.events, n, A.FILTER_EVENT),
               "needs_host": arr(v.needs_host, n, np.uint8), "genome_pairs": arr(v.genome_pairs, n, A.PAIRED_RESULT), "hits": [], "ch": []}
        for e in range(2):
            off = arr(v.hit_offsets[e], n + 1, np.uint32)
            t = int(off[-1]) if n else 0
            out["hits"].append((off, arr(v.hit_locations[e], t, np.uint32), arr(v.hit_rcs[e], t, np.uint8), arr(v.hit_scores[e], t, np.int32)))
            seg = arr(v.seg_offsets[e], 2 * n + 1, np.uint64)
            t = int(seg[-1]) if n else 0
            out["ch"].append((seg, arr(v.ch_locations[e], t, np.uint32), arr(v.ch_seed_offsets[e], t, np.uint16)))
        soff = arr(v.splice_offsets, n + 1, np.uint64)
        out["splice_offsets"], out["splices"] = soff, arr(v.splices, int(soff[-1]) if n else 0, A.SPLICE)
        out["splice_overflow"] = arr(v.splice_overflow, n, np.uint8)
        if v.sam_line_offsets:
            out["sam_line_offsets"] = arr(v.sam_line_offsets, 2 * n + 1, np.uint64)
            out["sam_text"] = arr(v.sam_text, int(out["sam_line_offsets"][-1]), np.uint8).tobytes()
        return out

    def device_count(self):
        return int(self.lib.snapb200_device_count())


class Session:
    """Device-resident batch (snapb200_session_*): upload once, run the kernels, download."""

    def __init__(self, lib, index, max_items, max_read_len=A.MAX_READ_LENGTH):
        self.lib = lib
        self.h = C.c_void_p()
        lib._check(lib.lib.snapb200_session_create(index, C.c_uint32(max_items), C.c_uint32(max_read_len), C.byref(self.h)),
                   "session_create")

    def upload(self, slot, batch):
        self.lib._check(self.lib.lib.snapb200_session_upload(self.h, C.c_int(slot), batch.byref()), "session_upload")

    def upload_raw(self, slot, n, offsets_ptr, bases_ptr, quals_ptr):
        rb = A.ReadBatch(n, C.cast(offsets_ptr, C.POINTER(C.c_uint32)), C.cast(bases_ptr, C.POINTER(C.c_uint8)),
                         C.cast(quals_ptr, C.POINTER(C.c_uint8)))
        self.lib._check(self.lib.lib.snapb200_session_upload(self.h, C.c_int(slot), C.byref(rb)), "session_upload")

    def run_single(self, params):
        self.lib._check(self.lib.lib.snapb200_session_run_single(self.h, C.byref(params)), "session_run_single")

    def run_paired(self, params):
        self.lib._check(self.lib.lib.snapb200_session_run_paired(self.h, C.byref(params)), "session_run_paired")

    def download_single(self, out):
        self.lib._check(self.lib.lib.snapb200_session_download_single(self.h, out.ctypes.data_as(C.c_void_p)), "download_single")
        return out

    def download_paired(self, out):
        self.lib._check(self.lib.lib.snapb200_session_download_paired(self.h, out.ctypes.data_as(C.c_void_p)), "download_paired")
        return out

    def download_paired_ptr(self, ptr):
        self.lib._check(self.lib.lib.snapb200_session_download_paired(self.h, C.c_void_p(ptr)), "download_paired")

    def sync(self):
        self.lib._check(self.lib.lib.snapb200_session_sync(self.h), "session_sync")

    def last_run(self):
        ms, n, tot = C.c_float(), C.c_uint32(), C.c_uint64()
        self.lib._check(self.lib.lib.snapb200_session_last_run(self.h, C.byref(ms), C.byref(n), C.byref(tot)), "last_run")
        return ms.value, n.value, tot.value

    def main_kernel_ms(self):
        ms = C.c_float()
        self.lib._check(self.lib.lib.snapb200_session_main_kernel_ms(self.h, C.byref(ms)), "main_kernel_ms")
        return ms.value

    def close(self):
        if self.h:
            self.lib.lib.snapb200_session_destroy.restype = None
            self.lib.lib.snapb200_session_destroy(self.h)
            self.h = None


def lib(device=0):
    """Load libsnapb200.so.  Raises if it has not been built: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(f"{SO_PATH} is missing -- build it with `python -m snap_rnaseq_b200.build` "
                               "(nvcc, sm_100a).  This package has no CPU fallback.")
        _lib = Snapb200(C.CDLL(SO_PATH), device=device)
    _lib.device = device
    return _lib


# ---- the reference's names for this path, in batch form ---------------------------------------------------------
class GenomeIndex:
    """GenomeIndex::loadFromDirectory (SNAPLib/GenomeIndex.cpp:844-963) -> index + genome resident in HBM."""

    def __init__(self, handle, device):
        self.h, self.device = handle, device

    @classmethod
    def loadFromDirectory(cls, directory, device=0):
        return cls(lib(device).load_index(directory, device), device)

    @classmethod
    def BuildIndex(cls, contigs, seedLen=20, chromosomePadding=500, slack=0.3, device=0):
        """GenomeIndex::BuildIndexToDirectory (SNAPLib/GenomeIndex.cpp:348-720), on the device, from a dict of contigs."""
        from .synth import snap_layout
        bases, offs = snap_layout(contigs, chromosomePadding)
        return cls(lib(device).build_index(bases, offs, list(contigs), seedLen, chromosomePadding, slack, device), device)

    def saveToDirectory(self, directory):
        lib(self.device).save_index(self.h, directory)

    def getSeedLength(self):
        return lib(self.device).index_info(self.h).seed_len

    def getCountOfBases(self):
        return lib(self.device).index_info(self.h).n_bases

    def lookupSeed(self, seeds, max_out=64):
        return lib(self.device).lookup(self.h, seeds, max_out)

    def close(self):
        if self.h:
            lib(self.device).close_index(self.h)
            self.h = None


class BaseAligner:
    """BaseAligner (SNAPLib/BaseAligner.h:41-143): same constructor arguments, AlignRead over a batch."""

    def __init__(self, index, maxHitsToConsider, maxK, maxReadSize, maxSeedsToUse, maxSeedCoverage, extraSearchDepth):
        self.index = index
        self.params = SingleParams(max_hits=maxHitsToConsider, max_k=maxK, max_read_size=maxReadSize,
                                   num_seeds=maxSeedsToUse, seed_coverage=maxSeedCoverage,
                                   extra_search_depth=extraSearchDepth)

    def setExplorePopularSeeds(self, v):
        self.params.explore_popular_seeds = int(bool(v))

    def setStopOnFirstHit(self, v):
        self.params.stop_on_first_hit = int(bool(v))

    def AlignRead(self, batch, maxHitsToGet=0):
        L = lib(self.index.device)
        if maxHitsToGet:
            p = SingleParams.from_buffer_copy(self.params)
            p.max_hits_to_get = maxHitsToGet
            return L.single_multihit(self.index.h, p, batch)
        return L.single(self.index.h, self.params, batch)


class ChimericPairedEndAligner:
    """ChimericPairedEndAligner over IntersectingPairedEndAligner (SNAPLib/ChimericPairedEndAligner.cpp:41-61,
    SNAPLib/IntersectingPairedEndAligner.cpp:34-49): align() over a batch of pairs."""

    def __init__(self, index, maxReadSize, maxHits, maxK, maxSeeds, maxSeedCoverage, minSpacing, maxSpacing, forceSpacing,
                 extraSearchDepth, maxBigHits=16000, maxCandidatePoolSize=1000000):
        self.index = index
        self.params = PairedParams(max_hits=maxHits, max_k=maxK, max_read_size=maxReadSize, num_seeds=maxSeeds,
                                   seed_coverage=maxSeedCoverage, min_spacing=minSpacing, max_spacing=maxSpacing,
                                   force_spacing=int(bool(forceSpacing)), max_big_hits=maxBigHits,
                                   extra_search_depth=extraSearchDepth, max_candidate_pool_size=maxCandidatePoolSize)

    def align(self, batch0, batch1):
        return lib(self.index.device).paired(self.index.h, self.params, batch0, batch1)


class SAMFormat:
    """The aligner call inside SAMFormat::computeCigarString (SNAPLib/SAM.cpp:1159-1189)."""

    @staticmethod
    def computeCigarString(index, batch, locations, directions, useM=False):
        return lib(index.device).cigar(index.h, batch, locations, directions, useM)

    @staticmethod
    def writeReads(index, reads, alignments, useM=False, readGroup=None):
        """SimpleReadWriter::writeRead over SAMFormat::writeRead (SNAPLib/ReadWriter.cpp:90-130, SAM.cpp:977-1153) for a batch of
        single-end reads (_abi.SamReads + SAM_ALIGNMENT records) -> (SAM bytes, line offsets)."""
        return lib(index.device).sam(index.h, reads, None, alignments, None, useM, readGroup)

    @staticmethod
    def writePairs(index, reads0, reads1, alignments0, alignments1, useM=False, readGroup=None):
        """SimpleReadWriter::writePair (SNAPLib/ReadWriter.cpp:132-217): two lines per pair, the end with the lower location first."""
        return lib(index.device).sam(index.h, reads0, reads1, alignments0, alignments1, useM, readGroup)


class FASTQReader:
    """FASTQReader (SNAPLib/FASTQ.cpp) in batch form: every complete record of a text at once."""

    def __init__(self, clipping=0, device=0):
        self.clipping = clipping  # ReadClippingType, SNAPLib/Read.h:85
        self.device = device

    def skipPartialRecord(self, text):
        """Offset of the first record in a buffer that may begin mid-record (FASTQ.cpp:113-184); len(text) if there is none."""
        return lib(self.device).fastq_record_start(text)

    def getReads(self, text):
        """getNextRead until the text is exhausted (FASTQ.cpp:188-246) -> (_abi.SamReads, bytes consumed)."""
        return lib(self.device).fastq_parse(text, self.clipping)

