"""Thin ctypes call layer over a library exporting the include/snapb200.h batch entry points.

The same layer drives the product (prefix ``snapb200_``) and, from tests only, the two CPU checkers
(``oracle_`` = the C restatement, ``ref_`` = the compiled reference), which export the same batch
signatures minus the device argument (plus a thread count for ``ref_``).
"""
import ctypes as C

import numpy as np

from . import _abi as A


class BatchLib:
    def __init__(self, lib, prefix, *, device=None, threads=None):
        self.lib = lib
        self.prefix = prefix
        self.device = device      # int for the CUDA library, None for CPU checkers
        self.threads = threads    # int for the threaded reference driver, None otherwise

    def fn(self, name):
        f = getattr(self.lib, self.prefix + name)
        f.restype = C.c_int
        return f

    def _check(self, rc, what):
        if rc != 0:
            msg = ""
            if self.prefix == "snapb200_":
                self.lib.snapb200_last_error.restype = C.c_char_p
                msg = (self.lib.snapb200_last_error() or b"").decode()
            raise RuntimeError(f"{self.prefix}{what} failed: rc={rc} {msg}")

    # -- building blocks -------------------------------------------------------------------------
    def lv(self, direction, texts, patterns, quals, ks):
        n = len(texts)
        t, to = A.strings_to_offsets(texts)
        p, po = A.strings_to_offsets(patterns)
        q = None
        if quals is not None:
            q, _ = A.strings_to_offsets(quals)
        k = np.asarray(ks, dtype=np.int32)
        score = np.zeros(n, np.int32)
        prob = np.zeros(n, np.float64)
        indel = np.zeros(n, np.int32)
        args = [C.c_int(direction), C.c_uint32(n), A.p32u(to), A.p8(t), A.p32u(po), A.p8(p),
                A.p8(q) if q is not None else None, A.p32i(k), A.p32i(score), A.pf64(prob), A.p32i(indel)]
        if self.device is not None:
            args.insert(0, C.c_int(self.device))
        self._check(self.fn("lv_batch")(*args), "lv_batch")
        return score, prob, indel

    def lv_cigar(self, texts, patterns, ks, use_m, stride=256):
        n = len(texts)
        t, to = A.strings_to_offsets(texts)
        p, po = A.strings_to_offsets(patterns)
        k = np.asarray(ks, dtype=np.int32)
        out = np.zeros(max(n, 1) * stride, np.uint8)
        ed = np.zeros(n, np.int32)
        args = [C.c_uint32(n), A.p32u(to), A.p8(t), A.p32u(po), A.p8(p), A.p32i(k), C.c_int(int(use_m)),
                out.ctypes.data_as(C.c_char_p), C.c_uint32(stride), A.p32i(ed)]
        if self.device is not None:
            args.insert(0, C.c_int(self.device))
        self._check(self.fn("lv_cigar_batch")(*args), "lv_cigar_batch")
        return _cstrings(out, n, stride), ed

    def mapq(self, p_all, p_best, score, popular):
        n = len(p_all)
        pa = np.ascontiguousarray(p_all, np.float64)
        pb = np.ascontiguousarray(p_best, np.float64)
        sc = np.ascontiguousarray(score, np.int32)
        po = np.ascontiguousarray(popular, np.int32)
        out = np.zeros(n, np.int32)
        args = [C.c_uint32(n), A.pf64(pa), A.pf64(pb), A.p32i(sc), A.p32i(po), A.p32i(out)]
        if self.device is not None:
            args.insert(0, C.c_int(self.device))
        self._check(self.fn("mapq_batch")(*args), "mapq_batch")
        return out

    # -- index-bound calls -------------------------------------------------------------------------
    def lookup(self, handle, seeds, max_out=64):
        n = len(seeds)
        s, _ = A.strings_to_offsets(seeds)
        nh = np.zeros((max(n, 1), 2), np.uint32)
        hits = np.zeros((max(n, 1), 2, max_out), np.uint32)
        self._check(self.fn("lookup_seed_batch")(handle, C.c_uint32(n), A.p8(s), C.c_uint32(max_out), A.p32u(nh),
                                                   A.p32u(hits)), "lookup_seed_batch")
        return nh[:n], hits[:n]

    def single(self, handle, params, batch):
        res = np.zeros(max(batch.n, 1), A.SINGLE_RESULT)
        args = [handle, C.byref(params), batch.byref(), res.ctypes.data_as(C.c_void_p)]
        if self.threads is not None:
            args.append(C.c_int(self.threads))
        self._check(self.fn("single_batch")(*args), "single_batch")
        return res[:batch.n]

    def single_multihit(self, handle, params, batch):
        mh = int(params.max_hits_to_get)
        n = max(batch.n, 1)
        res = np.zeros(n, A.SINGLE_RESULT)
        cnt = np.zeros(n, np.int32)
        locs = np.zeros((n, mh), np.uint32)
        rcs = np.zeros((n, mh), np.uint8)
        scores = np.zeros((n, mh), np.int32)
        args = [handle, C.byref(params), batch.byref(), res.ctypes.data_as(C.c_void_p), A.p32i(cnt), A.p32u(locs),
                A.p8(rcs), A.p32i(scores)]
        if self.threads is not None:
            args.append(C.c_int(self.threads))
        self._check(self.fn("single_multihit_batch")(*args), "single_multihit_batch")
        return res[:batch.n], cnt[:batch.n], locs[:batch.n], rcs[:batch.n], scores[:batch.n]

    def characterize(self, handle, params, batch):
        """BaseAligner::CharacterizeSeeds for a batch -> (seg_offsets[2n+1], locations, seed_offsets); segment 2*i+dir."""
        seg = np.zeros(2 * batch.n + 1, np.uint64)
        f = self.fn("characterize_batch")
        tail = [C.c_int(self.threads)] if self.threads is not None else []
        pseg = seg.ctypes.data_as(C.POINTER(C.c_uint64))
        self._check(f(handle, C.byref(params), batch.byref(), pseg, None, None, C.c_uint64(0), *tail), "characterize_batch")
        total = int(seg[-1])
        locs = np.zeros(max(total, 1), np.uint32)
        offs = np.zeros(max(total, 1), np.uint16)
        self._check(f(handle, C.byref(params), batch.byref(), pseg, A.p32u(locs), offs.ctypes.data_as(C.POINTER(C.c_uint16)),
                      C.c_uint64(total), *tail), "characterize_batch")
        assert int(seg[-1]) == total
        return seg, locs[:total], offs[:total]

    def paired(self, handle, params, b0, b1):
        assert b0.n == b1.n
        res = np.zeros(max(b0.n, 1), A.PAIRED_RESULT)
        args = [handle, C.byref(params), b0.byref(), b1.byref(), res.ctypes.data_as(C.c_void_p)]
        if self.threads is not None:
            args.append(C.c_int(self.threads))
        self._check(self.fn("paired_batch")(*args), "paired_batch")
        return res[:b0.n]

    def cigar(self, handle, batch, locations, directions, use_m, stride=512):
        n = batch.n
        loc = np.ascontiguousarray(locations, np.uint32)
        d = np.ascontiguousarray(directions, np.uint8)
        out = np.zeros(max(n, 1) * stride, np.uint8)
        ed = np.zeros(max(n, 1), np.int32)
        self._check(self.fn("cigar_batch")(handle, batch.byref(), A.p32u(loc), A.p8(d), C.c_int(int(use_m)),
                                             out.ctypes.data_as(C.c_char_p), C.c_uint32(stride), A.p32i(ed)),
                    "cigar_batch")
        return _cstrings(out, n, stride), ed[:n]


def _cstrings(buf, n, stride):
    out = []
    raw = buf.tobytes()
    for i in range(n):
        s = raw[i * stride:(i + 1) * stride]
        out.append(s.split(b"\0", 1)[0].decode())
    return out
