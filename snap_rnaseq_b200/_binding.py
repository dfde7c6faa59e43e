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


    # -- row f2: FASTQ text -> reads, alignments -> SAM text -------------------------------------------------------
    def fastq_parse(self, text, clipping=0, max_reads=None, bufs=None):
        """FASTQReader::getNextRead + Read::clip over every complete record of `text` -> (SamReads, bytes_consumed).
        The CUDA library and tests/hostsim take the text; the compiled reference (prefix ref_) reads a file through its own
        FASTQReader, so for it `text` is written to a temporary file first."""
        if isinstance(text, np.ndarray):  # used as is (e.g. a view of pinned memory)
            t = np.ascontiguousarray(text, np.uint8)
            nb = t.size
            raw = None
        else:
            raw = bytes(text)
            nb = len(raw)
            t = np.frombuffer(raw, np.uint8) if nb else np.zeros(1, np.uint8)
        if max_reads is None:
            max_reads = nb // 8 + 1
        if bufs is not None:  # preallocated outputs: (offsets, id_offsets, bases, quals, ids, front_clip, clipped_len)
            off, ioff, bases, quals, ids, fc, cl = bufs
            max_reads = min(off.size, ioff.size) - 1
            assert min(bases.size, quals.size, ids.size) >= nb and min(fc.size, cl.size) >= max_reads
        else:
            off = np.zeros(max_reads + 1, np.uint32)
            ioff = np.zeros(max_reads + 1, np.uint32)
            bases = np.zeros(max(nb, 1), np.uint8)
            quals = np.zeros(max(nb, 1), np.uint8)
            ids = np.zeros(max(nb, 1), np.uint8)
            fc = np.zeros(max_reads, np.uint16)
            cl = np.zeros(max_reads, np.uint16)
        n = C.c_uint32(0)
        used = C.c_uint64(0)
        tail = [C.byref(n)]
        if self.prefix == "ref_":
            import os
            import tempfile
            # the reference's reader soft_exits on an incomplete record at the end of a *file*; in a run the next buffer
            # supplies the rest, so the file it is given here ends with the last complete record
            pos = np.flatnonzero(t[:nb] == 10)
            keep = int(pos[len(pos) // 4 * 4 - 1]) + 1 if len(pos) >= 4 else 0
            with tempfile.NamedTemporaryFile(suffix=".fq", delete=False) as f:
                f.write(t[:keep].tobytes())
            try:
                self._check(self.fn("fastq_parse")(f.name.encode(), C.c_int(clipping), C.c_uint32(max_reads), C.byref(n), A.p32u(off), A.p8(bases),
                                                   A.p8(quals), A.p16u(fc), A.p16u(cl), A.p32u(ioff), A.p8(ids)), "fastq_parse")
            finally:
                os.unlink(f.name)
            used = None
        else:
            args = [A.p8(t), C.c_uint64(nb), C.c_int(clipping), C.c_uint32(max_reads), C.byref(n), C.byref(used), A.p32u(off), A.p8(bases),
                    A.p8(quals), A.p16u(fc), A.p16u(cl), A.p32u(ioff), A.p8(ids)]
            if self.device is not None:
                args.insert(0, C.c_int(self.device))
            self._check(self.fn("fastq_parse")(*args), "fastq_parse")
            used = int(used.value)
        k = int(n.value)
        return A.SamReads(off[:k + 1], bases[:off[k]], quals[:off[k]], fc[:k], cl[:k], ioff[:k + 1], ids[:ioff[k]]), used

    def fastq_record_start(self, text):
        """FASTQReader::skipPartialRecord: offset of the first record in a buffer that may begin mid-record (len(text) if none)."""
        raw = bytes(text)
        off = C.c_uint64(0)
        t = np.frombuffer(raw, np.uint8) if raw else np.zeros(1, np.uint8)
        self._check(self.fn("fastq_record_start")(A.p8(t), C.c_uint64(len(raw)), C.byref(off)), "fastq_record_start")
        return int(off.value)

    def sam(self, handle, reads0, reads1, aln0, aln1, use_m=False, read_group=None, out=None, rna=None, bam=False):
        """SimpleReadWriter::writeRead / writePair over SAMFormat::writeRead for a batch -> (SAM bytes, line_offsets).
        out: a uint8 array to write into with ONE call (returns a view of it); otherwise measure first, then write.
        rna = (annotation handle [the reference: its GTFReader], transcriptome index handle): alignments may be transcriptome ones
        (snapb200_sam_batch_rna).  bam: BAM records (SNAPB200_SAM_BAM_RECORDS) instead of SAM lines."""
        use_m = int(bool(use_m)) | (2 if bam else 0)
        a0 = np.ascontiguousarray(aln0, A.SAM_ALIGNMENT)
        a1 = np.ascontiguousarray(aln1, A.SAM_ALIGNMENT) if reads1 is not None else None
        n_lines = reads0.n * (2 if reads1 is not None else 1)
        rg = read_group.encode() if read_group else None
        r1 = reads1.byref() if reads1 is not None else None
        p1 = a1.ctypes.data_as(C.c_void_p) if a1 is not None else None
        if self.prefix == "ref_":
            import os
            import tempfile
            fd, path = tempfile.mkstemp(suffix=".sam")
            os.close(fd)
            try:
                if rna is not None:
                    self._check(self.fn("sam_batch_rna")(handle, rna[1], rna[0], reads0.byref(), r1, a0.ctypes.data_as(C.c_void_p), p1, C.c_int(int(use_m)),
                                                         rg, path.encode()), "sam_batch_rna")
                else:
                    self._check(self.fn("sam_batch")(handle, reads0.byref(), r1, a0.ctypes.data_as(C.c_void_p), p1, C.c_int(int(use_m)), rg,
                                                     path.encode()), "sam_batch")
                with open(path, "rb") as f:
                    return f.read(), None
            finally:
                os.unlink(path)
        lo = np.zeros(n_lines + 1, np.uint64)
        plo = lo.ctypes.data_as(C.POINTER(C.c_uint64))
        f = self.fn("sam_batch")
        if rna is not None:
            g = self.fn("sam_batch_rna")
            f = lambda h, *rest: g(rna[0], h, rna[1], *rest)
        if out is not None:
            self._check(f(handle, reads0.byref(), r1, a0.ctypes.data_as(C.c_void_p), p1, C.c_int(int(use_m)), rg, out.ctypes.data_as(C.c_char_p),
                          C.c_uint64(out.size), plo), "sam_batch")
            return out[:int(lo[-1])], lo
        self._check(f(handle, reads0.byref(), r1, a0.ctypes.data_as(C.c_void_p), p1, C.c_int(int(use_m)), rg, None, C.c_uint64(0), plo), "sam_batch")
        total = int(lo[-1])
        out = np.zeros(max(total, 1), np.uint8)
        self._check(f(handle, reads0.byref(), r1, a0.ctypes.data_as(C.c_void_p), p1, C.c_int(int(use_m)), rg, out.ctypes.data_as(C.c_char_p),
                      C.c_uint64(total), plo), "sam_batch")
        return out[:total].tobytes(), lo

    def bgzf_compress(self, data, chunk=0, device=0):
        """snapb200_bgzf_compress (hostsim: the serial specification of csrc/bgzf.h) -> (BGZF bytes, block offsets or None)."""
        raw = bytes(data)
        a = np.frombuffer(raw, np.uint8) if raw else np.zeros(1, np.uint8)
        c = chunk or 65024
        n_blocks = max(1, (len(raw) + c - 1) // c)
        out = np.zeros(len(raw) + 31 * n_blocks + 64, np.uint8)
        if self.prefix == "hostsim_":
            f = self.lib.hostsim_bgzf_compress
            f.restype = C.c_longlong
            n = f(A.p8(a), C.c_uint64(len(raw)), C.c_uint32(c), A.p8(out), C.c_uint64(out.size))
            if n < 0:
                raise RuntimeError("hostsim_bgzf_compress failed")
            return out[:n].tobytes(), None
        nb = C.c_uint64(0)
        off = np.zeros(n_blocks + 1, np.uint64)
        self._check(self.fn("bgzf_compress")(C.c_int(device), A.p8(a), C.c_uint64(len(raw)), C.c_uint32(chunk), A.p8(out), C.c_uint64(out.size), C.byref(nb),
                                             off.ctypes.data_as(C.POINTER(C.c_uint64))), "bgzf_compress")
        return out[:int(nb.value)].tobytes(), off

    def bgzf_last_kernel_ms(self):
        v = C.c_float(0)
        self.fn("bgzf_last_kernel_ms")(C.byref(v))
        return v.value

    def io_last_kernel_ms(self):
        a, b = C.c_float(0), C.c_float(0)
        self.fn("io_last_kernel_ms")(C.byref(a), C.byref(b))
        return a.value, b.value


def _cstrings(buf, n, stride):
    out = []
    raw = buf.tobytes()
    for i in range(n):
        s = raw[i * stride:(i + 1) * stride]
        out.append(s.split(b"\0", 1)[0].decode())
    return out
