"""Randomised differential test of the per-record logic the FASTQ / SAM kernels compile (snap_rnaseq_b200/csrc/iofmt.h, run on the
host by tests/hostsim) against the compiled reference's FASTQReader and SimpleReadWriter + SAMFormat.  No GPU.  Shapes the fixed
cases of tests/io_cases.py do not reach: mixed LF / CR LF inside one record, quality strings longer or shorter than the bases,
'@' '+' and spaces anywhere in ids and qualities, IUPAC and lower-case letters inside reads, reads of 1..500 bases, arbitrary
clipping, alignments at the first and last base of a contig, in the padding between contigs, mates on other contigs, equal locations,
MAPQ and status values of every kind.  FUZZ_ROUNDS scales it (default sized for the CPU suite)."""
import ctypes as C
import os

import numpy as np
import pytest

from snap_rnaseq_b200 import _abi as A
from snap_rnaseq_b200 import synth
from test_io_edges import assert_same_sam, hostsim, hostsim_index  # noqa: F401  (hostsim is a fixture)
from tests_genome import small_genome

ROUNDS = int(os.environ.get("FUZZ_ROUNDS", "1"))
PRINTABLE = np.arange(ord("!"), ord("~") + 1, dtype=np.uint8)


def random_fastq(rng, n):
    out = []
    for i in range(n):
        L = int(rng.choice([1, 2, 49, 50, 51, 100, 151, 500, int(rng.integers(1, 501))]))
        alphabet = b"ACGTN" if rng.random() < 0.7 else b"ACGTNacgtnRYKMSWrykm."
        seq = bytes(rng.choice(np.frombuffer(alphabet, np.uint8), size=L))
        seq = bytes([rng.choice(np.frombuffer(b"ACGTNacgtn", np.uint8))]) + seq[1:]
        qL = L if (rng.random() < 0.85 or i == n - 1) else max(1, L + int(rng.integers(-5, 6)))
        q = rng.choice(PRINTABLE, size=qL)
        for _ in range(int(rng.integers(0, 3))):
            a = int(rng.integers(0, qL))
            q[a:a + int(rng.integers(1, 60))] = ord("#")
        if rng.random() < 0.3:
            q[:int(rng.integers(1, 70))] = ord("#")
        if rng.random() < 0.3:
            q[qL - int(rng.integers(1, 70)):] = ord("#")
        idl = int(rng.integers(0, 40))
        ident = bytes(rng.choice(np.frombuffer(b"abcXYZ019_:/ @+#", np.uint8), size=idl))
        plus = b"+" + (ident if rng.random() < 0.3 else b"")
        nl = [b"\r\n" if rng.random() < 0.2 else b"\n" for _ in range(4)]
        out.append(b"@" + ident + nl[0] + seq + nl[1] + plus + nl[2] + bytes(q) + nl[3])
    return b"".join(out)


@pytest.mark.parametrize("seed", range(4 * ROUNDS))
def test_fuzz_fastq_parse(hostsim, ref, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(12):
        text = random_fastq(rng, int(rng.integers(1, 40)))
        clipping = int(rng.integers(0, 4))
        want, _ = ref.fastq_parse(text, clipping)
        got, used = hostsim.fastq_parse(text, clipping)
        assert used == len(text)
        assert got.same_as(want), (seed, clipping, text[:300])
        st = int(rng.integers(0, max(1, len(text) - 1)))
        k = hostsim.fastq_record_start(text[st:])
        assert k == len(text) - st or text[st + k:st + k + 1] == b"@"


def random_sam_case(rng, n, paired, piece_off, piece_len):
    total_end = int(piece_off[-1]) + int(piece_len[-1])
    ends = 2 if paired else 1
    reads, alns = [], []
    for e in range(ends):
        seqs, quals, ids, fcs, cls = [], [], [], [], []
        for i in range(n):
            L = int(rng.choice([1, 30, 50, 100, 250, 500, int(rng.integers(1, 501))]))
            alphabet = b"ACGTN" if rng.random() < 0.8 else b"ACGTNRY"
            seqs.append(bytes(rng.choice(np.frombuffer(alphabet, np.uint8), size=L)))
            quals.append(bytes(rng.choice(PRINTABLE, size=L)))
            fc = int(rng.integers(0, L)) if rng.random() < 0.3 else 0
            cl = int(rng.integers(1, L - fc + 1)) if rng.random() < 0.3 else L - fc
            fcs.append(fc)
            cls.append(cl)
            base = bytes(rng.choice(np.frombuffer(b"abc01 _/", np.uint8), size=int(rng.integers(1, 20))))
            ids.append(base + rng.choice([b"", b"/1", b"/2", b"/3"]).item())
        reads.append(A.SamReads.from_lists(ids, seqs, quals, fcs, cls))
        a = np.zeros(n, A.SAM_ALIGNMENT)
        p = rng.integers(0, len(piece_off), size=n)
        kind = rng.random(n)
        loc = piece_off[p].astype(np.int64) + (rng.random(n) * piece_len[p]).astype(np.int64)
        loc = np.where(kind < 0.1, piece_off[p], loc)                                  # first base of a contig
        loc = np.where((kind >= 0.1) & (kind < 0.2), piece_off[p] + piece_len[p] - 1, loc)   # last base
        loc = np.where((kind >= 0.2) & (kind < 0.3), piece_off[p] + piece_len[p] + rng.integers(0, 400, size=n), loc)  # padding after it
        loc = np.where((kind >= 0.3) & (kind < 0.33), total_end + rng.integers(0, 499, size=n), loc)  # the padding after the last contig
        a["location"] = np.where(kind > 0.93, A.INVALID_LOCATION, loc).astype(np.uint32)
        a["status"] = rng.choice([A.NOT_FOUND, A.SINGLE_HIT, A.MULTIPLE_HITS], size=n, p=[0.15, 0.6, 0.25])
        a["direction"] = rng.integers(0, 2, size=n)
        a["mapq"] = rng.choice([-1000, -1, 0, 1, 35, 69, 70, 71, 255, 100000], size=n)
        alns.append(a)
    if paired:  # mates at the same place, and pairs that share an id exactly
        same = rng.random(n) < 0.1
        alns[1]["location"][same] = alns[0]["location"][same]
    return reads, alns


@pytest.mark.parametrize("seed", range(3 * ROUNDS))
def test_fuzz_sam_text(hostsim, ref, small_index_dir, seed):
    import io_cases
    rng = np.random.default_rng(2000 + seed)
    contigs = small_genome()
    _, piece_off = synth.snap_layout(contigs, 500)
    piece_len = np.array([len(v) for v in contigs.values()], np.int64)
    h = ref.load_index(small_index_dir)
    for use_m in (False, True):
        for paired in (False, True):
            n = int(rng.integers(1, 120))
            reads, alns = random_sam_case(rng, n, paired, piece_off, piece_len)
            rg = None if rng.random() < 0.5 else "grp x"
            want, _ = ref.sam(h, reads[0], reads[1] if paired else None, alns[0], alns[1] if paired else None, use_m, rg)
            cig, eds = [], []
            for e in range(len(reads)):
                b, loc, d = io_cases.clipped_for_cigar(reads[e], alns[e])
                cg, ed = ref.cigar(h, b, loc, d, use_m, stride=256)
                cig.append(np.array([c.encode() for c in cg], dtype="S256"))
                eds.append(np.ascontiguousarray(ed, np.int32))
            ix, keep = hostsim_index(cig, eds)
            got, lo = hostsim.sam(C.byref(ix), reads[0], reads[1] if paired else None, alns[0], alns[1] if paired else None, use_m, rg)
            assert_same_sam(want, got, f"seed {seed} use_m {use_m} paired {paired}")


# ---- the same random cases through the CUDA library (the warp-parallel parts: newline scan, QNAME / NUL scans, CIGAR walk) -------
@pytest.fixture(scope="module")
def cuda_handle(cuda, small_index_dir):
    h = cuda.load_index(small_index_dir)
    yield h
    cuda.close_index(h)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_fuzz_fastq_parse_cuda(cuda, ref, seed):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(12):
        text = random_fastq(rng, int(rng.integers(1, 40)))
        clipping = int(rng.integers(0, 4))
        want, _ = ref.fastq_parse(text, clipping)
        got, used = cuda.fastq_parse(text, clipping)
        assert used == len(text)
        assert got.same_as(want), (seed, clipping, text[:300])
        rng.integers(0, max(1, len(text) - 1))  # keeps the stream of random numbers aligned with the host test


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_fuzz_sam_text_cuda(cuda, cuda_handle, ref, small_index_dir, seed):
    rng = np.random.default_rng(2000 + seed)
    contigs = small_genome()
    _, piece_off = synth.snap_layout(contigs, 500)
    piece_len = np.array([len(v) for v in contigs.values()], np.int64)
    h = ref.load_index(small_index_dir)
    for use_m in (False, True):
        for paired in (False, True):
            n = int(rng.integers(1, 120))
            reads, alns = random_sam_case(rng, n, paired, piece_off, piece_len)
            rg = None if rng.random() < 0.5 else "grp x"
            want, _ = ref.sam(h, reads[0], reads[1] if paired else None, alns[0], alns[1] if paired else None, use_m, rg)
            got, lo = cuda.sam(cuda_handle, reads[0], reads[1] if paired else None, alns[0], alns[1] if paired else None, use_m, rg)
            assert_same_sam(want, got, f"seed {seed} use_m {use_m} paired {paired}")


# ---- BAM records (SNAPB200_SAM_BAM_RECORDS): BAMFormat::writeRead's bytes, uncompressed ------------------------------------------
def bam_records(raw, mask_undefined=True):
    """Splits a stream of BAM records.  NM of a read without a location, or whose location has no reference text under it, is an
    uninitialised variable in the reference (SNAPLib/Bam.cpp:644, 808-825: whatever the previous call left on the stack); those four
    bytes are zeroed on both sides (records without CIGAR operations)."""
    import struct
    raw = bytes(raw)
    out, p = [], 0
    while p < len(raw):
        size, = struct.unpack_from("<i", raw, p)
        rec = bytearray(raw[p:p + 4 + size])
        flag, = struct.unpack_from("<H", rec, 18)
        n_ops, = struct.unpack_from("<H", rec, 16)
        if mask_undefined and ((flag & 4) or n_ops == 0):
            rec[-4:] = b"\0\0\0\0"
        out.append(bytes(rec))
        p += 4 + size
    assert p == len(raw)
    return out


def assert_same_bam(want, got, what):
    a, b = bam_records(want), bam_records(got)
    assert len(a) == len(b), (what, len(a), len(b))
    for k, (x, y) in enumerate(zip(a, b)):
        assert x == y, (what, k, x[:36].hex(), y[:36].hex(), x[36:], y[36:])


@pytest.mark.parametrize("seed", range(3 * ROUNDS))
def test_fuzz_bam_records(hostsim, ref, small_index_dir, seed):
    """bam_* of iofmt.h (host simulation) against BAMFormat::writeRead through the reference's own writer: random reads, clipping,
    locations, strands, read groups, ids with spaces (kept in BAM) and /1 /2 suffixes; pairs and single reads; = / X and M."""
    import io_cases
    rng = np.random.default_rng(5000 + seed)
    contigs = small_genome()
    _, piece_off = synth.snap_layout(contigs, 500)
    piece_len = np.array([len(v) for v in contigs.values()], np.int64)
    h = ref.load_index(small_index_dir)
    for use_m in (False, True):
        for paired in (False, True):
            n = int(rng.integers(1, 120))
            reads, alns = random_sam_case(rng, n, paired, piece_off, piece_len)
            rg = None if rng.random() < 0.5 else "grp x"
            want, _ = ref.sam(h, reads[0], reads[1] if paired else None, alns[0], alns[1] if paired else None, use_m, rg, bam=True)
            cig, eds = [], []
            for e in range(len(reads)):
                b, loc, d = io_cases.clipped_for_cigar(reads[e], alns[e])
                cg, ed = ref.cigar(h, b, loc, d, use_m, stride=256)
                cig.append(np.array([c.encode() for c in cg], dtype="S256"))
                eds.append(np.ascontiguousarray(ed, np.int32))
            ix, keep = hostsim_index(cig, eds)
            got, lo = hostsim.sam(C.byref(ix), reads[0], reads[1] if paired else None, alns[0], alns[1] if paired else None, use_m, rg, bam=True)
            assert_same_bam(want, got, f"seed {seed} use_m {use_m} paired {paired}")
            assert int(lo[-1]) == len(want)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_fuzz_bam_records_cuda(cuda, cuda_handle, ref, small_index_dir, seed):
    rng = np.random.default_rng(6000 + seed)
    contigs = small_genome()
    _, piece_off = synth.snap_layout(contigs, 500)
    piece_len = np.array([len(v) for v in contigs.values()], np.int64)
    h = ref.load_index(small_index_dir)
    for use_m in (False, True):
        for paired in (False, True):
            n = int(rng.integers(1, 400))
            reads, alns = random_sam_case(rng, n, paired, piece_off, piece_len)
            rg = None if rng.random() < 0.5 else "grp x"
            want, _ = ref.sam(h, reads[0], reads[1] if paired else None, alns[0], alns[1] if paired else None, use_m, rg, bam=True)
            got, lo = cuda.sam(cuda_handle, reads[0], reads[1] if paired else None, alns[0], alns[1] if paired else None, use_m, rg, bam=True)
            assert_same_bam(want, got, f"seed {seed} use_m {use_m} paired {paired}")
            assert int(lo[-1]) == len(want)


@pytest.mark.gpu
def test_bam_record_name_limit_cuda(cuda, cuda_handle):
    """BAM's l_read_name is one byte: the reference exits for a name of more than 254 bytes (Bam.cpp:723); the library reports it."""
    ok = A.SamReads.from_lists([b"n" * 254], [b"ACGT" * 10], [b"I" * 40])
    bad = A.SamReads.from_lists([b"n" * 255], [b"ACGT" * 10], [b"I" * 40])
    a = np.zeros(1, A.SAM_ALIGNMENT)
    a["location"] = A.INVALID_LOCATION
    got, lo = cuda.sam(cuda_handle, ok, None, a, None, False, None, bam=True)
    assert len(got) == 36 + 255 + 20 + 40 + 8 + 7 and got[12] == 255
    with pytest.raises(RuntimeError, match="254"):
        cuda.sam(cuda_handle, bad, None, a, None, False, None, bam=True)
    assert len(cuda.sam(cuda_handle, bad, None, a, None, False, None)[0]) > 255  # SAM has no such limit
