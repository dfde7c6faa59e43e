"""FASTQ text -> reads and alignments -> SAM text: snapb200_fastq_parse / snapb200_sam_batch
against the reference's own FASTQReader and SimpleReadWriter + SAMFormat.

  golden:    tests/golden/io_cases.npz, written by the compiled reference (tests/golden/make_golden_io.py)
  not gpu:   the compiled reference reproduces the golden file; the per-record logic the kernels compile
             (snap_rnaseq_b200/csrc/iofmt.h) run on the host by tests/hostsim reproduces it too
  gpu:       the CUDA path through the C ABI against the golden file, and against the compiled reference on fresh inputs
             (FASTQ text -> parse -> align -> SAM text on the device, byte for byte)
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import io_cases
from golden.make_golden_io import FASTQ_CASES, SAM_CASES
from snap_rnaseq_b200 import _abi as A
from snap_rnaseq_b200 import synth
from snap_rnaseq_b200._binding import BatchLib
from tests_genome import small_genome

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="session")
def io_golden():
    return np.load(os.path.join(HERE, "golden", "io_cases.npz"))


def golden_reads(g, key):
    return A.SamReads(*[g[f"{key}_{f}"] for f in ("offsets", "bases", "quals", "front_clip", "clipped_len", "id_offsets", "ids")])


# ---- the host simulation of the device logic (test infrastructure) ------------------------------------------------------
class HostsimIndex(C.Structure):
    _fields_ = [("piece_begin", C.POINTER(C.c_uint32)), ("n_pieces", C.c_uint32), ("names_blob", C.c_char_p),
                ("names_off", C.POINTER(C.c_uint32)), ("cigars", C.c_void_p * 2), ("edit_distance", C.c_void_p * 2),
                ("cigar_stride", C.c_uint32)]


@pytest.fixture(scope="session")
def hostsim():
    src = os.path.join(HERE, "hostsim", "io_hostsim.cpp")
    so = os.path.join(HERE, "hostsim", "libiohostsim.so")
    csrc = os.path.join(os.path.dirname(HERE), "snap_rnaseq_b200", "csrc")
    hdrs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(f) for f in [src] + hdrs):
        subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", so, src], check=True)
    return BatchLib(C.CDLL(so), "hostsim_")


def hostsim_index(cigars, eds):
    contigs = small_genome()
    _, piece_off = synth.snap_layout(contigs, 500)
    blob, off = A.strings_to_offsets([k.encode() for k in contigs])
    ix = HostsimIndex()
    keep = [piece_off, blob, off, cigars, eds, bytes(blob[:off[-1]])]
    ix.piece_begin = A.p32u(piece_off)
    ix.n_pieces = len(piece_off)
    ix.names_blob = keep[-1]
    ix.names_off = A.p32u(off)
    for e in range(len(cigars)):
        ix.cigars[e] = cigars[e].ctypes.data
        ix.edit_distance[e] = eds[e].ctypes.data
    ix.cigar_stride = 256
    return ix, keep


def check_fastq_golden(impl, g):
    for seed, n, rlen, clipping in FASTQ_CASES:
        text = io_cases.fastq_text(seed, n, rlen)
        got, used = impl.fastq_parse(text, clipping)
        want = golden_reads(g, f"fq{clipping}")
        assert got.n == n and got.same_as(want), f"clipping {clipping}"
        if used is not None:
            assert text[used:] == b"@partial record\nACGTACGT\n+\n"
        lens = np.diff(got.offsets)
        if clipping in (1, 3):
            assert (got.front_clip[:n] > 0).any()
        if clipping in (2, 3):
            assert (got.front_clip[:n].astype(np.int64) + got.clipped_len[:n] < lens).any()
        if clipping == 0:
            assert (got.front_clip[:n] == 0).all() and (got.clipped_len[:n] == lens).all()


def sam_lines(text, offsets=None):
    return text.split(b"\n")


def assert_same_sam(want, got, what):
    if want == got:
        return
    w, g = want.split(b"\n"), got.split(b"\n")
    for i, (a, b) in enumerate(zip(w, g)):
        if a != b:
            raise AssertionError(f"{what}: line {i} differs\n  expected {a!r}\n  got      {b!r}")
    raise AssertionError(f"{what}: {len(w)} lines expected, {len(g)} produced")


# ---- not gpu -----------------------------------------------------------------------------------------------------------
def test_reference_reproduces_io_golden(ref, io_golden, small_index_dir):
    check_fastq_golden(ref, io_golden)
    h = ref.load_index(small_index_dir)
    for k, (seed, n, rlen, paired, use_m) in enumerate(SAM_CASES):
        reads, aln = io_cases.sam_case(seed, n, rlen, paired)
        sam, _ = ref.sam(h, reads[0], reads[1] if paired else None, aln[0], aln[1] if paired else None, use_m, "grp1" if k == 1 else None)
        assert_same_sam(io_golden[f"sam{k}"].tobytes(), sam, f"case {k}")


def test_hostsim_fastq(hostsim, io_golden):
    check_fastq_golden(hostsim, io_golden)
    # the shapes the reference rejects (FASTQ.cpp:214-223): a blank line, a bad starting character
    for bad in (b"@a\nACGT\n+\nIIII\n\n@b\nACGT\n+\nIIII\n", b"@a\nACGT\n+\nIIII\n@b\nXCGT\n+\nIIII\n", b"a\nACGT\n+\nIIII\n"):
        with pytest.raises(RuntimeError):
            hostsim.fastq_parse(bad, 0)
    r, used = hostsim.fastq_parse(b"", 0)
    assert r.n == 0 and used == 0


def test_hostsim_sam(hostsim, io_golden):
    for k, (seed, n, rlen, paired, use_m) in enumerate(SAM_CASES):
        reads, aln = io_cases.sam_case(seed, n, rlen, paired)
        ends = 2 if paired else 1
        cig = [np.ascontiguousarray(io_golden[f"sam{k}_cigar{e}"]) for e in range(ends)]
        eds = [np.ascontiguousarray(io_golden[f"sam{k}_ed{e}"], np.int32) for e in range(ends)]
        ix, keep = hostsim_index(cig, eds)
        sam, lo = hostsim.sam(C.byref(ix), reads[0], reads[1] if paired else None, aln[0], aln[1] if paired else None, use_m,
                              "grp1" if k == 1 else None)
        want = io_golden[f"sam{k}"].tobytes()
        assert_same_sam(want, sam, f"case {k}")
        assert int(lo[-1]) == len(want) and all(sam[int(o) - 1:int(o)] == b"\n" for o in lo[1:])


def test_record_start_matches_the_reference(cuda, hostsim, ref, tmp_path):
    """FASTQReader::skipPartialRecord: for byte ranges that begin anywhere (inside an id, the bases, the '+' line, a quality
    string that starts with '@' or '+'), the first record the reference's reader stands on is the one the C ABI reports.
    snapb200_fastq_record_start is host logic, so this runs without a GPU."""
    import ctypes
    text = io_cases.fastq_text(5, 400, 100, crlf_frac=0.15, partial_tail=False)
    path = tmp_path / "x.fq"
    path.write_bytes(text)
    ref.lib.ref_fastq_record_start.restype = ctypes.c_int
    rng = np.random.default_rng(3)
    starts = sorted({int(x) for x in rng.integers(1, len(text) - 2000, size=300)} | {1, 2, 5})
    record_starts = {0} | {m + 1 for m in np.flatnonzero(np.frombuffer(text, np.uint8) == 10)}
    n_mid_quality = 0
    for st in starts:
        off = ctypes.c_longlong(0)
        assert ref.lib.ref_fastq_record_start(str(path).encode(), ctypes.c_longlong(st), ctypes.byref(off)) == 0
        want = off.value - st
        got = cuda.fastq_record_start(text[st:])
        assert got == hostsim.fastq_record_start(text[st:])
        assert got == want, f"range starting at {st}: reference skips {want} bytes, C ABI {got}"
        assert st + got in record_starts and text[st + got:st + got + 1] == b"@"
        n_mid_quality += text[st:st + 1] in (b"@", b"+")
    assert n_mid_quality > 0
    # parsing from the reported start gives the tail of the records parsed from the beginning
    st = starts[len(starts) // 2]
    got = hostsim.fastq_record_start(text[st:])
    whole, _ = hostsim.fastq_parse(text, 3)
    part, _ = hostsim.fastq_parse(text[st + got:], 3)
    k = whole.n - part.n
    assert part.n > 0 and np.array_equal(part.clipped_len[:part.n], whole.clipped_len[k:whole.n])
    assert part.ids[:part.id_offsets[-1]].tobytes() == whole.ids[whole.id_offsets[k]:whole.id_offsets[-1]].tobytes()
    # nothing that looks like a record
    assert cuda.fastq_record_start(b"ACGTACGT\nIIIIIIII\n") == len(b"ACGTACGT\nIIIIIIII\n")
    assert cuda.fastq_record_start(b"") == 0


def test_golden_covers_the_branches(io_golden):
    """The golden SAM text must contain what the cases were built to produce (so a silent change of the generator shows)."""
    text = b"".join(io_golden[f"sam{k}"].tobytes() for k in range(len(SAM_CASES)))
    lines = [ln.split(b"\t") for ln in text.split(b"\n") if ln]
    flags = {int(f[1]) for f in lines}
    assert any(f & 0x4 for f in flags) and any(f & 0x8 for f in flags) and any(f & 0x2 for f in flags) and any(f & 0x10 for f in flags)
    assert any(f[5] == b"*" and not int(f[1]) & 0x4 for f in lines)            # mapped but no CIGAR (NM:i:-1)
    assert any(b"S" in f[5] for f in lines) and any(f[6] not in (b"=", b"*") for f in lines)
    assert any(int(f[8]) < 0 for f in lines) and any(int(f[8]) > 0 for f in lines)
    assert any(len(f[9]) < len(f[10]) for f in lines)                           # SEQ cut at the NUL of COMPLEMENT[]
    assert any(f[-3].startswith(b"RG:Z:grp1") for f in lines) and any(b" " not in f[0] for f in lines)
    assert any(f[0].endswith(b"/1") for f in lines) and any(not f[0].endswith((b"/1", b"/2")) for f in lines)


# ---- gpu ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def handle(cuda, small_index_dir):
    h = cuda.load_index(small_index_dir)
    yield h
    cuda.close_index(h)


@pytest.mark.gpu
def test_cuda_fastq_golden(cuda, io_golden):
    check_fastq_golden(cuda, io_golden)
    for bad in (b"@a\nACGT\n+\nIIII\n\n@b\nACGT\n+\nIIII\n", b"@a\nACGT\n+\nIIII\n@b\nXCGT\n+\nIIII\n", b"a\nACGT\n+\nIIII\n"):
        with pytest.raises(RuntimeError):
            cuda.fastq_parse(bad, 0)
    r, used = cuda.fastq_parse(b"", 0)
    assert r.n == 0 and used == 0
    r, used = cuda.fastq_parse(b"@only\nACGT\n+\n", 0)
    assert r.n == 0 and used == 0


@pytest.mark.gpu
def test_cuda_sam_golden(cuda, handle, io_golden):
    for k, (seed, n, rlen, paired, use_m) in enumerate(SAM_CASES):
        reads, aln = io_cases.sam_case(seed, n, rlen, paired)
        sam, lo = cuda.sam(handle, reads[0], reads[1] if paired else None, aln[0], aln[1] if paired else None, use_m,
                           "grp1" if k == 1 else None)
        want = io_golden[f"sam{k}"].tobytes()
        assert_same_sam(want, sam, f"case {k}")
        assert int(lo[-1]) == len(want)
    # skip: nothing is written for the read, the neighbours are unchanged
    reads, aln = io_cases.sam_case(11, 300, 100, False)
    full, lo = cuda.sam(handle, reads[0], None, aln[0], None)
    aln[0]["skip"][::3] = 1
    part, lo2 = cuda.sam(handle, reads[0], None, aln[0], None)
    keep = b"".join(full[int(lo[i]):int(lo[i + 1])] for i in range(300) if i % 3)
    assert part == keep and all(lo2[i] == lo2[i + 1] for i in range(0, 300, 3))
    # empty batch
    e = A.SamReads(np.zeros(1, np.uint32), [], [], [], [], np.zeros(1, np.uint32), [])
    sam, lo = cuda.sam(handle, e, None, np.zeros(0, A.SAM_ALIGNMENT), None)
    assert sam == b"" and list(lo) == [0]


@pytest.mark.gpu
@pytest.mark.parametrize("rlen,clipping,seed", [(100, 3, 41), (150, 2, 42), (250, 0, 43)])
def test_cuda_fastq_to_sam_pipeline(cuda, handle, ref, small_index_dir, rlen, clipping, seed):
    """FASTQ text of both mates -> snapb200_fastq_parse -> snapb200_paired_batch -> snapb200_sam_batch, against the
    reference's FASTQReader -> ChimericPairedEndAligner -> SimpleReadWriter::writePair on the same bytes."""
    hc = ref.load_index(small_index_dir)
    texts = [io_cases.fastq_text(seed + e, 1500, rlen, crlf_frac=0.05 * e, partial_tail=False) for e in range(2)]
    got = [cuda.fastq_parse(t, clipping)[0] for t in texts]
    want = [ref.fastq_parse(t, clipping)[0] for t in texts]
    assert got[0].same_as(want[0]) and got[1].same_as(want[1])
    pp = A.paired_defaults(max_k=15 if rlen < 250 else 20)
    b0, b1 = got[0].clipped_batch(), got[1].clipped_batch()
    res = cuda.paired(handle, pp, b0, b1)
    res_ref = ref.paired(hc, pp, b0, b1)
    aln = []
    for e in range(2):
        a = np.zeros(b0.n, A.SAM_ALIGNMENT)
        for f in ("location", "mapq", "status", "direction"):
            a[f] = res[f][:, e]
            assert np.array_equal(res[f][:, e], res_ref[f][:, e])
        aln.append(a)
    for use_m in (False, True):
        sam, _ = cuda.sam(handle, got[0], got[1], aln[0], aln[1], use_m)
        sam_ref, _ = ref.sam(hc, want[0], want[1], aln[0], aln[1], use_m)
        assert_same_sam(sam_ref, sam, f"use_m={use_m}")
    # the same through the reference-named mirror of the package
    import snap_rnaseq_b200 as S
    gi = S.GenomeIndex(handle, 0)
    assert S.SAMFormat.writePairs(gi, got[0], got[1], aln[0], aln[1], True)[0] == sam
    assert S.FASTQReader(clipping).getReads(texts[0])[0].same_as(got[0])
    # single-end lines of the same reads
    sam, _ = cuda.sam(handle, got[1], None, aln[1], None, False, "rg7")
    sam_ref, _ = ref.sam(hc, want[1], None, aln[1], None, False, "rg7")
    assert_same_sam(sam_ref, sam, "single")


@pytest.mark.gpu
def test_cuda_io_concurrent_callers(cuda, handle, io_golden):
    """The reference runs these per worker thread (-t N); every calling thread has its own device buffers and stream."""
    import threading
    seed, n, rlen, paired, use_m = SAM_CASES[0]
    reads, aln = io_cases.sam_case(seed, n, rlen, paired)
    want_sam = io_golden["sam0"].tobytes()
    text = io_cases.fastq_text(*FASTQ_CASES[3][:3])
    want_fq = golden_reads(io_golden, "fq3")
    errors = []

    def worker(k):
        try:
            for _ in range(5):
                sam, _ = cuda.sam(handle, reads[0], reads[1], aln[0], aln[1], use_m)
                assert sam == want_sam
                got, _ = cuda.fastq_parse(text, 3)
                assert got.same_as(want_fq)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


@pytest.mark.gpu
def test_cuda_io_round_trip_properties(cuda, handle):
    """Size-independent properties on a batch too large to hand-check (200 k reads): the parsed arrays rebuild the FASTQ text
    they came from; every SAM line carries its read (reversed and complemented when flagged 0x10), line offsets tile the output,
    and mates point at each other."""
    contigs = small_genome()
    n, rlen = 100_000, 150
    sim = synth.simulate(contigs, n, rlen, paired=True, err=0.02, seed=77, frag=(250, 450))
    b0, b1 = sim["batches"]
    texts = [synth.fastq_fixed(b0, 0), synth.fastq_fixed(b1, 1)]
    reads = []
    for t, b in zip(texts, (b0, b1)):
        r, used = cuda.fastq_parse(t, 0)
        assert used == t.size and r.n == n
        ids = r.ids[:r.id_offsets[-1]].reshape(n, 11)
        rebuilt = np.concatenate([np.full((n, 1), ord("@"), np.uint8), ids, np.full((n, 1), 10, np.uint8), r.bases[:n * rlen].reshape(n, rlen),
                                  np.frombuffer(b"\n+\n", np.uint8)[None, :].repeat(n, 0), r.quals[:n * rlen].reshape(n, rlen),
                                  np.full((n, 1), 10, np.uint8)], axis=1).reshape(-1)
        assert np.array_equal(rebuilt, t)
        reads.append(r)
    res = cuda.paired(handle, A.paired_defaults(), b0, b1)
    aln = []
    for e in range(2):
        a = np.zeros(n, A.SAM_ALIGNMENT)
        for f in ("location", "mapq", "status", "direction"):
            a[f] = res[f][:, e]
        aln.append(a)
    sam, lo = cuda.sam(handle, reads[0], reads[1], aln[0], aln[1])
    assert len(lo) == 2 * n + 1 and int(lo[-1]) == len(sam) and np.all(np.diff(lo.astype(np.int64)) > 0)
    lines = sam.split(b"\n")
    assert lines[-1] == b"" and len(lines) == 2 * n + 1
    comp = bytes.maketrans(b"ACGTN", b"TGCAN")
    for p in range(0, n, 37):
        f1, f2 = lines[2 * p].split(b"\t"), lines[2 * p + 1].split(b"\t")
        assert f1[0] == f2[0] == b"r%08x" % p
        fl1, fl2 = int(f1[1]), int(f2[1])
        assert fl1 & 0x41 == 0x41 and fl2 & 0x81 == 0x81
        assert bool(fl1 & 0x20) == bool(fl2 & 0x10) or fl2 & 0x4      # mate strand mirrors the mate's own strand when it is mapped
        assert abs(int(f1[8])) == abs(int(f2[8]))
        mates = {b0.read(p), b1.read(p)}
        for f, fl in ((f1, fl1), (f2, fl2)):
            seq, qual = f[9], f[10]
            if fl & 0x10:
                seq, qual = seq.translate(comp)[::-1], qual[::-1]
            assert (seq.decode(), qual.decode()) in mates
            assert f[-1].startswith(b"NM:i:") and f[-2] == b"PG:Z:SNAP"
