"""AlignmentFilter on the device: the warp-per-pair kernel behind snapb200_filter_paired_batch and the
whole RNA pair loop of a batch (snapb200_rna_batch_*) against the compiled reference's own AlignmentFilter, driven as
SNAPLib/PairedAligner.cpp:575-663 drives it -- records field by field, and the per-pair event records replayed into the reference's
GTFReader so that the nine statistics files it writes are compared byte for byte."""
import ctypes as C
import os

import numpy as np
import pytest

import filter_cases as F
from test_filter_oracle import genome_pieces

pytestmark = pytest.mark.gpu
FIELDS = [f for f in F.FILTER_RESULT.names if f != "pad"]


@pytest.fixture(scope="module")
def ws(tmp_path_factory, ref, cuda):
    """Genome + GTF + both indices built by the reference's command line, loaded by the reference and by the CUDA library."""
    from oracle import oracle as O
    d = str(tmp_path_factory.mktemp("cuda_filter"))
    contigs = F.build_workspace(d, O.REF_BIN)
    gdir, tdir, gtf = os.path.join(d, "gidx"), os.path.join(d, "tidx"), os.path.join(d, "a.gtf")
    hg, ht = cuda.load_index(gdir), cuda.load_index(tdir)
    ann = cuda.annotation_open(hg, ht, gtf)
    rg, rt = ref.load_index(gdir), ref.load_index(tdir)
    yield dict(d=d, contigs=contigs, gdir=gdir, tdir=tdir, gtf=gtf, hg=hg, ht=ht, ann=ann, rg=rg, rt=rt)
    cuda.annotation_close(ann)
    cuda.close_index(hg)
    cuda.close_index(ht)


def assert_same_records(want, got, what):
    bad = [i for i in range(len(want)) if any(not np.array_equal(want[f][i], got[f][i]) for f in FIELDS)]
    assert not bad, (what, len(bad), bad[:10], [(want[i], got[i]) for i in bad[:3]])


def replay_and_compare(ref, cuda, w, sam_reads, events, want_prefix, tag, splices=None):
    """The event records through the reference's own public GTFReader methods on a fresh GTFReader: same statistics files.
    splices = (offsets, records): the novel-splice records replace the UnalignedRead calls."""
    lib = ref.lib
    lib.ref_gtf_load.restype = C.c_void_p
    d = w["d"]
    g2 = C.c_void_p(lib.ref_gtf_load(w["gtf"].encode(), os.path.join(d, tag).encode()))
    t_ids, chr_names = cuda.annotation_names(w["ann"])
    assert chr_names == genome_pieces(w["gdir"])[0]
    arr = lambda names: (C.c_char_p * len(names))(*[x.encode() for x in names])
    evc = np.ascontiguousarray(events)
    if splices is None:
        assert lib.ref_filter_replay_events(w["rg"], w["rt"], g2, sam_reads[0].byref(), sam_reads[1].byref(), C.c_uint(15), C.c_void_p(evc.ctypes.data), arr(t_ids),
                                            arr(chr_names)) == 0
    else:
        off, recs = np.ascontiguousarray(splices[0], np.uint64), np.ascontiguousarray(splices[1])
        assert lib.ref_filter_replay_events2(w["rg"], w["rt"], g2, sam_reads[0].byref(), sam_reads[1].byref(), C.c_uint(15), C.c_void_p(evc.ctypes.data), arr(t_ids),
                                             arr(chr_names), off.ctypes.data_as(C.c_void_p), C.c_void_p(recs.ctypes.data)) == 0
    lib.ref_gtf_finish(g2)
    produced = sorted(x for x in os.listdir(d) if x.startswith(tag))
    assert len(produced) >= 8
    for f in produced:
        assert open(os.path.join(d, f), "rb").read() == open(os.path.join(d, want_prefix + f[len(tag):]), "rb").read(), f


def test_cuda_filter_matches_the_reference_and_the_golden_file(ref, cuda, ws):
    """1500 simulated spliced / chimeric pairs; every input of the filter produced by the CUDA library itself."""
    from snap_rnaseq_b200 import _abi as A
    w = ws
    (b0, b1), sam_reads = F.reads(w["contigs"], w["d"])
    hits, genome_res, pp = F.alignments(cuda, w["hg"], w["ht"], b0, b1)
    ch = [cuda.characterize(w["hg"], A.single_defaults(max_hits=300, num_seeds=12), b) for b in (b0, b1)]
    prm = A.FilterParams(pp.max_spacing, pp.force_spacing, 2, 15, F.MAX_HITS_TO_GET)
    res, ev, needs_host = cuda.filter_paired(w["ann"], prm, np.diff(b0.offsets), np.diff(b1.offsets), hits[0], hits[1], genome_res, ch[0], ch[1])
    assert not needs_host.any()
    want = F.run_reference_filter(ref, w["rg"], w["rt"], w["gtf"], os.path.join(w["d"], "want_a"), sam_reads, hits, genome_res, pp)
    assert_same_records(want, res, "CUDA filter vs reference")
    golden = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filter_cases.npz"))
    assert_same_records(golden["result"], res, "CUDA filter vs golden file")
    assert res["aligned_as_pair"].sum() > 500 and (res["is_transcriptome"] == 1).sum() > 100
    replay_and_compare(ref, cuda, w, sam_reads, ev, "want_a", "replay_a")


def test_cuda_filter_on_fabricated_hits(ref, cuda, ws):
    """Dozens of fabricated hits per end: classes of hundreds of combinations, most of them tied -- the order of the map keys, the
    de-duplication rule and the introsort mirror decide the reported pair (the > 16 combinations with a tie at the minimum path)."""
    from snap_rnaseq_b200 import _abi as A
    w = ws
    (b0, b1), sam_reads = F.reads(w["contigs"], w["d"], n=500, seed=19)
    hits, genome_res, pp = F.alignments(cuda, w["hg"], w["ht"], b0, b1)
    res_in = np.ascontiguousarray(genome_res, A.PAIRED_RESULT).copy()
    chr_names, piece_begin = genome_pieces(w["gdir"])
    _, tpiece_begin = genome_pieces(w["tdir"])
    F.fabricate_hits(hits, res_in, tpiece_begin, piece_begin, chr_names, b0.n)
    ch = [cuda.characterize(w["hg"], A.single_defaults(max_hits=300, num_seeds=12), b) for b in (b0, b1)]
    prm = A.FilterParams(pp.max_spacing, pp.force_spacing, 2, 15, F.MAX_HITS_TO_GET)
    res, ev, needs_host = cuda.filter_paired(w["ann"], prm, np.diff(b0.offsets), np.diff(b1.offsets), hits[0], hits[1], res_in, ch[0], ch[1])
    assert not needs_host.any()
    want = F.run_reference_filter(ref, w["rg"], w["rt"], w["gtf"], os.path.join(w["d"], "want_b"), sam_reads, hits, res_in, pp)
    assert_same_records(want, res, "CUDA filter vs reference, fabricated hits")
    kinds = np.bincount(ev["kind"], minlength=4)
    assert kinds[1] > 20 and kinds[2] > 5 and kinds[3] > 2 and (res["status"] == 2).any()
    replay_and_compare(ref, cuda, w, sam_reads, ev, "want_b", "replay_b")
    # empty batch and argument errors
    e = cuda.filter_paired(w["ann"], prm, np.zeros(0, np.uint32), np.zeros(0, np.uint32), tuple(a[:0] for a in hits[0]), tuple(a[:0] for a in hits[1]),
                           res_in[:0], (np.zeros(1, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint16)),
                           (np.zeros(1, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint16)))
    assert len(e[0]) == 0
    bad = (hits[0][0].copy(), hits[0][1], hits[0][2], hits[0][3])
    bad[0][3] = F.MAX_HITS_TO_GET + 1
    with pytest.raises(RuntimeError, match="hit count"):
        cuda.filter_paired(w["ann"], prm, np.diff(b0.offsets), np.diff(b1.offsets), bad, hits[1], res_in, ch[0], ch[1])


def test_rna_batch_is_the_pair_loop_of_the_reference(ref, cuda, ws):
    """snapb200_rna_batch_*: transcriptome multi-hits, genome pair, CharacterizeSeeds and the filter in one submission with the
    intermediates resident in HBM -- against the reference's aligners + AlignmentFilter with the same per-thread aligner parameters
    (PairedAligner.cpp:470-527), twice through the same batch object (buffers are reused) and with two objects in flight."""
    from snap_rnaseq_b200 import _abi as A
    w = ws
    P = A.rna_defaults()
    objs = [cuda.rna_batch_create(w["ann"], w["hg"], w["ht"]) for _ in range(2)]
    cases = []
    for k, (n, seed) in enumerate(((1500, 8), (700, 23))):
        (b0, b1), sam_reads = F.reads(w["contigs"], w["d"], n=n, seed=seed)
        cases.append((b0, b1, sam_reads))
        cuda.rna_batch_submit(objs[k], P, b0, b1)       # both in flight
    outs = [cuda.rna_batch_wait(objs[k]) for k in range(2)]
    cuda.rna_batch_submit(objs[0], P, cases[1][0], cases[1][1])  # the first object again, with the other batch
    again = cuda.rna_batch_wait(objs[0])
    for f in FIELDS:
        assert np.array_equal(again["results"][f], outs[1]["results"][f]), f
    assert np.array_equal(again["events"], outs[1]["events"])
    for k, (b0, b1, sam_reads) in enumerate(cases):
        o = outs[k]
        assert o["n"] == b0.n and not o["needs_host"].any()
        # the intermediates the view exposes are what the separate entry points return
        hits = []
        for e, b in enumerate((b0, b1)):
            _, cnt, locs, rcs, scores = ref.single_multihit(w["rt"], P.transcriptome, b)
            off, hl, hr, hs = o["hits"][e]
            assert np.array_equal(np.diff(off.astype(np.int64)), cnt)
            m = np.arange(locs.shape[1])[None, :] < cnt[:, None]
            assert np.array_equal(hl, locs[m]) and np.array_equal(hr, rcs[m]) and np.array_equal(hs, scores[m])
            hits.append((np.ascontiguousarray(cnt, np.int32), np.ascontiguousarray(locs, np.uint32), np.ascontiguousarray(rcs, np.uint8),
                         np.ascontiguousarray(scores, np.int32)))
            seg, cl, co = ref.characterize(w["rg"], P.partial, b)
            assert np.array_equal(o["ch"][e][0], seg) and np.array_equal(o["ch"][e][1], cl) and np.array_equal(o["ch"][e][2], co)
        genome_res = ref.paired(w["rg"], P.paired, b0, b1)
        for f in ("location", "score", "mapq", "status", "direction", "p_all", "p_best"):
            assert np.array_equal(o["genome_pairs"][f], genome_res[f]), f
        want = F.run_reference_filter(ref, w["rg"], w["rt"], w["gtf"], os.path.join(w["d"], f"want_r{k}"), sam_reads, hits, genome_res, P.paired)
        assert_same_records(want, o["results"], f"rna batch {k} vs reference")
        replay_and_compare(ref, cuda, w, sam_reads, o["events"], f"want_r{k}", f"replay_r{k}")
        # the novel-splice records of the flagged reads instead of UnalignedRead on the host: same files
        assert not o["splice_overflow"].any() and len(o["splices"]) > 0 and (o["events"]["unaligned"] > 0).sum() > 20
        assert np.all(np.diff(o["splice_offsets"].astype(np.int64))[o["events"]["unaligned"] == 0] == 0)
        replay_and_compare(ref, cuda, w, sam_reads, o["events"], f"want_r{k}", f"replay_s{k}", splices=(o["splice_offsets"], o["splices"]))
    # the SAM lines as the last stage of the same submission: what writePair writes for the filter's results
    b0, b1, sam_reads = cases[0]
    cuda.rna_batch_submit(objs[0], P, b0, b1, sam=(sam_reads[0], sam_reads[1], False, "grp"))
    o = cuda.rna_batch_wait(objs[0])
    assert_same_records(outs[0]["results"], o["results"], "with the SAM stage")
    aln = []
    for e in range(2):
        a = np.zeros(b0.n, A.SAM_ALIGNMENT)
        for f in ("location", "mapq", "status", "direction", "is_transcriptome", "tlocation"):
            a[f] = o["results"][f][:, e]
        aln.append(a)
    lib = ref.lib
    lib.ref_gtf_load.restype = C.c_void_p
    g = C.c_void_p(lib.ref_gtf_load(w["gtf"].encode(), os.path.join(w["d"], "sam_stage").encode()))
    want = ref.sam(w["rg"], sam_reads[0], sam_reads[1], aln[0], aln[1], False, "grp", rna=(g, w["rt"]))[0]
    assert o["sam_text"] == want
    lo = o["sam_line_offsets"].astype(np.int64)
    assert len(lo) == 2 * b0.n + 1 and lo[-1] == len(want) and all(want[lo[k + 1] - 1:lo[k + 1]] == b"\n" for k in range(2 * b0.n))
    # ... and as BAM records
    from test_io_fuzz import assert_same_bam
    cuda.rna_batch_submit(objs[0], P, b0, b1, sam=(sam_reads[0], sam_reads[1], 2, "grp"))
    o = cuda.rna_batch_wait(objs[0])
    want = ref.sam(w["rg"], sam_reads[0], sam_reads[1], aln[0], aln[1], False, "grp", rna=(g, w["rt"]), bam=True)[0]
    assert_same_bam(want, o["sam_text"], "rna batch, BAM records")
    with pytest.raises(RuntimeError, match="clipped_len"):
        bad = A.SamReads(sam_reads[0].offsets, sam_reads[0].bases, sam_reads[0].quals, sam_reads[0].front_clip, sam_reads[0].clipped_len - 1,
                         sam_reads[0].id_offsets, sam_reads[0].ids)
        cuda.rna_batch_submit(objs[0], P, b0, b1, sam=(bad, sam_reads[1], False, None))
    # an empty batch is legal
    cuda.rna_batch_submit(objs[1], P, cases[0][0].slice(0, 0), cases[0][1].slice(0, 0))
    assert cuda.rna_batch_wait(objs[1])["n"] == 0
    for h in objs:
        cuda.rna_batch_destroy(h)


def test_cuda_filter_fuzz_against_the_serial_specification(cuda, ws, ref):
    """The warp schedule of filter_warp.cuh against the serial specification it parallelises (flt_filter_pair of filterfmt.h, run on
    the host by tests/hostsim and verified against the reference by tests/test_filter_oracle.py): random pairs with up to 600 hits
    per end over all transcripts, few distinct scores (ties everywhere), repeated places -- lists far longer than a warp, classes of
    up to tens of thousands of combinations, so every strided loop, the rank sort, the ordered compaction and the introsort path run
    many times.  Pairs either side reports as too large for its scratch are compared only where both decided."""
    import ctypes as C
    import subprocess
    from snap_rnaseq_b200 import _abi as A
    from test_filter_oracle import FLT_EVENT, FLT_RESULT, flat_tables
    w = ws
    lib = ref.lib
    lib.ref_gtf_load.restype = C.c_void_p
    g = C.c_void_p(lib.ref_gtf_load(w["gtf"].encode(), os.path.join(w["d"], "fuzz").encode()))
    assert lib.ref_gtf_export(g, os.path.join(w["d"], "gtf_fuzz.tsv").encode()) == 0
    T, keep = flat_tables(os.path.join(w["d"], "gtf_fuzz.tsv"), w["gdir"], w["tdir"])
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "hostsim", "libiohostsim.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(here, "hostsim", "io_hostsim.cpp")], check=True)
    hs = C.CDLL(so)
    n, mh = 240, F.MAX_HITS_TO_GET
    (b0, b1), _ = F.reads(w["contigs"], w["d"], n=n, seed=31)
    rng = np.random.default_rng(9)
    tb = keep["tpiece_begin"].astype(np.int64)
    tlen = np.diff(np.append(tb, tb[-1] + 3000))
    pb = keep["piece_begin"].astype(np.int64)
    hits = []
    for e in range(2):
        cnt, loc, rcs, sc = np.zeros(n, np.int32), np.zeros((n, mh), np.uint32), np.zeros((n, mh), np.uint8), np.zeros((n, mh), np.int32)
        for i in range(n):
            k = int(rng.choice([0, 1, 3, 20, 70, 200, 600]))
            p = rng.integers(1, len(tb), size=k)
            places = (tb[p] + (rng.random(k) * np.maximum(1, tlen[p] - 250)).astype(np.int64)).astype(np.uint32)
            if k > 4:  # repeated places with other scores / strands: HashAlignment's replace rule
                places[rng.integers(0, k, size=k // 4)] = places[rng.integers(0, k, size=k // 4)]
            cnt[i], loc[i, :k], rcs[i, :k], sc[i, :k] = k, places, rng.integers(0, 2, size=k), rng.integers(0, int(rng.choice([2, 4, 17])), size=k)
        hits.append((cnt, loc, rcs, sc))
    res_in = np.zeros(n, A.PAIRED_RESULT)
    for e in range(2):
        res_in["location"][:, e] = np.where(rng.random(n) < 0.85, pb[rng.integers(1, len(pb), size=n)] + rng.integers(0, 150000, size=n), 0xFFFFFFFF)
        res_in["score"][:, e], res_in["mapq"][:, e] = rng.integers(0, 17, size=n), rng.integers(0, 71, size=n)
        res_in["direction"][:, e], res_in["status"][:, e] = rng.integers(0, 2, size=n), rng.integers(0, 3, size=n)
    ch = [cuda.characterize(w["hg"], A.single_defaults(max_hits=300, num_seeds=12), b) for b in (b0, b1)]
    prm = A.FilterParams(1000, 0, 2, 15, mh)
    res, ev, needs_host = cuda.filter_paired(w["ann"], prm, np.diff(b0.offsets), np.diff(b1.offsets), hits[0], hits[1], res_in, ch[0], ch[1])
    p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint64))
    p16 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint16))
    lens0, lens1 = np.diff(b0.offsets), np.diff(b1.offsets)
    (n0, l0, r0, s0), (n1, l1, r1, s1) = hits
    compared = sorted_path = 0
    for i in range(n):
        want, wev = np.zeros(1, FLT_RESULT), np.zeros(1, FLT_EVENT)
        rc = hs.hostsim_filter_pair_flat(C.byref(T), C.c_uint(int(lens0[i])), C.c_uint(int(lens1[i])), C.c_uint(15), C.c_uint(1000), C.c_uint(2), C.c_int(0),
                                         C.c_int(int(n0[i])), A.p32u(l0[i]), A.p8(r0[i]), A.p32i(s0[i]), C.c_int(int(n1[i])), A.p32u(l1[i]), A.p8(r1[i]),
                                         A.p32i(s1[i]), C.c_void_p(res_in[i:i + 1].ctypes.data), p64(ch[0][0]), A.p32u(ch[0][1]), p16(ch[0][2]), p64(ch[1][0]),
                                         A.p32u(ch[1][1]), p16(ch[1][2]), C.c_uint(i), C.c_uint(1001), C.c_uint(1 << 20), C.c_uint(1 << 16),
                                         C.c_void_p(want.ctypes.data), C.c_void_p(wev.ctypes.data))
        if rc != 0 or needs_host[i]:
            continue
        compared += 1
        sorted_path += int(n0[i]) * int(n1[i]) > 64
        for f in FIELDS:
            assert np.array_equal(want[f][0], res[f][i]), (i, f, int(n0[i]), int(n1[i]), want[0], res[i])
        for f in FLT_EVENT.names:
            assert np.array_equal(wev[f][0], ev[f][i]), (i, f, wev[0], ev[i])
    assert compared > 0.6 * n and sorted_path > 40, (compared, sorted_path, int(needs_host.sum()))


def test_cuda_sam_of_transcriptome_alignments(ref, cuda, ws):
    """snapb200_sam_batch_rna against SimpleReadWriter::writePair / writeRead of the reference with its transcriptome and GTFReader
    (SAM.cpp:1046-1061, LandauVishkin.cpp:119-250): the filter's own results for 1500 simulated spliced / chimeric pairs (several
    hundred transcriptome alignments, most of them across junctions), pairs and single reads, with = / X and with M; then the same
    with soft clips forced on a third of the reads and some transcriptome alignments moved to where no CIGAR can be computed (the
    reference then prints an empty field for a transcriptome alignment, where a genome one gets *)."""
    from snap_rnaseq_b200 import _abi as A
    w = ws
    (b0, b1), sam_reads = F.reads(w["contigs"], w["d"])
    hits, genome_res, pp = F.alignments(cuda, w["hg"], w["ht"], b0, b1)
    ch = [cuda.characterize(w["hg"], A.single_defaults(max_hits=300, num_seeds=12), b) for b in (b0, b1)]
    prm = A.FilterParams(pp.max_spacing, pp.force_spacing, 2, 15, F.MAX_HITS_TO_GET)
    res, _, _ = cuda.filter_paired(w["ann"], prm, np.diff(b0.offsets), np.diff(b1.offsets), hits[0], hits[1], genome_res, ch[0], ch[1])
    aln = []
    for e in range(2):
        a = np.zeros(b0.n, A.SAM_ALIGNMENT)
        for f in ("location", "mapq", "status", "direction", "is_transcriptome", "tlocation"):
            a[f] = res[f][:, e]
        aln.append(a)
    assert (aln[0]["is_transcriptome"] == 1).sum() > 100
    lib = ref.lib
    lib.ref_gtf_load.restype = C.c_void_p
    g = C.c_void_p(lib.ref_gtf_load(w["gtf"].encode(), os.path.join(w["d"], "sam_rna").encode()))
    rng = np.random.default_rng(3)
    for clipped in (False, True):
        if clipped:
            for r in sam_reads:
                lens = np.diff(r.offsets).astype(np.int64)
                pick = rng.random(r.n) < 0.33
                front = np.where(pick, rng.integers(0, 6, size=r.n), 0)
                back = np.where(pick, rng.integers(0, 6, size=r.n), 0)
                r.front_clip[:] = front.astype(np.uint16)
                r.clipped_len[:] = (lens - front - back).astype(np.uint16)
            # transcriptome alignments moved to where the read does not align (no CIGAR within 30 edits) or onto a piece boundary
            _, tpiece_begin = genome_pieces(w["tdir"])
            for a in aln:
                t = np.flatnonzero(a["is_transcriptome"] == 1)
                a["tlocation"][t[:12]] += 41
                a["tlocation"][t[12:16]] = tpiece_begin[2:6] - 20
        for use_m in (False, True):
            want = ref.sam(w["rg"], sam_reads[0], sam_reads[1], aln[0], aln[1], use_m, None, rna=(g, w["rt"]))[0]
            got, lo = cuda.sam(w["hg"], sam_reads[0], sam_reads[1], aln[0], aln[1], use_m, None, rna=(w["ann"], w["ht"]))
            assert bytes(got) == want, next((a, b) for a, b in zip(want.split(b"\n"), bytes(got).split(b"\n")) if a != b)
            assert lo[-1] == len(want)
            n_junction = sum(1 for ln in want.split(b"\n") if ln and b"N" in ln.split(b"\t")[5])
            assert n_junction > 20
            if clipped:
                assert sum(1 for ln in want.split(b"\n") if ln and ln.split(b"\t")[5] == b"") > 0  # the empty field of a failed transcriptome CIGAR
            # the same pairs as BAM records (BAMFormat::writeRead: binary CIGAR operations with the N runs, bin from the reference span)
            from test_io_fuzz import assert_same_bam, bam_records
            want = ref.sam(w["rg"], sam_reads[0], sam_reads[1], aln[0], aln[1], use_m, "grp", rna=(g, w["rt"]), bam=True)[0]
            got, lo = cuda.sam(w["hg"], sam_reads[0], sam_reads[1], aln[0], aln[1], use_m, "grp", rna=(w["ann"], w["ht"]), bam=True)
            assert_same_bam(want, got, f"BAM records, clipped {clipped} use_m {use_m}")
            assert int(lo[-1]) == len(want) and len(bam_records(want)) == 2 * b0.n
        want = ref.sam(w["rg"], sam_reads[1], None, aln[1], None, False, "grp", rna=(g, w["rt"]))[0]
        got, _ = cuda.sam(w["hg"], sam_reads[1], None, aln[1], None, False, "grp", rna=(w["ann"], w["ht"]))
        assert bytes(got) == want
    # the genome-only entry point refuses what it cannot format
    with pytest.raises(RuntimeError, match="snapb200_sam_batch_rna"):
        cuda.sam(w["hg"], sam_reads[0], sam_reads[1], aln[0], aln[1])
    with pytest.raises(RuntimeError, match="index pair"):
        cuda.sam(w["ht"], sam_reads[0], sam_reads[1], aln[0], aln[1], rna=(w["ann"], w["ht"]))
