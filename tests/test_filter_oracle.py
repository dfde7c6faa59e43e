"""AlignmentFilter (SURVEY.md section 8 row f3 -- the next row; there is no CUDA version yet): the oracle entry point
ref_filter_paired_batch (oracle/ref_driver.cpp: the reference's own AlignmentFilter driven as PairedAligner.cpp:575-663 drives it)
is pinned by tests/golden/filter_cases.npz, and gives the same records whether its inputs come from the compiled reference's
aligners or from the C restatement's -- so the device version will be checked against a fixed target from its first line."""
import os

import numpy as np
import pytest

import filter_cases as F
from conftest import assert_records_equal

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def golden_filter():
    return np.load(os.path.join(HERE, "golden", "filter_cases.npz"))


def test_reference_filter_reproduces_golden(ref, port, golden_filter, tmp_path):
    from oracle import oracle as O
    d = str(tmp_path)
    contigs = F.build_workspace(d, O.REF_BIN)
    (b0, b1), sam_reads = F.reads(contigs, d)
    hg, ht = ref.load_index(os.path.join(d, "gidx")), ref.load_index(os.path.join(d, "tidx"))
    hits, genome_res, pp = F.alignments(ref, hg, ht, b0, b1)
    assert np.array_equal(hits[0][0], golden_filter["hit_counts0"]) and np.array_equal(hits[1][0], golden_filter["hit_counts1"])
    assert np.array_equal(genome_res["location"], golden_filter["genome_location"])
    out = F.run_reference_filter(ref, hg, ht, os.path.join(d, "a.gtf"), os.path.join(d, "run1"), sam_reads, hits, genome_res, pp)
    assert_records_equal(golden_filter["result"], out, what="AlignmentFilter vs golden")
    for f in sorted(os.listdir(d)):
        if f.startswith("run1"):
            key = "file_" + (f[len("run1"):].strip("._") or "main")
            assert open(os.path.join(d, f), "rb").read() == golden_filter[key].tobytes(), f
    # the filter's inputs from the C restatement of the aligners are the same, hence its outputs
    hp, htp = port.load_index(os.path.join(d, "gidx")), port.load_index(os.path.join(d, "tidx"))
    hits_p, genome_p, _ = F.alignments(port, hp, htp, b0, b1)
    for e in range(2):
        n = hits[e][0]
        assert np.array_equal(n, hits_p[e][0])
        m = np.arange(F.MAX_HITS_TO_GET)[None, :] < n[:, None]
        for a, b in zip(hits[e][1:], hits_p[e][1:]):
            assert np.array_equal(a[m], b[m])
    assert_records_equal(genome_res, genome_p, fields=("location", "score", "mapq", "status", "direction"), what="genome pair, port vs reference")
    # what the cases exercise
    r = golden_filter["result"]
    assert r["is_transcriptome"].sum() > 100 and (r["status"] == 0).sum() > 50
    assert (r["location"] != golden_filter["genome_location"]).any(axis=1).sum() > 100


# ---- first pieces of the device filter (snap_rnaseq_b200/csrc/filterfmt.h) on the host ------------------------------------------
import ctypes as C  # noqa: E402


class FltTables(C.Structure):
    _fields_ = [("piece_begin", C.POINTER(C.c_uint32)), ("n_pieces", C.c_uint32), ("chr_names", C.c_char_p), ("chr_name_off", C.POINTER(C.c_uint32)),
                ("tpiece_begin", C.POINTER(C.c_uint32)), ("n_tpieces", C.c_uint32), ("tpiece_transcript", C.POINTER(C.c_int32)),
                ("t_chr", C.POINTER(C.c_int32)), ("t_gene", C.POINTER(C.c_int32)), ("t_end", C.POINTER(C.c_uint32)),
                ("t_feat_first", C.POINTER(C.c_uint32)), ("f_type", C.POINTER(C.c_uint32)), ("f_start", C.POINTER(C.c_uint32)),
                ("f_end", C.POINTER(C.c_uint32)), ("g_chr", C.POINTER(C.c_int32)), ("g_start", C.POINTER(C.c_uint32)), ("g_end", C.POINTER(C.c_uint32)),
                ("n_genes", C.c_uint32), ("gene_tree_min_stop", C.c_int32)]


def genome_pieces(index_dir):
    """(names, beginning offsets) from the `Genome` file of an index directory (SNAPLib/Genome.cpp:126-158)."""
    with open(os.path.join(index_dir, "Genome"), "rb") as f:
        n_pieces = int(f.readline().split()[1])
        rows = [f.readline().decode().rstrip("\n").split(" ", 1) for _ in range(n_pieces)]
    return [r[1] for r in rows], np.array([int(r[0]) for r in rows], np.uint32)


def flat_tables(export_path, gdir, tdir):
    from snap_rnaseq_b200 import _abi as A
    chr_names, piece_begin = genome_pieces(gdir)
    t_names, tpiece_begin = genome_pieces(tdir)
    transcripts, genes = {}, {}
    for line in open(export_path):
        p = line.rstrip("\n").split("\t")
        if p[0] == "T":
            n = int(p[6])
            feats = [(int(p[7 + 3 * k]), int(p[8 + 3 * k]), int(p[9 + 3 * k])) for k in range(n)]
            transcripts[p[1]] = (p[2], p[3], int(p[4]), int(p[5]), feats)
        else:
            genes[p[1]] = (p[2], int(p[3]), int(p[4]))
    t_ids, g_ids = sorted(transcripts), sorted(genes)
    keep = {}
    keep["piece_begin"], keep["tpiece_begin"] = piece_begin, tpiece_begin
    blob, off = A.strings_to_offsets([c.encode() for c in chr_names])
    keep["chr_blob"], keep["chr_off"] = bytes(blob[:off[-1]]), off
    keep["tpiece_transcript"] = np.array([t_ids.index(n) for n in t_names], np.int32)
    keep["t_chr"] = np.array([chr_names.index(transcripts[t][0]) for t in t_ids], np.int32)
    keep["t_gene"] = np.array([g_ids.index(transcripts[t][1]) for t in t_ids], np.int32)
    keep["t_end"] = np.array([transcripts[t][3] for t in t_ids], np.uint32)
    first = np.zeros(len(t_ids) + 1, np.uint32)
    np.cumsum([len(transcripts[t][4]) for t in t_ids], out=first[1:])
    keep["t_feat_first"] = first
    feats = [f for t in t_ids for f in transcripts[t][4]]
    for k, name in enumerate(("f_type", "f_start", "f_end")):
        keep[name] = np.array([f[k] for f in feats], np.uint32)
    T = FltTables()
    T.piece_begin, T.n_pieces = A.p32u(piece_begin), len(piece_begin)
    T.chr_names, T.chr_name_off = keep["chr_blob"], A.p32u(off)
    T.tpiece_begin, T.n_tpieces = A.p32u(tpiece_begin), len(tpiece_begin)
    T.tpiece_transcript = A.p32i(keep["tpiece_transcript"])
    T.t_chr, T.t_gene, T.t_end = A.p32i(keep["t_chr"]), A.p32i(keep["t_gene"]), A.p32u(keep["t_end"])
    T.t_feat_first = A.p32u(first)
    T.f_type, T.f_start, T.f_end = A.p32u(keep["f_type"]), A.p32u(keep["f_start"]), A.p32u(keep["f_end"])
    keep["g_chr"] = np.array([chr_names.index(genes[g][0]) for g in g_ids], np.int32)
    keep["g_start"] = np.array([genes[g][1] for g in g_ids], np.uint32)
    keep["g_end"] = np.array([genes[g][2] for g in g_ids], np.uint32)
    T.g_chr, T.g_start, T.g_end = A.p32i(keep["g_chr"]), A.p32u(keep["g_start"]), A.p32u(keep["g_end"])
    T.n_genes = len(g_ids)
    T.gene_tree_min_stop = int(keep["g_start"][0]) if 0 < len(g_ids) < 64 else -(1 << 31)  # flt_gene_found: the unsorted single-node tree
    return T, keep


def test_alignment_lists_match_the_reference(ref, tmp_path):
    """AddAlignment + HashAlignment for every pair: the de-duplicated lists of both ends, in the string order of the reference's map
    keys, from the flat-table functions of filterfmt.h against the maps inside the reference's own AlignmentFilter."""
    import subprocess
    from oracle import oracle as O
    from snap_rnaseq_b200 import _abi as A
    d = str(tmp_path)
    contigs = F.build_workspace(d, O.REF_BIN)
    (b0, b1), sam_reads = F.reads(contigs, d, n=600)
    hg, ht = ref.load_index(os.path.join(d, "gidx")), ref.load_index(os.path.join(d, "tidx"))
    hits, genome_res, pp = F.alignments(ref, hg, ht, b0, b1)
    lib = ref.lib
    lib.ref_gtf_load.restype = C.c_void_p
    g = C.c_void_p(lib.ref_gtf_load(os.path.join(d, "a.gtf").encode(), os.path.join(d, "lists").encode()))
    assert lib.ref_gtf_export(g, os.path.join(d, "gtf.tsv").encode()) == 0
    T, keep = flat_tables(os.path.join(d, "gtf.tsv"), os.path.join(d, "gidx"), os.path.join(d, "tidx"))
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "hostsim", "libiohostsim.so")
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(here, "hostsim", "io_hostsim.cpp")], check=True)
    hs = C.CDLL(so)
    cap, mh = 2048, F.MAX_HITS_TO_GET
    (n0, l0, r0, s0), (n1, l1, r1, s1) = hits
    res = np.ascontiguousarray(genome_res, A.PAIRED_RESULT)
    lens0, lens1 = np.diff(b0.offsets), np.diff(b1.offsets)
    # the simulated reads give short lists; pairs 300.. get fabricated hits instead (the filter takes whatever the aligners hand it):
    # dozens of transcriptome locations per end with scores around the maxDist gate, exact repeats with other scores (the tie rules of
    # HashAlignment), reads that run off the end of a transcript, and genome locations anywhere
    rng = np.random.default_rng(12)
    tb = keep["tpiece_begin"].astype(np.int64)
    tlen = np.diff(np.append(tb, tb[-1] + 3000))
    glen = int(keep["piece_begin"][-1]) + 150000
    for i in range(300, b0.n):
        for (cnt, loc, rcs, sc) in hits:
            k = int(rng.integers(1, 60))
            p = rng.integers(0, len(tb), size=k)
            loc[i, :k] = (tb[p] + (rng.random(k) * (tlen[p] - 400)).astype(np.int64)).astype(np.uint32)
            sc[i, :k] = rng.integers(0, 19, size=k)
            rcs[i, :k] = rng.integers(0, 2, size=k)
            dup = rng.integers(0, k, size=k // 3)          # the same place again, with another score / strand
            loc[i, k:k + len(dup)] = loc[i, dup]
            sc[i, k:k + len(dup)] = np.maximum(0, sc[i, dup] + rng.integers(-1, 2, size=len(dup)))
            rcs[i, k:k + len(dup)] = rng.integers(0, 2, size=len(dup))
            cnt[i] = k + len(dup)
        for e in range(2):
            res["location"][i, e] = int(rng.integers(600, glen)) if rng.random() < 0.9 else 0xFFFFFFFF
            res["score"][i, e] = int(rng.integers(0, 19))
            res["mapq"][i, e] = int(rng.integers(0, 71))
            res["direction"][i, e] = int(rng.integers(0, 2))
    n_lists = n_entries = n_transcriptome = 0
    for i in range(b0.n):
        cw, cg = np.zeros(2, np.uint32), np.zeros(2, np.uint32)
        rw, rg = np.zeros((2, cap, 7), np.uint32), np.zeros((2, cap, 7), np.uint32)
        keys = C.create_string_buffer(1 << 16)
        rc = lib.ref_filter_alignments(hg, ht, g, sam_reads[0].byref(), sam_reads[1].byref(), C.c_uint(i), C.c_uint(15), C.c_uint(mh), A.p32i(n0), A.p32u(l0),
                                       A.p8(r0), A.p32i(s0), A.p32i(n1), A.p32u(l1), A.p8(r1), A.p32i(s1), res.ctypes.data_as(C.c_void_p), C.c_uint(cap),
                                       A.p32u(cw), A.p32u(rw), keys, C.c_uint(1 << 16))
        assert rc == 0
        rc = hs.hostsim_filter_alignments(C.byref(T), C.c_uint(int(lens0[i])), C.c_uint(int(lens1[i])), C.c_uint(15), C.c_int(int(n0[i])), A.p32u(l0[i]),
                                          A.p8(r0[i]), A.p32i(s0[i]), C.c_int(int(n1[i])), A.p32u(l1[i]), A.p8(r1[i]), A.p32i(s1[i]),
                                          C.c_void_p(res[i:i + 1].ctypes.data), C.c_uint(cap), A.p32u(cg), A.p32u(rg))
        assert rc == 0
        assert np.array_equal(cw, cg), (i, cw, cg)
        for e in range(2):
            assert np.array_equal(rw[e, :cw[e]], rg[e, :cg[e]]), (i, e, rw[e, :cw[e]], rg[e, :cg[e]])
            n_lists += 1
            n_entries += int(cw[e])
            n_transcriptome += int(rw[e, :cw[e], 6].sum())
    assert n_entries > 4 * n_lists and n_transcriptome > 3000  # lists with many entries, transcriptome alignments among them


FLT_RESULT = np.dtype([("location", "<u4", (2,)), ("tlocation", "<u4", (2,)), ("score", "<i4", (2,)), ("mapq", "<i4", (2,)),
                       ("status", "u1", (2,)), ("direction", "u1", (2,)), ("is_transcriptome", "u1", (2,)), ("aligned_as_pair", "u1"), ("pad", "u1")], align=True)


FLT_EVENT = np.dtype([("kind", "<i4"), ("unaligned", "<i4"), ("transcript", "<i4", (2,)), ("chr", "<i4", (2,)), ("pos_original", "<u4", (2,)),
                      ("pos", "<u4", (2,)), ("pos_end", "<u4", (2,))], align=True)


FLT_SPLICE = np.dtype([("pair", "<u4"), ("kind", "<i4"), ("chr", "<i4", (2,)), ("pos", "<u4", (2,)), ("pos_end", "<u4", (2,))], align=True)


def splice_records(hs, T, events, ch, lens, seed_len=20, seg_cap=1 << 16):
    """AlignmentFilter::UnalignedRead of every flagged read as records (the serial specification in filterfmt.h): offsets[n + 1], records."""
    p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint64))
    p16 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint16))
    hs.hostsim_unaligned_splices.restype = C.c_longlong
    n = len(events)
    counts = np.zeros(n + 1, np.uint64)
    parts = []
    for i in np.flatnonzero(events["unaligned"] > 0):
        e = int(events["unaligned"][i]) - 1
        args = (C.byref(T), C.c_uint(int(lens[e][i])), C.c_uint(seed_len), p64(ch[e][0]), A_p32u(ch[e][1]), p16(ch[e][2]), C.c_uint(int(i)), C.c_uint(seg_cap))
        c = hs.hostsim_unaligned_splices(*args, None)
        assert c >= 0
        if c:
            rec = np.zeros(c, FLT_SPLICE)
            assert hs.hostsim_unaligned_splices(*args, C.c_void_p(rec.ctypes.data)) == c
            parts.append(rec)
        counts[i + 1] = c
    off = np.cumsum(counts).astype(np.uint64)
    return off, (np.concatenate(parts) if parts else np.zeros(0, FLT_SPLICE))


def A_p32u(a):
    from snap_rnaseq_b200 import _abi as A
    return A.p32u(a)


def test_filter_decision_matches_the_reference(ref, golden_filter, tmp_path):
    """The whole per-pair decision (classification of every combination, ProcessPairs, CheckNoRC, FindPartialMatches, forceSpacing,
    the MAPQ halving) from the flat-table functions against the golden records of the reference's AlignmentFilter, 1500 pairs."""
    import subprocess
    from oracle import oracle as O
    from snap_rnaseq_b200 import _abi as A
    d = str(tmp_path)
    contigs = F.build_workspace(d, O.REF_BIN)
    (b0, b1), sam_reads = F.reads(contigs, d)
    hg, ht = ref.load_index(os.path.join(d, "gidx")), ref.load_index(os.path.join(d, "tidx"))
    hits, genome_res, pp = F.alignments(ref, hg, ht, b0, b1)
    lib = ref.lib
    lib.ref_gtf_load.restype = C.c_void_p
    g = C.c_void_p(lib.ref_gtf_load(os.path.join(d, "a.gtf").encode(), os.path.join(d, "dec").encode()))
    assert lib.ref_gtf_export(g, os.path.join(d, "gtf.tsv").encode()) == 0
    T, keep = flat_tables(os.path.join(d, "gtf.tsv"), os.path.join(d, "gidx"), os.path.join(d, "tidx"))
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "hostsim", "libiohostsim.so")
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(here, "hostsim", "io_hostsim.cpp")], check=True)
    hs = C.CDLL(so)
    # the partialAligner's CharacterizeSeeds (PairedAligner.cpp:518-527: maxHits 300, 12 seeds)
    cp = A.single_defaults(max_hits=300, num_seeds=12)
    ch = [ref.characterize(hg, cp, b) for b in (b0, b1)]
    (n0, l0, r0, s0), (n1, l1, r1, s1) = hits
    res = np.ascontiguousarray(genome_res, A.PAIRED_RESULT)
    lens0, lens1 = np.diff(b0.offsets), np.diff(b1.offsets)
    out = np.zeros(b0.n, FLT_RESULT)
    events = np.zeros(b0.n, FLT_EVENT)
    p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint64))
    p16 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint16))
    for i in range(b0.n):
        rc = hs.hostsim_filter_pair(C.byref(T), C.c_uint(int(lens0[i])), C.c_uint(int(lens1[i])), C.c_uint(15), C.c_uint(pp.max_spacing), C.c_uint(2),
                                    C.c_int(int(pp.force_spacing)), C.c_int(int(n0[i])), A.p32u(l0[i]), A.p8(r0[i]), A.p32i(s0[i]), C.c_int(int(n1[i])),
                                    A.p32u(l1[i]), A.p8(r1[i]), A.p32i(s1[i]), C.c_void_p(res[i:i + 1].ctypes.data), p64(ch[0][0]), A.p32u(ch[0][1]),
                                    p16(ch[0][2]), p64(ch[1][0]), A.p32u(ch[1][1]), p16(ch[1][2]), C.c_uint(i), C.c_void_p(out[i:i + 1].ctypes.data),
                                    C.c_void_p(events[i:i + 1].ctypes.data))
        assert rc == 0
    want = golden_filter["result"]
    bad = [i for i in range(b0.n) if any(not np.array_equal(want[f][i], out[f][i]) for f in FLT_RESULT.names)]
    assert not bad, (len(bad), bad[:10], [(want[i], out[i]) for i in bad[:3]])
    # flt_filter_pair, the allocation-free form a kernel runs (fixed scratch, classes counted first, only the deciding class
    # materialised and sorted): same records, same events; and it reports scratch that is too small instead of overrunning it
    out2, events2 = np.zeros(b0.n, FLT_RESULT), np.zeros(b0.n, FLT_EVENT)
    overflow = 0
    for i in range(b0.n):
        for caps in ((4, 4, 64), (2048, 1 << 16, 1 << 16)):
            rc = hs.hostsim_filter_pair_flat(C.byref(T), C.c_uint(int(lens0[i])), C.c_uint(int(lens1[i])), C.c_uint(15), C.c_uint(pp.max_spacing), C.c_uint(2),
                                             C.c_int(int(pp.force_spacing)), C.c_int(int(n0[i])), A.p32u(l0[i]), A.p8(r0[i]), A.p32i(s0[i]), C.c_int(int(n1[i])),
                                             A.p32u(l1[i]), A.p8(r1[i]), A.p32i(s1[i]), C.c_void_p(res[i:i + 1].ctypes.data), p64(ch[0][0]), A.p32u(ch[0][1]),
                                             p16(ch[0][2]), p64(ch[1][0]), A.p32u(ch[1][1]), p16(ch[1][2]), C.c_uint(i), C.c_uint(caps[0]), C.c_uint(caps[1]),
                                             C.c_uint(caps[2]), C.c_void_p(out2[i:i + 1].ctypes.data), C.c_void_p(events2[i:i + 1].ctypes.data))
            if rc == 0:
                break
            overflow += 1
        assert rc == 0
    assert np.array_equal(out2, out) and np.array_equal(events2, events) and overflow > 0
    # The statistics: the per-pair event records, replayed in input order through the reference's own public GTFReader methods on a
    # fresh GTFReader (what the shim does once the decision comes from the device), must leave the files the reference's filter leaves.
    assert (events["kind"] == 1).sum() > 300 and (events["kind"] >= 2).sum() > 5 and (events["unaligned"] > 0).sum() > 20
    g2 = C.c_void_p(lib.ref_gtf_load(os.path.join(d, "a.gtf").encode(), os.path.join(d, "replay").encode()))
    t_ids = [ln.split("\t")[1] for ln in open(os.path.join(d, "gtf.tsv")) if ln.startswith("T")]
    chr_names, _ = genome_pieces(os.path.join(d, "gidx"))
    arr = lambda names: (C.c_char_p * len(names))(*[n.encode() for n in names])
    rc = lib.ref_filter_replay_events(hg, ht, g2, sam_reads[0].byref(), sam_reads[1].byref(), C.c_uint(15), C.c_void_p(events.ctypes.data), arr(t_ids),
                                      arr(chr_names))
    assert rc == 0
    lib.ref_gtf_finish(g2)
    produced = sorted(f for f in os.listdir(d) if f.startswith("replay"))
    assert len(produced) >= 8
    for f in produced:
        key = "file_" + (f[len("replay"):].strip("._") or "main")
        assert open(os.path.join(d, f), "rb").read() == golden_filter[key].tobytes(), f
    # the same with UnalignedRead replaced by its records (what the device emits): the files must not change.  This annotation has 11
    # genes, i.e. the reference's gene interval tree is the single unsorted node whose scan can be skipped (flt_gene_found).
    off, recs = splice_records(hs, T, events, ch, (lens0, lens1))
    g3 = C.c_void_p(lib.ref_gtf_load(os.path.join(d, "a.gtf").encode(), os.path.join(d, "replay2").encode()))
    rc = lib.ref_filter_replay_events2(hg, ht, g3, sam_reads[0].byref(), sam_reads[1].byref(), C.c_uint(15), C.c_void_p(events.ctypes.data), arr(t_ids),
                                       arr(chr_names), off.ctypes.data_as(C.c_void_p), C.c_void_p(recs.ctypes.data))
    assert rc == 0
    lib.ref_gtf_finish(g3)
    for f in sorted(x for x in os.listdir(d) if x.startswith("replay2")):
        key = "file_" + (f[len("replay2"):].strip("._") or "main")
        assert open(os.path.join(d, f), "rb").read() == golden_filter[key].tobytes(), f


def test_unaligned_read_records_on_a_larger_annotation(ref, tmp_path):
    """UnalignedRead as records against the reference's own UnalignedRead on a 6 Mbp genome with > 64 genes (the reference's gene
    interval tree splits and sorts there) and repeat families that give unaligned reads hundreds of partial alignments: the interval
    maps after the replay -- and the files written from them -- must be identical."""
    import subprocess
    from oracle import oracle as O
    from snap_rnaseq_b200 import _abi as A, synth
    d = str(tmp_path)
    contigs = {"chrDecoy": synth.random_contigs([2000], seed=99)["chr1"]}
    contigs.update(synth.random_contigs([3_000_000] * 2, seed=20))
    synth.inject_repeats({k: v for k, v in contigs.items() if k != "chrDecoy"}, frac=0.04, seed=21)
    synth.write_fasta(os.path.join(d, "g.fa"), contigs)
    synth.make_gtf(os.path.join(d, "a.gtf"), contigs)
    for cmd in ([O.REF_BIN, "index", "g.fa", "gidx", "-s", "20", "-t4"], [O.REF_BIN, "transcriptome", "a.gtf", "g.fa", "tidx", "-t4", "-s", "20"]):
        r = subprocess.run(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-2000:]
    (b0, b1), sam_reads = F.reads(contigs, d, n=3000, seed=5)
    hg, ht = ref.load_index(os.path.join(d, "gidx")), ref.load_index(os.path.join(d, "tidx"))
    lib = ref.lib
    lib.ref_gtf_load.restype = C.c_void_p
    g = C.c_void_p(lib.ref_gtf_load(os.path.join(d, "a.gtf").encode(), os.path.join(d, "want").encode()))
    assert lib.ref_gtf_export(g, os.path.join(d, "gtf.tsv").encode()) == 0
    T, keep = flat_tables(os.path.join(d, "gtf.tsv"), os.path.join(d, "gidx"), os.path.join(d, "tidx"))
    assert T.n_genes >= 64
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "hostsim", "libiohostsim.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(here, "hostsim", "io_hostsim.cpp")], check=True)
    hs = C.CDLL(so)
    ch = [ref.characterize(hg, A.single_defaults(max_hits=300, num_seeds=12), b) for b in (b0, b1)]
    lens = (np.diff(b0.offsets), np.diff(b1.offsets))
    # every read gets the search (alternating mates), whatever the aligners would have said about it
    events = np.zeros(b0.n, FLT_EVENT)
    events["unaligned"] = 1 + (np.arange(b0.n) & 1)
    t_ids = [ln.split("\t")[1] for ln in open(os.path.join(d, "gtf.tsv")) if ln.startswith("T")]
    chr_names, _ = genome_pieces(os.path.join(d, "gidx"))
    arr = lambda names: (C.c_char_p * len(names))(*[n.encode() for n in names])
    assert lib.ref_filter_replay_events(hg, ht, g, sam_reads[0].byref(), sam_reads[1].byref(), C.c_uint(15), C.c_void_p(events.ctypes.data), arr(t_ids),
                                        arr(chr_names)) == 0
    want_counts = np.zeros(4, np.uint64)
    lib.ref_gtf_interval_counts(g, want_counts.ctypes.data_as(C.c_void_p))
    lib.ref_gtf_finish(g)
    off, recs = splice_records(hs, T, events, ch, lens)
    g2 = C.c_void_p(lib.ref_gtf_load(os.path.join(d, "a.gtf").encode(), os.path.join(d, "mine").encode()))
    assert lib.ref_filter_replay_events2(hg, ht, g2, sam_reads[0].byref(), sam_reads[1].byref(), C.c_uint(15), C.c_void_p(events.ctypes.data), arr(t_ids),
                                         arr(chr_names), off.ctypes.data_as(C.c_void_p), C.c_void_p(recs.ctypes.data)) == 0
    got_counts = np.zeros(4, np.uint64)
    lib.ref_gtf_interval_counts(g2, got_counts.ctypes.data_as(C.c_void_p))
    assert np.array_equal(want_counts, got_counts), (want_counts, got_counts)
    assert want_counts[1] > 1000 and want_counts[1] == 2 * (recs["kind"] == 2).sum() and want_counts[3] == 2 * (recs["kind"] == 3).sum()
    lib.ref_gtf_finish(g2)
    for f in sorted(x for x in os.listdir(d) if x.startswith("mine")):
        assert open(os.path.join(d, f), "rb").read() == open(os.path.join(d, "want" + f[len("mine"):]), "rb").read(), f


def test_sort_mirror_is_std_sort(tmp_path):
    """ProcessPairs takes pairs[0] and pairs[1] after std::sort on the score alone: the mirror of libstdc++'s introsort in
    filterfmt.h must leave every element where std::sort leaves it, for short and long sequences, with few and many ties, sorted,
    reversed and organ-pipe inputs (the shapes that drive the median-of-three and the depth limit)."""
    import subprocess
    from snap_rnaseq_b200 import _abi as A
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "hostsim", "libiohostsim.so")
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(here, "hostsim", "io_hostsim.cpp")], check=True)
    hs = C.CDLL(so)
    rng = np.random.default_rng(4)
    cases = []
    for n in list(range(1, 40)) + [63, 64, 65, 100, 257, 1000, 4096, 20000]:
        for hi in (1, 2, 3, 10, 31, 1000):
            cases.append(rng.integers(0, hi, size=n).astype(np.uint32))
        cases.append(np.arange(n, dtype=np.uint32))
        cases.append(np.arange(n, dtype=np.uint32)[::-1].copy())
        cases.append(np.minimum(np.arange(n), np.arange(n)[::-1]).astype(np.uint32))
        cases.append((np.arange(n) // 3).astype(np.uint32))
    for sc in cases:
        mine, theirs = np.zeros(sc.size, np.uint32), np.zeros(sc.size, np.uint32)
        hs.hostsim_sort_check(A.p32u(sc), C.c_uint(sc.size), A.p32u(mine), A.p32u(theirs))
        assert np.array_equal(mine, theirs), (sc.size, sc[:20])
        assert np.all(np.diff(sc[mine].astype(np.int64)) >= 0)
        # the depth-limit fallback (heapsort) is not reached by these inputs inside std::sort, so it is compared on its own
        hs.hostsim_sort_check(A.p32u(sc), C.c_uint(sc.size | 0x80000000), A.p32u(mine), A.p32u(theirs))
        assert np.array_equal(mine, theirs), ("heap", sc.size, sc[:20])


def random_gtf(rng, path):
    """Annotations with what the loader's quirks react to: several isoforms per gene sharing exons, lines in any order, single-exon
    transcripts as the first or a later line of their gene, non-exon features, comments, GFF3-style Parent, attributes in any order,
    doubled spaces and a single quote inside an attribute (the reference splits fields at tabs AND single quotes)."""
    lines = ["# a comment line"]
    for gi in range(int(rng.integers(3, 12))):
        chrom = f"chr{int(rng.integers(1, 4))}"
        gene = f"G{gi:03d}" if rng.random() < 0.9 else f"G{gi:03d}x"
        base = int(rng.integers(1000, 100000))
        exons = []
        p = base
        for _ in range(int(rng.integers(1, 7))):
            ln = int(rng.integers(50, 400))
            exons.append((p, p + ln - 1))
            p += ln + int(rng.integers(50, 900))
        glines = []
        for ti in range(int(rng.integers(1, 4))):
            tid = f"{gene}.t{ti}"
            k = int(rng.integers(1, len(exons) + 1))
            chosen = sorted(rng.choice(len(exons), size=k, replace=False))
            for e in chosen:
                s, en = exons[e]
                if rng.random() < 0.15:
                    en += int(rng.integers(1, 30))       # an isoform-specific exon end: another feature key
                style = rng.random()
                if style < 0.6:
                    attr = f'gene_id "{gene}"; transcript_id "{tid}"; gene_name "{gene}n"; transcript_name "{tid}n";'
                elif style < 0.8:
                    attr = f'transcript_id "{tid}"; gene_id "{gene}";'
                elif style < 0.9:
                    attr = f'Parent={gene};transcript_id={tid};'
                else:
                    attr = f'gene_id "{gene}";'                  # no transcript id: the gene id stands in
                glines.append("\t".join([chrom, "src", "exon", str(s), str(en), ".", "+-"[int(rng.integers(0, 2))], ".", attr]))
            if rng.random() < 0.3:
                glines.append("\t".join([chrom, "src", "CDS", str(exons[0][0]), str(exons[0][1]), ".", "+", "0", f'gene_id "{gene}"; transcript_id "{tid}";']))
        if rng.random() < 0.5:
            rng.shuffle(glines)
        lines += glines
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def test_annotation_tables_match_the_reference(ref, tmp_path):
    """gtf_tables.h (the loader the device filter will use) against the reference's GTFReader, through the same text export: the
    annotation of the test workspace and 40 random ones."""
    import subprocess
    from oracle import oracle as O
    d = str(tmp_path)
    F.build_workspace(d, O.REF_BIN)
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "hostsim", "libiohostsim.so")
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(here, "hostsim", "io_hostsim.cpp")], check=True)
    hs = C.CDLL(so)
    lib = ref.lib
    lib.ref_gtf_load.restype = C.c_void_p
    rng = np.random.default_rng(21)
    paths = [os.path.join(d, "a.gtf")]
    for k in range(40):
        paths.append(os.path.join(d, f"r{k}.gtf"))
        random_gtf(rng, paths[-1])
    n_unprocessed = 0
    for p in paths:
        g = C.c_void_p(lib.ref_gtf_load(p.encode(), (p + ".out").encode()))
        assert lib.ref_gtf_export(g, (p + ".ref.tsv").encode()) == 0
        assert hs.hostsim_gtf_export(p.encode(), (p + ".mine.tsv").encode()) == 0
        want, got = open(p + ".ref.tsv").read(), open(p + ".mine.tsv").read()
        assert want == got, (p, [(a, b) for a, b in zip(want.split("\n"), got.split("\n")) if a != b][:3])
        n_unprocessed += sum(1 for ln in want.split("\n") if ln.startswith("T") and ln.split("\t")[6] == "0")
    assert n_unprocessed > 10  # transcripts the reference never processes (empty exon lists) occur and are reproduced


def test_filter_header_compiles_for_the_device(tmp_path):
    """filterfmt.h is host/device code: a one-thread-per-pair kernel around flt_filter_pair must cross-compile for sm_100a without a
    single host-only call (the kernel itself is the next round's work; this keeps the header device-clean until then)."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    root = os.path.dirname(HERE)
    src = tmp_path / "flt_check.cu"
    src.write_text('#include "%s"\n'
                   'struct KArgs { FltTables t; FltParams p; const FltPairInput *in; const FltScratch *sc; FltResult *out; FltEvent *ev; int *rc; unsigned n; };\n'
                   '__global__ void flt_kernel(KArgs a) {\n'
                   '    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;\n'
                   '    if (i < a.n) a.rc[i] = flt_filter_pair(a.t, a.p, a.in[i], a.sc[i], &a.out[i], &a.ev[i]);\n'
                   '}\n' % os.path.join(root, "snap_rnaseq_b200", "csrc", "filterfmt.h"))
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-Werror", "all-warnings", "-c", str(src), "-o",
                        str(tmp_path / "flt_check.o")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]


def test_filter_decision_on_fabricated_hits(ref, tmp_path):
    """The simulated reads give the filter short lists.  Here every pair gets dozens of fabricated transcriptome hits per end (scores
    around the maxDist gate, repeats of the same place, both strands) and a genome pair anywhere, so that the classes hold hundreds
    of combinations with many equal scores -- the regime where the string order of the maps and libstdc++'s sort decide which pair is
    reported.  flt_filter_pair against the reference's AlignmentFilter on the same inputs, records and statistics files."""
    import subprocess
    from oracle import oracle as O
    from snap_rnaseq_b200 import _abi as A
    d = str(tmp_path)
    contigs = F.build_workspace(d, O.REF_BIN)
    (b0, b1), sam_reads = F.reads(contigs, d, n=500, seed=19)
    hg, ht = ref.load_index(os.path.join(d, "gidx")), ref.load_index(os.path.join(d, "tidx"))
    hits, genome_res, pp = F.alignments(ref, hg, ht, b0, b1)
    res = np.ascontiguousarray(genome_res, A.PAIRED_RESULT).copy()
    lib = ref.lib
    lib.ref_gtf_load.restype = C.c_void_p
    g = C.c_void_p(lib.ref_gtf_load(os.path.join(d, "a.gtf").encode(), os.path.join(d, "want").encode()))
    assert lib.ref_gtf_export(g, os.path.join(d, "gtf.tsv").encode()) == 0
    T, keep = flat_tables(os.path.join(d, "gtf.tsv"), os.path.join(d, "gidx"), os.path.join(d, "tidx"))
    F.fabricate_hits(hits, res, keep["tpiece_begin"], keep["piece_begin"], genome_pieces(os.path.join(d, "gidx"))[0], b0.n)
    want = F.run_reference_filter(ref, hg, ht, os.path.join(d, "a.gtf"), os.path.join(d, "want"), sam_reads, hits, res, pp)
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "hostsim", "libiohostsim.so")
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(here, "hostsim", "io_hostsim.cpp")], check=True)
    hs = C.CDLL(so)
    ch = [ref.characterize(hg, A.single_defaults(max_hits=300, num_seeds=12), b) for b in (b0, b1)]
    (n0, l0, r0, s0), (n1, l1, r1, s1) = hits
    lens0, lens1 = np.diff(b0.offsets), np.diff(b1.offsets)
    out, events = np.zeros(b0.n, FLT_RESULT), np.zeros(b0.n, FLT_EVENT)
    p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint64))
    p16 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint16))
    for i in range(b0.n):
        rc = hs.hostsim_filter_pair_flat(C.byref(T), C.c_uint(int(lens0[i])), C.c_uint(int(lens1[i])), C.c_uint(15), C.c_uint(pp.max_spacing), C.c_uint(2),
                                         C.c_int(int(pp.force_spacing)), C.c_int(int(n0[i])), A.p32u(l0[i]), A.p8(r0[i]), A.p32i(s0[i]), C.c_int(int(n1[i])),
                                         A.p32u(l1[i]), A.p8(r1[i]), A.p32i(s1[i]), C.c_void_p(res[i:i + 1].ctypes.data), p64(ch[0][0]), A.p32u(ch[0][1]),
                                         p16(ch[0][2]), p64(ch[1][0]), A.p32u(ch[1][1]), p16(ch[1][2]), C.c_uint(i), C.c_uint(2048), C.c_uint(1 << 16),
                                         C.c_uint(1 << 16), C.c_void_p(out[i:i + 1].ctypes.data), C.c_void_p(events[i:i + 1].ctypes.data))
        assert rc == 0
    bad = [i for i in range(b0.n) if any(not np.array_equal(want[f][i], out[f][i]) for f in FLT_RESULT.names)]
    assert not bad, (len(bad), bad[:10], [(want[i], out[i]) for i in bad[:3]])
    kinds = np.bincount(events["kind"], minlength=4)
    assert kinds[1] > 20 and kinds[2] > 5 and kinds[3] > 2 and (out["status"] == 2).any() and (out["is_transcriptome"] == 1).sum() > 200
    g2 = C.c_void_p(lib.ref_gtf_load(os.path.join(d, "a.gtf").encode(), os.path.join(d, "replay").encode()))
    t_ids = [ln.split("\t")[1] for ln in open(os.path.join(d, "gtf.tsv")) if ln.startswith("T")]
    chr_names, _ = genome_pieces(os.path.join(d, "gidx"))
    arr = lambda names: (C.c_char_p * len(names))(*[n.encode() for n in names])
    assert lib.ref_filter_replay_events(hg, ht, g2, sam_reads[0].byref(), sam_reads[1].byref(), C.c_uint(15), C.c_void_p(events.ctypes.data), arr(t_ids),
                                        arr(chr_names)) == 0
    lib.ref_gtf_finish(g2)
    for f in sorted(x for x in os.listdir(d) if x.startswith("replay")):
        assert open(os.path.join(d, f), "rb").read() == open(os.path.join(d, "want" + f[len("replay"):]), "rb").read(), f




def feature_tables(export_path):
    """FltTables with only what sam_splice_cigar reads: the exon / intron list of every transcript, in transcript-id order."""
    from snap_rnaseq_b200 import _abi as A
    transcripts = {}
    for line in open(export_path):
        p = line.rstrip("\n").split("\t")
        if p[0] == "T":
            transcripts[p[1]] = [(int(p[7 + 3 * k]), int(p[8 + 3 * k]), int(p[9 + 3 * k])) for k in range(int(p[6]))]
    t_ids = sorted(transcripts)
    keep = {"first": np.zeros(len(t_ids) + 1, np.uint32)}
    np.cumsum([len(transcripts[t]) for t in t_ids], out=keep["first"][1:])
    feats = [f for t in t_ids for f in transcripts[t]] or [(0, 0, 0)]
    for k, name in enumerate(("f_type", "f_start", "f_end")):
        keep[name] = np.array([f[k] for f in feats], np.uint32)
    T = FltTables()
    T.t_feat_first = A.p32u(keep["first"])
    T.f_type, T.f_start, T.f_end = A.p32u(keep["f_type"]), A.p32u(keep["f_start"]), A.p32u(keep["f_end"])
    return T, keep, t_ids, transcripts


def test_spliced_cigar_matches_the_reference(ref, tmp_path):
    """sam_splice_cigar (iofmt.h; the leader lane of sam_measure_kernel runs it for transcriptome alignments) against
    LandauVishkinWithCigar::insertSpliceJunctions over GTFTranscript::Junctions (SNAPLib/LandauVishkin.cpp:119-250,
    SNAPLib/GTFReader.cpp:1109-1139): random run lists (= X M I D, soft clips) at random positions of the transcripts of random
    annotations, plus one with abutting and overlapping exons (introns of length 0 and below) and one-base exons; reads that start on
    an exon's first base, end on its last base (the reference then ends the CIGAR with an N run) or run past the transcript."""
    import subprocess
    d = str(tmp_path)
    here = os.path.dirname(os.path.abspath(__file__))
    so = os.path.join(here, "hostsim", "libiohostsim.so")
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", so, os.path.join(here, "hostsim", "io_hostsim.cpp")], check=True)
    hs = C.CDLL(so)
    lib = ref.lib
    lib.ref_gtf_load.restype = C.c_void_p
    rng = np.random.default_rng(5)
    paths = []
    for k in range(12):
        paths.append(os.path.join(d, f"r{k}.gtf"))
        random_gtf(rng, paths[-1])
    odd = os.path.join(d, "odd.gtf")
    with open(odd, "w") as f:
        rows = [(100, 149), (150, 199), (200, 200), (201, 260), (250, 300), (400, 400), (402, 450), (1000, 1100)]  # abutting, one base, overlapping
        for gi in range(2):  # the first line of a gene does not register its transcript: give every gene a throw-away first line
            f.write(f'chr1\ts\texon\t10\t20\t.\t+\t.\tgene_id "O{gi}"; transcript_id "O{gi}.first";\n')
            for (s, e) in rows[gi:]:
                f.write(f'chr1\ts\texon\t{s}\t{e}\t.\t+\t.\tgene_id "O{gi}"; transcript_id "O{gi}.t";\n')
    paths.append(odd)
    n_cases = n_spliced = n_tail_n = n_silent = 0
    for p in paths:
        g = C.c_void_p(lib.ref_gtf_load(p.encode(), (p + ".out").encode()))
        assert lib.ref_gtf_export(g, (p + ".tsv").encode()) == 0
        T, keep, t_ids, transcripts = feature_tables(p + ".tsv")
        for tr, tid in enumerate(t_ids):
            feats = transcripts[tid]
            exon_len = [e - s + 1 for (ty, s, e) in feats if ty == 1]
            total = sum(exon_len)
            ends = np.cumsum(exon_len) if exon_len else np.array([50])
            for _ in range(25):
                # a read of 30..150 transcript bases; start / end snapped to exon boundaries now and then
                span = int(rng.integers(30, 151))
                pos = int(rng.integers(1, max(2, total + 20)))
                r = rng.random()
                if r < 0.2:
                    pos = int(rng.choice(ends)) + 1            # the first base of an exon
                elif r < 0.4:
                    pos = max(1, int(rng.choice(ends)) - span + 1)   # ends on the last base of an exon (without I / D)
                runs, left, with_indels = [], span, rng.random() < 0.5
                style_m = rng.random() < 0.3
                while left > 0:
                    n = int(min(left, rng.integers(1, 60)))
                    op = "M" if style_m else "=X"[int(rng.integers(0, 2))]
                    if with_indels and rng.random() < 0.25:
                        op = "ID"[int(rng.integers(0, 2))]
                        n = int(rng.integers(1, 4))
                    if runs and runs[-1][1] == op:
                        runs[-1] = (runs[-1][0] + n, op)
                    else:
                        runs.append((n, op))
                    if op != "I":
                        left -= n
                cb, ca = (int(rng.integers(0, 12)) if rng.random() < 0.3 else 0 for _ in range(2))
                lv = "".join(f"{n}{op}" for n, op in runs).encode()
                tokens = ([cb, ord("S")] if cb else []) + [x for n, op in runs for x in (n, ord(op))] + ([ca, ord("S")] if ca else [])
                tok = np.array(tokens, np.uint32)
                want = C.create_string_buffer(4096)
                rc = lib.ref_splice_cigar(g, tok.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_uint(len(tok)), tid.encode(), C.c_uint(pos), want, C.c_int(4096))
                assert rc >= 0
                got = C.create_string_buffer(4096)
                calls = C.c_uint(0)
                n = hs.hostsim_splice_cigar(C.byref(T), C.c_int(tr), C.c_uint(pos), lv, C.c_uint(len(lv)), C.c_uint(cb), C.c_uint(ca), got, C.c_uint(4095), C.byref(calls))
                assert n >= 0 and got.raw[:n] == want.value, (p, tid, pos, lv, cb, ca, want.value, got.raw[:max(n, 0)])
                assert calls.value == rc  # the reference's operator count, which includes runs that print nothing (BAM's n_cigar_op)
                n_silent += rc != sum(ch.isalpha() or ch == "=" for ch in want.value.decode())
                n_cases += 1
                n_spliced += b"N" in want.value
                n_tail_n += want.value.endswith(b"N")
                # a slot that is one character too small is reported, never overrun
                if n > 0:
                    small = C.create_string_buffer(b"\xee" * (n + 8))
                    assert hs.hostsim_splice_cigar(C.byref(T), C.c_int(tr), C.c_uint(pos), lv, C.c_uint(len(lv)), C.c_uint(cb), C.c_uint(ca), small, C.c_uint(n - 1), None) == -1
                    assert small.raw[n - 1:n + 8] == b"\xee" * 9
    assert n_cases > 500 and n_spliced > 100 and n_tail_n > 5 and n_silent > 0, (n_cases, n_spliced, n_tail_n, n_silent)
