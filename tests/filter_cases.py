"""Inputs for the AlignmentFilter differential tests (the next row): a small genome + GTF +
transcriptome built with the reference's own command line, spliced / chimeric pairs from the repo's simulator, and the alignments the
filter consumes (transcriptome multi-hits of both ends, the genome pair) computed by whichever aligner implementation is handed in.
Shared by tests/golden/make_golden_filter.py and tests/test_filter_oracle.py."""
import os
import subprocess

import numpy as np

from snap_rnaseq_b200 import _abi as A
from snap_rnaseq_b200 import synth

MAX_HITS_TO_GET = 1000  # SNAPLib/PairedAligner.cpp:584
N_PAIRS = 1500


def build_workspace(d, ref_bin):
    """Genome, GTF, genome index and transcriptome index, exactly as tests/test_dropin_sam.py builds them."""
    contigs = {"chrDecoy": synth.random_contigs([2000], seed=99)["chr1"]}
    contigs.update(synth.random_contigs([300000, 200000], seed=20))
    synth.inject_repeats({k: v for k, v in contigs.items() if k != "chrDecoy"}, frac=0.04, seed=21, max_len=800, max_copies=60)
    synth.write_fasta(os.path.join(d, "g.fa"), contigs)
    synth.make_gtf(os.path.join(d, "a.gtf"), contigs)
    for cmd in ([ref_bin, "index", "g.fa", "gidx", "-s", "20", "-t1"], [ref_bin, "transcriptome", "a.gtf", "g.fa", "tidx", "-t1", "-s", "20"]):
        r = subprocess.run(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout[-2000:]
    return contigs


def reads(contigs, d, n=N_PAIRS, seed=8):
    b0, b1 = synth.simulate_rna(contigs, os.path.join(d, "a.gtf"), n, 100, seed=seed)
    out = []
    for e, b in enumerate((b0, b1)):
        ids = [f"x{i:x}/{e + 1}" for i in range(b.n)]
        lens = np.diff(b.offsets).astype(np.uint16)
        idb, idoff = A.strings_to_offsets([s.encode() for s in ids])
        out.append(A.SamReads(b.offsets, b.bases, b.quals, np.zeros(b.n, np.uint16), lens, idoff, idb[:idoff[-1]]))
    return (b0, b1), out


def alignments(impl, h_genome, h_transcriptome, b0, b1):
    """What the run loop feeds the filter (SNAPLib/PairedAligner.cpp:586-617), from any implementation of the aligners."""
    tp = A.single_defaults(max_hits_to_get=MAX_HITS_TO_GET)
    pp = A.paired_defaults()
    hits = []
    for b in (b0, b1):
        _, cnt, locs, rcs, scores = impl.single_multihit(h_transcriptome, tp, b)
        hits.append((np.ascontiguousarray(cnt, np.int32), np.ascontiguousarray(locs, np.uint32), np.ascontiguousarray(rcs, np.uint8),
                     np.ascontiguousarray(scores, np.int32)))
    return hits, impl.paired(h_genome, pp, b0, b1), pp


FILTER_RESULT = np.dtype([("location", "<u4", (2,)), ("tlocation", "<u4", (2,)), ("score", "<i4", (2,)), ("mapq", "<i4", (2,)),
                          ("status", "u1", (2,)), ("direction", "u1", (2,)), ("is_transcriptome", "u1", (2,)), ("aligned_as_pair", "u1"), ("pad", "u1")])


def run_reference_filter(ref, h_genome, h_transcriptome, gtf_path, out_prefix, sam_reads, hits, genome_res, pp, conf_diff=2, max_dist=15):
    """ref_filter_paired_batch + the GTF epilogue; returns the per-pair records (FILTER_RESULT)."""
    import ctypes as C
    lib = ref.lib
    lib.ref_gtf_load.restype = C.c_void_p
    g = C.c_void_p(lib.ref_gtf_load(str(gtf_path).encode(), str(out_prefix).encode()))
    n = sam_reads[0].n
    out = np.zeros(n, FILTER_RESULT)
    res = np.ascontiguousarray(genome_res, A.PAIRED_RESULT)
    (n0, l0, r0, s0), (n1, l1, r1, s1) = hits
    rc = lib.ref_filter_paired_batch(h_genome, h_transcriptome, g, sam_reads[0].byref(), sam_reads[1].byref(), C.c_uint(pp.min_spacing),
                                     C.c_uint(pp.max_spacing), C.c_int(int(pp.force_spacing)), C.c_uint(conf_diff), C.c_uint(max_dist),
                                     C.c_uint(MAX_HITS_TO_GET), A.p32i(n0), A.p32u(l0), A.p8(r0), A.p32i(s0), A.p32i(n1), A.p32u(l1), A.p8(r1),
                                     A.p32i(s1), res.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    lib.ref_gtf_finish(g)
    return out


def fabricate_hits(hits, res, tpiece_begin, piece_begin, chr_names, n, seed=77):
    """Overwrites the filter's inputs of every pair with dozens of fabricated transcriptome hits per end (scores around the maxDist
    gate, repeats of the same place with other scores / strands) and a genome pair anywhere, so that the classes hold hundreds of
    combinations with many equal scores -- the regime where the string order of the map keys and libstdc++'s sort decide which pair
    is reported.  Some pairs keep one or two hits so that every class gets to decide somewhere."""
    rng = np.random.default_rng(seed)
    tb = tpiece_begin.astype(np.int64)
    tlen = np.diff(np.append(tb, tb[-1] + 3000))
    real = np.flatnonzero(np.array([c != "chrDecoy" for c in chr_names]))
    pb = piece_begin.astype(np.int64)
    for i in range(n):
        few = rng.random() < 0.3
        for (cnt, loc, rcs, sc) in hits:
            k = int(rng.integers(1, 3)) if few else int(rng.integers(5, 45))
            p = rng.integers(1, len(tb), size=k)  # not the decoy transcript
            loc[i, :k] = (tb[p] + (rng.random(k) * np.maximum(1, tlen[p] - 300)).astype(np.int64)).astype(np.uint32)
            sc[i, :k] = rng.integers(0, 18, size=k) if rng.random() < 0.5 else rng.integers(0, 3, size=k)
            rcs[i, :k] = rng.integers(0, 2, size=k)
            dup = rng.integers(0, k, size=k // 3)
            loc[i, k:k + len(dup)] = loc[i, dup]
            sc[i, k:k + len(dup)] = np.maximum(0, sc[i, dup] + rng.integers(-1, 2, size=len(dup)))
            rcs[i, k:k + len(dup)] = rng.integers(0, 2, size=len(dup))
            cnt[i] = k + len(dup)
        for e in range(2):
            c = int(rng.choice(real))
            res["location"][i, e] = int(pb[c] + rng.integers(0, 150000)) if rng.random() < 0.85 else 0xFFFFFFFF
            res["score"][i, e] = int(rng.integers(0, 18))
            res["mapq"][i, e] = int(rng.integers(0, 71))
            res["direction"][i, e] = int(rng.integers(0, 2))
            res["status"][i, e] = int(rng.integers(0, 3))
