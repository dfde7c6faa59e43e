#!/usr/bin/env python3
"""Generate the committed golden fixtures from the COMPILED REFERENCE (oracle/_ref).

Run in the build container (needs oracle/_ref, i.e. /root/reference):
    python tests/golden/make_golden.py
Writes, next to this file:
    small_index.tar.gz   index directory of a 36 kbp repeat-injected genome built by `snap-rna index -s 20 -t1`
    small_cases.npz      reads + the reference's per-read outputs for single / multi-hit / paired / CIGAR /
                         lookupSeed / LandauVishkin<+1,-1> / computeMAPQ on that index
The GPU box has no /root/reference; tests compare the oracle port and the CUDA path with these files.
"""
import io
import os
import sys
import tarfile
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from snap_rnaseq_b200 import _abi as A  # noqa: E402
from snap_rnaseq_b200 import synth  # noqa: E402


def small_genome():
    contigs = synth.random_contigs([16000, 12000, 8000], seed=20)
    synth.inject_repeats(contigs, frac=0.10, seed=21, min_len=100, max_len=600, max_copies=40)
    return contigs


def lv_tuples(rng, n, maxlen=120):
    """Random (text, pattern, qual, k) tuples: pattern = edited prefix of text."""
    texts, pats, quals, ks = [], [], [], []
    alpha = np.frombuffer(b"ACGT", np.uint8)
    for _ in range(n):
        pl = int(rng.integers(0, maxlen))
        t = alpha[rng.integers(0, 4, size=pl + 31)]
        p = list(t[:pl])
        for _e in range(int(rng.integers(0, 7))):
            if not p:
                break
            pos = int(rng.integers(0, len(p)))
            op = rng.integers(0, 3)
            if op == 0:
                p[pos] = alpha[rng.integers(0, 4)]
            elif op == 1:
                del p[pos]
            else:
                p.insert(pos, alpha[rng.integers(0, 4)])
        p = bytes(bytearray(int(c) for c in p))
        tl = len(p) + 31 if rng.random() < 0.8 else int(rng.integers(0, len(t) + 1))
        texts.append(t[:tl].tobytes())
        pats.append(p)
        quals.append(synth.QUAL_LEVELS[rng.integers(0, 8, size=len(p))].tobytes())
        ks.append(int(rng.integers(0, 31)))
    return texts, pats, quals, ks


def main():
    if not O.have_ref():
        raise SystemExit("oracle/_ref missing: run python oracle/build_ref.py first")
    ref = O.ref()
    contigs = small_genome()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "small.fa")
        synth.write_fasta(fa, contigs)
        idx = os.path.join(tmp, "small_index")
        os.makedirs(idx)
        O.ref_build_index(fa, idx, seed_len=20, threads=1)
        with tarfile.open(os.path.join(HERE, "small_index.tar.gz"), "w:gz") as tf:
            for f in ("Genome", "GenomeIndex", "GenomeIndexHash", "OverflowTable"):
                tf.add(os.path.join(idx, f), arcname="small_index/" + f)
        h = ref.load_index(idx)

        def put(prefix, batch):
            out[prefix + "_bases"] = batch.bases
            out[prefix + "_quals"] = batch.quals
            out[prefix + "_offsets"] = batch.offsets

        # single end, defaults; includes junk reads, Ns, and a few short / empty reads
        sim = synth.simulate(contigs, 1500, 100, err=0.03, seed=11, junk_frac=0.03, n_rate=0.03)
        b = sim["batches"][0]
        seqs = [b.read(i) for i in range(b.n)]
        seqs += [("", ""), ("ACGT", "IIII"), (seqs[0][0][:19], seqs[0][1][:19]), (seqs[1][0][:20], seqs[1][1][:20]),
                 (seqs[2][0][:49], seqs[2][1][:49]), ("N" * 60, "I" * 60), (seqs[3][0][:30] + "N" * 16 + seqs[3][0][46:], seqs[3][1])]
        b = A.Batch.from_strings([s for s, _ in seqs], [q for _, q in seqs])
        put("single", b)
        out["single_res"] = ref.single(h, A.single_defaults(), b)
        pm = A.single_defaults(max_hits_to_get=1000, max_hits=16000, num_seeds=8, max_k=15)
        r, cnt, locs, rcs, scores = ref.single_multihit(h, pm, b)
        out["multihit_res"], out["multihit_cnt"] = r, cnt
        out["multihit_locs"], out["multihit_rcs"], out["multihit_scores"] = locs[:, :64], rcs[:, :64], scores[:, :64]
        assert cnt.max() <= 64, cnt.max()
        # CIGAR at the locations the reference reported
        res = out["single_res"]
        for use_m in (0, 1):
            cg, ed = ref.cigar(h, b, res["location"], res["direction"], use_m)
            out[f"cigar{use_m}_str"] = np.array(cg)
            out[f"cigar{use_m}_ed"] = ed
        # paired, defaults
        sim = synth.simulate(contigs, 1500, 100, paired=True, err=0.03, seed=12, junk_frac=0.05, n_rate=0.02)
        b0, b1 = sim["batches"]
        put("pair0", b0)
        put("pair1", b1)
        out["paired_res"] = ref.paired(h, A.paired_defaults(), b0, b1)
        # long, divergent reads (C5-like): 250 bp at 4 %, -d 20
        sim = synth.simulate(contigs, 400, 250, err=0.04, seed=13)
        b = sim["batches"][0]
        put("long", b)
        out["long_res"] = ref.single(h, A.single_defaults(max_k=20), b)
        # lookupSeed on seeds drawn from the reads (hits) and random seeds (mostly misses)
        rng = np.random.default_rng(14)
        seeds = []
        for i in range(300):
            s, _ = b.read(i)
            o = int(rng.integers(0, len(s) - 20))
            seeds.append(s[o:o + 20].encode())
        seeds += [bytes(bytearray(int(c) for c in synth.BASES[rng.integers(0, 4, size=20)])) for _ in range(100)]
        seeds.append(b"ACGTACGTACGTACGTACGT")  # its own reverse complement
        seeds.append(b"ACGTNCGTACGTACGTACGT")
        nh, hits = ref.lookup(h, seeds, max_out=64)
        out["lookup_seeds"] = np.frombuffer(b"".join(seeds), np.uint8).reshape(-1, 20)
        out["lookup_nhits"], out["lookup_hits"] = nh, hits
        # LandauVishkin<+1>/<-1> tuples
        t, p, q, k = lv_tuples(rng, 3000)
        for d in (1, -1):
            s, pr, ni = ref.lv(d, t, p, q, k)
            out[f"lv{'f' if d == 1 else 'r'}_score"], out[f"lv{'f' if d == 1 else 'r'}_prob"] = s, pr
            out[f"lv{'f' if d == 1 else 'r'}_indel"] = ni
        out["lv_texts"], out["lv_text_off"] = A.strings_to_offsets(t)
        out["lv_pats"], out["lv_pat_off"] = A.strings_to_offsets(p)
        out["lv_quals"], _ = A.strings_to_offsets(q)
        out["lv_k"] = np.array(k, np.int32)
        # computeMAPQ
        pa = rng.random(2000) * 10 ** rng.uniform(-12, 0, 2000)
        pb = pa * np.where(rng.random(2000) < 0.3, 1.0, rng.random(2000))
        sc = rng.integers(0, 20, 2000).astype(np.int32)
        po = rng.integers(0, 30, 2000).astype(np.int32) * (rng.random(2000) < 0.5)
        out["mapq_pall"], out["mapq_pbest"], out["mapq_score"], out["mapq_pop"] = pa, pb, sc, po.astype(np.int32)
        out["mapq_out"] = ref.mapq(pa, pb, sc, po.astype(np.int32))
    np.savez_compressed(os.path.join(HERE, "small_cases.npz"), **out)
    for f in ("small_index.tar.gz", "small_cases.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
