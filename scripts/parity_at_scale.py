"""Parity at BASELINE.json's full sizes: the CUDA path against the COMPILED REFERENCE (oracle/_ref, all host threads) on the C3
index (3.1 Gbp, built on the device, saved in the reference's file format and loaded by the reference's own loader).
Prints one JSON line: reads compared and % bit-exact for paired (location, strand, edit distance, MAPQ, status per read),
single-end, the multi-hit form, CharacterizeSeeds tuples and CIGAR strings.  TEST INFRASTRUCTURE (uses oracle/).
usage: parity_at_scale.py [pairs] [c3|c2|c5]      (c5: the stress shape -- 1 Gbp genome, 250 bp reads at 4 % error, -d 20)"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import snap_rnaseq_b200 as S
from oracle import oracle as O
from snap_rnaseq_b200 import synth, _abi as A
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
cfg = sys.argv[2] if len(sys.argv) > 2 else "c3"
max_k = 15
if cfg == "c3":
    bench.GENOME_CONTIGS, bench.READ_LEN, bench.ERR_RATE = [25_000_000] * 124, 150, 0.01
elif cfg == "c5":
    bench.GENOME_CONTIGS, bench.READ_LEN, bench.ERR_RATE = [25_000_000] * 40, 250, 0.04
    max_k = 20
L = S.lib(0)
contigs = bench.make_genome()
bases, offs = synth.snap_layout(contigs, 500)
h = L.build_index(bases, offs, list(contigs), seed_len=20)
ref = O.ref(threads=os.cpu_count() or 8)
out = {"config": cfg, "genome_mbp": sum(bench.GENOME_CONTIGS) // 1_000_000, "read_len": bench.READ_LEN, "checker": "compiled reference, %d threads" % (os.cpu_count() or 8)}
with tempfile.TemporaryDirectory(dir=bench.scratch_dir(sum(bench.GENOME_CONTIGS) * 18)) as tmp:
    d = os.path.join(tmp, "idx")
    L.save_index(h, d)
    hc = ref.load_index(d)
    pp = A.paired_defaults(max_k=max_k)
    n_same = n_tot = 0
    t_ref = t_gpu = 0.0
    chunk = 1_000_000
    for k, lo in enumerate(range(0, pairs, chunk)):
        n = min(chunk, pairs - lo)
        b0, b1 = bench.make_pairs(contigs, n, 5000 + k)
        t = time.perf_counter(); got = L.paired(h, pp, b0, b1); t_gpu += time.perf_counter() - t
        t = time.perf_counter(); want = ref.paired(hc, pp, b0, b1); t_ref += time.perf_counter() - t
        same = np.ones((n, 2), bool)
        for f in ("location", "mapq", "status", "score", "direction"):
            same &= want[f] == got[f]
        same &= (want["aligned_as_pair"] == got["aligned_as_pair"])[:, None]
        n_same += int(same.sum()); n_tot += 2 * n
    out["paired"] = {"reads": n_tot, "bit_exact_pct": 100.0 * n_same / n_tot, "gpu_reads_per_s_e2e": n_tot / t_gpu, "reference_reads_per_s": n_tot / t_ref}
    m = min(200_000, b0.n)
    s0 = b0.slice(0, m)
    ps = A.single_defaults(max_k=min(max_k, 20))
    g, w = L.single(h, ps, s0), ref.single(hc, ps, s0)
    ok = np.ones(m, bool)
    for f in ("location", "mapq", "status", "score", "direction", "n_scored", "n_lookups"):
        ok &= g[f] == w[f]
    ok &= (g["p_all"] == w["p_all"]) & (g["p_best"] == w["p_best"])  # the FP64 probabilities computeMAPQ consumed, bit for bit
    out["single"] = {"reads": m, "bit_exact_pct": 100.0 * ok.mean()}
    pm = A.single_defaults(max_hits_to_get=1000, max_hits=16000, num_seeds=8, max_k=15)
    mm = min(50_000, m)
    s1 = b1.slice(0, mm)
    gm, wm = L.single_multihit(h, pm, s1), ref.single_multihit(hc, pm, s1)
    okm = np.array_equal(gm[1], wm[1])
    for i in range(mm):
        c = int(wm[1][i])
        okm = okm and np.array_equal(gm[2][i, :c], wm[2][i, :c]) and np.array_equal(gm[3][i, :c], wm[3][i, :c]) and np.array_equal(gm[4][i, :c], wm[4][i, :c])
    out["multihit"] = {"reads": mm, "hits": int(wm[1].sum()), "identical": bool(okm)}
    pc = A.single_defaults(max_hits=300, num_seeds=12, max_k=15)
    gc, wc = L.characterize(h, pc, s1), ref.characterize(hc, pc, s1)
    out["characterize"] = {"reads": mm, "tuples": int(wc[0][-1]), "identical": bool(all(np.array_equal(a, b) for a, b in zip(gc, wc)))}
    cg, eg = L.cigar(h, s0, g["location"], g["direction"], False)
    cw, ew = ref.cigar(hc, s0, w["location"], w["direction"], False)
    out["cigar"] = {"reads": m, "identical": bool(cg == cw and np.array_equal(eg, ew))}
print(json.dumps(out))
