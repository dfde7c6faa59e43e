"""Single-end path timing: BaseAligner::AlignRead (snap-rna single defaults) and the multi-hit form the RNA pipeline runs on
the transcriptome index (maxHitsToGet 1000, -h 16000 -n 8 -d 15), plus the C5 stress shape.
usage: single_bench.py [n_reads] [genome_mbp] [read_len] [err]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snap_rnaseq_b200 as S
from snap_rnaseq_b200 import synth, _abi as A
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
mbp = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
rlen = int(sys.argv[3]) if len(sys.argv) > 3 else 100
err = float(sys.argv[4]) if len(sys.argv) > 4 else 0.02
L = S.lib(0)
contigs = synth.random_contigs([25_000_000] * (mbp // 25), seed=20)
synth.inject_repeats(contigs, frac=0.05, seed=21)
bases, offs = synth.snap_layout(contigs, 500)
h = L.build_index(bases, offs, list(contigs), seed_len=20)
sim = synth.simulate(contigs, n, rlen, paired=True, err=err, indel_frac=0.15, seed=1000)
b0, b1 = sim["batches"]
sess = S.Session(L, h, n, 500)
sess.upload(0, b0); sess.upload(1, b1)
cases = [("single defaults (-h 300 -d 14 -n 25)", A.single_defaults()),
         ("transcriptome aligner of the paired loop (-h 16000 -d 15 -n 8, maxHitsToGet 1000)", A.single_defaults(max_hits=16000, num_seeds=8, max_k=15, max_hits_to_get=1000))]
if rlen >= 250:
    cases.append(("C5 stress: -d 20", A.single_defaults(max_k=20)))
for name, p in cases:
    for it in range(3):
        sess.run_single(p)
        ms, launches, _ = sess.last_run()
    out = np.zeros(n, A.SINGLE_RESULT); sess.download_single(out)
    print("%s: %.1f ms / %d reads = %.2f M reads/s (%d launches); scored/read mean %.1f max %d; aligned %.3f" % (
        name, ms, n, n / ms / 1e3, launches, out["n_scored"].mean(), out["n_scored"].max(), (out["status"] != 0).mean()))
pp = A.paired_defaults(max_k=20) if rlen >= 250 else A.paired_defaults()
for it in range(3):
    sess.run_paired(pp)
    ms, launches, _ = sess.last_run()
print("paired (%s): %.1f ms / %d pairs = %.2f M reads/s" % ("-d 20" if rlen >= 250 else "defaults", ms, n, 2 * n / ms / 1e3))
