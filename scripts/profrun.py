"""Cycle-accounting run of the paired path (SNAPB200_PROF=1): where does a pair's time go?"""
import os, sys, time
os.environ["SNAPB200_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("PROF_PLAIN"):  # time the production build instead (no phase breakdown)
    del os.environ["SNAPB200_PROF"]
else:                             # the cycle accounting lives in the -DSNAPB200_PROFILE build (python -m snap_rnaseq_b200.build --profile)
    os.environ["SNAPB200_SO"] = os.path.join(ROOT, "snap_rnaseq_b200", "libsnapb200_prof.so")
import numpy as np
import snap_rnaseq_b200 as S
from snap_rnaseq_b200 import synth, _abi as A
import bench
L = S.lib(0)
if len(sys.argv) > 2 and sys.argv[2] == "c3":
    bench.GENOME_CONTIGS = [25_000_000] * 124
    bench.READ_LEN, bench.ERR_RATE = 150, 0.01
if os.environ.get("PROF_NOREPEAT"):  # ordinary pairs only: no repeat families in the genome
    contigs = synth.random_contigs(bench.GENOME_CONTIGS, seed=20)
else:
    contigs = bench.make_genome()
bases, offs = synth.snap_layout(contigs, 500)
h = L.build_index(bases, offs, list(contigs), seed_len=20)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
b0, b1 = bench.make_pairs(contigs, n, 1000)
sess = S.Session(L, h, n, 256)
sess.upload(0, b0); sess.upload(1, b1)
p = A.paired_defaults()
sweep = os.environ.get("PROF_SWEEP", "").split(",") if os.environ.get("PROF_SWEEP") else [None] * 3
for it, cfg in enumerate(sweep):
    if cfg is not None:
        os.environ[os.environ.get("PROF_SWEEP_VAR", "SNAPB200_CTAS_PER_SM")] = cfg
    sess.run_paired(p)
    ms, launches, _ = sess.last_run()
    print("run %d total %.1f ms main %.1f ms launches %d" % (it, ms, sess.main_kernel_ms(), launches))
out = np.zeros(n, A.PAIRED_RESULT); sess.download_paired(out)
lv = out["n_lv_calls"]
print("lv calls per pair: mean %.1f median %d p90 %d p99 %d p999 %d max %d; pairs with >1000: %d; share of lv in top 1%% pairs: %.2f" % (
    lv.mean(), np.median(lv), np.percentile(lv, 90), np.percentile(lv, 99), np.percentile(lv, 99.9), lv.max(), (lv > 1000).sum(),
    np.sort(lv)[-n // 100:].sum() / lv.sum()))
