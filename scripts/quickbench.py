import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import snap_rnaseq_b200 as S
from snap_rnaseq_b200 import synth, _abi as A
L = S.lib(0)
t = time.time()
contigs = synth.random_contigs([25_000_000] * 4, seed=20)
synth.inject_repeats(contigs, frac=0.05, seed=21)
print("genome gen %.1fs" % (time.time() - t)); t = time.time()
bases, offs = synth.snap_layout(contigs, 500)
h = L.build_index(bases, offs, list(contigs), seed_len=20)
print("index build %.1fs" % (time.time() - t), [(f[0], getattr(L.index_info(h), f[0])) for f in A.IndexInfo._fields_]); t = time.time()
n = 500_000
sim = synth.simulate(contigs, n, 100, paired=True, err=0.02, seed=7)
b0, b1 = sim["batches"]
print("sim %.1fs" % (time.time() - t))
sess = S.Session(L, h, n, 128)
sess.upload(0, b0); sess.upload(1, b1)
p = A.paired_defaults()
for it in range(4):
    t = time.time(); sess.run_paired(p); sess.sync(); dt = time.time() - t
    ms, launches, tot = sess.last_run()
    print("run %d: wall %.1f ms, kernel %.1f ms, launches %d -> %.2f M reads/s" % (it, dt * 1e3, ms, launches, 2 * n / ms / 1e3))
out = np.zeros(n, A.PAIRED_RESULT); sess.download_paired(out)
print("status", np.bincount(out["status"].ravel()), "as pair", out["aligned_as_pair"].mean(), "lv/pair", out["n_lv_calls"].mean(), "lookups/pair", out["n_lookups"].mean())
# single
ps = A.single_defaults()
for it in range(3):
    t = time.time(); sess.run_single(ps); sess.sync(); dt = time.time() - t
    ms, launches, tot = sess.last_run()
    print("single run %d: kernel %.1f ms -> %.2f M reads/s" % (it, ms, n / ms / 1e3))
t = time.time(); r = L.paired(h, p, b0, b1); print("e2e paired batch %.1f ms -> %.2f M reads/s" % ((time.time() - t) * 1e3, 2 * n / (time.time() - t) / 1e6))
