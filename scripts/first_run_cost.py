import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np
import snap_rnaseq_b200 as S
from snap_rnaseq_b200 import _abi as A, synth
import bench
L = S.lib()
bench.GENOME_CONTIGS = [20_000_000] * 2
contigs = bench.make_genome()
bases, offs = synth.snap_layout(contigs, 500)
b0, b1 = bench.make_pairs(contigs, 20000, 1000)
t = time.perf_counter(); h = L.build_index(bases, offs, list(contigs), seed_len=20); print("index build", time.perf_counter() - t)
p = A.paired_defaults()
for rep in range(2):
    t = time.perf_counter(); sess = S.Session(L, h, 20000, 256); t1 = time.perf_counter()
    sess.upload(0, b0); sess.upload(1, b1); t2 = time.perf_counter()
    sess.run_paired(p); t3 = time.perf_counter()
    sess.run_paired(p); t4 = time.perf_counter()
    print(f"session {rep}: create {t1-t:.3f} upload {t2-t1:.3f} first run_paired {t3-t2:.3f} second {t4-t3:.3f}")
    tp = A.single_defaults(max_hits_to_get=1000)
    t = time.perf_counter(); sess.run_single(tp); t1 = time.perf_counter(); sess.run_single(tp); t2 = time.perf_counter()
    print(f"   first run_single(multihit) {t1-t:.3f} second {t2-t1:.3f}")
    sess.close()
