"""Small fixed workload for ncu captures: one paired step (100k pairs, C2 genome) + one single-end step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
b0, b1 = bench.make_pairs(contigs, n, 1000)
sess = S.Session(L, h, n, 256)
sess.upload(0, b0); sess.upload(1, b1)
for _ in range(2):
    sess.run_paired(A.paired_defaults())
sess.run_single(A.single_defaults())
print("ok", sess.last_run())
