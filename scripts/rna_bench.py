"""The device side of the RNA pair loop alone (snapb200_rna_batch_*): C4 workload (BASELINE.json configs[3]), batches of `batch`
pairs submitted from `threads` host threads with two batch objects each, as the extension's worker threads do -- without the
reference's host code around it.  SNAPB200_RNA_TIMING=1 prints where every batch spends its time.
usage: rna_bench.py [pairs] [genome_mbp] [batch] [threads]"""
import json, os, subprocess, sys, tempfile, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import snap_rnaseq_b200 as S
from snap_rnaseq_b200 import _abi as A, synth
REF = os.path.join(ROOT, "oracle", "_ref", "snap-rna")
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
mbp = int(sys.argv[2]) if len(sys.argv) > 2 else 40
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 32768
threads = int(sys.argv[4]) if len(sys.argv) > 4 else 4
d = tempfile.mkdtemp(prefix="rnabench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
contigs = {"chrDecoy": synth.random_contigs([2000], seed=99)["chr1"]}
contigs.update(synth.random_contigs([mbp * 500_000] * 2, seed=20))
synth.inject_repeats({k: v for k, v in contigs.items() if k != "chrDecoy"}, frac=0.04, seed=21)
synth.write_fasta(os.path.join(d, "g.fa"), contigs)
synth.make_gtf(os.path.join(d, "a.gtf"), contigs)
for cmd in ([REF, "index", "g.fa", "gidx", "-s", "20", "-t16"], [REF, "transcriptome", "a.gtf", "g.fa", "tidx", "-t16", "-s", "20"]):
    subprocess.run(cmd, cwd=d, check=True, stdout=subprocess.DEVNULL)
r0, r1 = synth.simulate_rna(contigs, os.path.join(d, "a.gtf"), pairs, 100, seed=8)
L = S.lib(0)
hg, ht = L.load_index(os.path.join(d, "gidx")), L.load_index(os.path.join(d, "tidx"))
ann = L.annotation_open(hg, ht, os.path.join(d, "a.gtf"))
P = A.rna_defaults()
chunks = [(lo, min(pairs, lo + batch)) for lo in range(0, pairs, batch)]
todo = list(range(len(chunks)))
lock = threading.Lock()
stats = {"needs_host": 0, "unaligned": 0, "events": 0, "device_ms": 0.0}


def worker():
    objs = [L.rna_batch_create(ann, hg, ht) for _ in range(2)]
    inflight = []
    while True:
        with lock:
            k = todo.pop(0) if todo else None
        if k is not None:
            lo, hi = chunks[k]
            o = objs[len(inflight) % 2] if len(inflight) < 2 else None
            if o is None:
                o = inflight.pop(0)
                collect(L.rna_batch_wait(o))
            L.rna_batch_submit(o, P, r0.slice(lo, hi), r1.slice(lo, hi))
            inflight.append(o)
        else:
            for o in inflight:
                collect(L.rna_batch_wait(o))
            break
    for o in objs:
        L.rna_batch_destroy(o)


def collect(out):
    with lock:
        stats["needs_host"] += int(out["needs_host"].sum())
        stats["unaligned"] += int((out["events"]["unaligned"] > 0).sum())
        stats["events"] += int((out["events"]["kind"] > 0).sum())
        stats["device_ms"] += out["device_ms"]


for rep in range(2):  # the second pass is warm (buffers allocated, sessions created)
    todo[:] = list(range(len(chunks)))
    for k in stats:
        stats[k] = 0
    t0 = time.perf_counter()
    th = [threading.Thread(target=worker) for _ in range(threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    print(json.dumps({"pass": rep, "pairs": pairs, "batch": batch, "threads": threads, "wall_s": dt, "reads_per_s": 2 * pairs / dt, **stats}), flush=True)
import shutil
shutil.rmtree(d, ignore_errors=True)
