"""Where the reference's per-pair host work goes in RNA mode (BASELINE.json configs[3]) -- CPU only, test infrastructure.
Times, on one thread of the compiled reference (oracle/_ref): the whole AlignmentFilter section of the paired run loop
(ref_filter_paired_batch = AddAlignment x hits + Filter, SNAPLib/PairedAligner.cpp:582-663) against the part that stays on the host
once the decision comes from the device (ref_filter_replay_events: IncrementReadCount / IntrachromosomalPair / InterchromosomalPair /
UnalignedRead through GTFReader's public methods).
usage: filter_host_profile.py [pairs] [genome_mbp]"""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from snap_rnaseq_b200 import _abi as A, synth  # noqa: E402
import filter_cases as F  # noqa: E402
from test_filter_oracle import FLT_EVENT, FLT_RESULT, flat_tables, genome_pieces  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
mbp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d = tempfile.mkdtemp(prefix="fltprof_")
contigs = {"chrDecoy": synth.random_contigs([2000], seed=99)["chr1"]}
contigs.update(synth.random_contigs([mbp * 500_000] * 2, seed=20))
synth.inject_repeats({k: v for k, v in contigs.items() if k != "chrDecoy"}, frac=0.04, seed=21)
synth.write_fasta(os.path.join(d, "g.fa"), contigs)
synth.make_gtf(os.path.join(d, "a.gtf"), contigs)
for cmd in ([O.REF_BIN, "index", "g.fa", "gidx", "-s", "20", "-t8"], [O.REF_BIN, "transcriptome", "a.gtf", "g.fa", "tidx", "-t8", "-s", "20"]):
    subprocess.run(cmd, cwd=d, check=True, stdout=subprocess.DEVNULL)
(b0, b1), sam_reads = F.reads(contigs, d, n=pairs)
ref = O.ref(threads=8)
hg, ht = ref.load_index(os.path.join(d, "gidx")), ref.load_index(os.path.join(d, "tidx"))
t = time.perf_counter()
hits, genome_res, pp = F.alignments(ref, hg, ht, b0, b1)
t_align = time.perf_counter() - t
t = time.perf_counter()
want = F.run_reference_filter(ref, hg, ht, os.path.join(d, "a.gtf"), os.path.join(d, "want"), sam_reads, hits, genome_res, pp)
t_filter = time.perf_counter() - t
lib = ref.lib
lib.ref_gtf_load.restype = C.c_void_p
g = C.c_void_p(lib.ref_gtf_load(os.path.join(d, "a.gtf").encode(), os.path.join(d, "exp").encode()))
lib.ref_gtf_export(g, os.path.join(d, "gtf.tsv").encode())
T, keep = flat_tables(os.path.join(d, "gtf.tsv"), os.path.join(d, "gidx"), os.path.join(d, "tidx"))
so = os.path.join(ROOT, "tests", "hostsim", "libiohostsim.so")
subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "hostsim", "io_hostsim.cpp")], check=True)
hs = C.CDLL(so)
ch = [ref.characterize(hg, A.single_defaults(max_hits=300, num_seeds=12), b) for b in (b0, b1)]
(n0, l0, r0, s0), (n1, l1, r1, s1) = hits
res = np.ascontiguousarray(genome_res, A.PAIRED_RESULT)
lens0, lens1 = np.diff(b0.offsets), np.diff(b1.offsets)
out, events = np.zeros(b0.n, FLT_RESULT), np.zeros(b0.n, FLT_EVENT)
p64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint64))
p16 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint16))
t = time.perf_counter()
for i in range(b0.n):
    rc = hs.hostsim_filter_pair_flat(C.byref(T), C.c_uint(int(lens0[i])), C.c_uint(int(lens1[i])), C.c_uint(15), C.c_uint(pp.max_spacing), C.c_uint(2),
                                     C.c_int(int(pp.force_spacing)), C.c_int(int(n0[i])), A.p32u(l0[i]), A.p8(r0[i]), A.p32i(s0[i]), C.c_int(int(n1[i])),
                                     A.p32u(l1[i]), A.p8(r1[i]), A.p32i(s1[i]), C.c_void_p(res[i:i + 1].ctypes.data), p64(ch[0][0]), A.p32u(ch[0][1]),
                                     p16(ch[0][2]), p64(ch[1][0]), A.p32u(ch[1][1]), p16(ch[1][2]), C.c_uint(i), C.c_uint(2048), C.c_uint(1 << 16),
                                     C.c_uint(1 << 16), C.c_void_p(out[i:i + 1].ctypes.data), C.c_void_p(events[i:i + 1].ctypes.data))
    assert rc == 0
t_flat = time.perf_counter() - t
same = all(np.array_equal(want[f], out[f]) for f in FLT_RESULT.names)
g2 = C.c_void_p(lib.ref_gtf_load(os.path.join(d, "a.gtf").encode(), os.path.join(d, "replay").encode()))
t_ids = [ln.split("\t")[1] for ln in open(os.path.join(d, "gtf.tsv")) if ln.startswith("T")]
chr_names, _ = genome_pieces(os.path.join(d, "gidx"))
arr = lambda names: (C.c_char_p * len(names))(*[n.encode() for n in names])
t = time.perf_counter()
lib.ref_filter_replay_events(hg, ht, g2, sam_reads[0].byref(), sam_reads[1].byref(), C.c_uint(15), C.c_void_p(events.ctypes.data), arr(t_ids), arr(chr_names))
t_replay = time.perf_counter() - t
ic = np.zeros(4, np.uint64)
lib.ref_gtf_interval_counts(g2, ic.ctypes.data_as(C.c_void_p))
# the same without the UnalignedRead calls: what the GTFReader counting alone costs
ev2 = events.copy()
ev2["unaligned"] = 0
g3 = C.c_void_p(lib.ref_gtf_load(os.path.join(d, "a.gtf").encode(), os.path.join(d, "replay2").encode()))
t = time.perf_counter()
lib.ref_filter_replay_events(hg, ht, g3, sam_reads[0].byref(), sam_reads[1].byref(), C.c_uint(15), C.c_void_p(ev2.ctypes.data), arr(t_ids), arr(chr_names))
t_replay2 = time.perf_counter() - t
nloc = [np.diff(c[0].astype(np.int64)).reshape(-1, 2).sum(axis=1) for c in ch]
un = events["unaligned"]
tuples_unaligned = np.where(un == 1, nloc[0], 0) + np.where(un == 2, nloc[1], 0)
kinds = np.bincount(events["kind"], minlength=4).tolist()
print(json.dumps({"pairs": pairs, "genome_mbp": mbp, "aligners_s_8_threads": t_align, "reference_filter_us_per_pair": 1e6 * t_filter / pairs,
                  "event_replay_us_per_pair": 1e6 * t_replay / pairs, "event_replay_without_unaligned_us_per_pair": 1e6 * t_replay2 / pairs,
                  "intervals_recorded[intra pairs, intra splices, inter pairs, inter splices]": [int(x) for x in ic],
                  "unaligned_tuples_mean": float(tuples_unaligned[un > 0].mean()) if (un > 0).any() else 0, "unaligned_tuples_max": int(tuples_unaligned.max()),
                  "unaligned_tuples_sq_sum": float((tuples_unaligned.astype(np.float64) ** 2).sum()), "flat_filter_python_loop_us_per_pair": 1e6 * t_flat / pairs,
                  "flat_equals_reference": bool(same), "event_kinds": kinds, "unaligned_events": int((events["unaligned"] > 0).sum()),
                  "mean_hits": [float(n0.mean()), float(n1.mean())], "max_hits": [int(n0.max()), int(n1.max())],
                  "transcripts": len(t_ids)}))
import shutil
shutil.rmtree(d, ignore_errors=True)
