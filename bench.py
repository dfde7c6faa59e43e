#!/usr/bin/env python3
"""bench.py -- reads aligned per second on the paired-end hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo: sm_100a kernels behind the C ABI
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation

Workload (default, BASELINE.json configs[2] "C3", the configuration the metric and the target are quoted on; it fits
one GPU): 3.1 Gbp repeat-injected synthetic genome (124 x 25 Mbp, seeds 20/21; 47 GB of hash tables, built on the device in
about 2 s), index seed length 20, WGsim-model 2x150 bp FR pairs (1 % error, 15 % of mutations indels, fragment 250-450),
`snap-rna paired` default options (-d 15 -n 8 -h 16000 -H 16000 -s 50 1000).  `--config c2` runs configs[1] (100 Mbp,
2x100 bp, 2 %).  One step = one pass of
ChimericPairedEndAligner::align over one batch of pairs per GPU.  The index and genome are replicated per GPU;
reads are sharded (each rank aligns its own batch, no collective on the data path) => weak scaling.  NCCL is used
only for the end-of-run AlignerStats all-reduce and the timing reduction.

JSON keys: see the task contract.  `value` times the kernels with the batch already resident in HBM; `e2e` times
snapb200_paired_batch (the C-ABI call the reference's AlignerExtension would make) from pinned host buffers,
host<->device copies included; `roofline` is the dominant kernel (paired_kernel) against the measured HBM copy
peak; `cpu_baseline` is the compiled reference (oracle/_ref) on the host cores over a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads_aligned_per_sec"
UNIT = "reads/s"
GENOME_CONTIGS = [25_000_000] * 4
READ_LEN = 100
ERR_RATE = 0.02
PAIRS_PER_STEP = 1_000_000          # per GPU
CPU_SAMPLE_PAIRS = 200_000


def measured_traffic(pairs):
    """dram__bytes_read + dram__bytes_write of paired_kernel per launch, from the committed `ncu --set full` capture of the
    same configuration (profiles/r2_traffic.json: bytes per pair of a 200 k-pair launch), scaled to this launch's pairs.
    None for configurations that were not captured."""
    if (sum(GENOME_CONTIGS) // 1_000_000, READ_LEN) != (3100, 150):
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_pair"]) * pairs
    except Exception:
        return None


def int_alu_roofline(pairs, kernel_ms, sm_mhz):
    """The second roof SURVEY.md 8(d) names for this path: integer issue.  Thread-instructions per pair come from the committed
    ncu capture of the same configuration (smsp__inst_executed.sum x threads per instruction / pairs); the peak is
    148 SMs x 128 lanes x the SM clock sampled during the run.  None for configurations that were not captured."""
    if (sum(GENOME_CONTIGS) // 1_000_000, READ_LEN) != (3100, 150):
        return None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            per_pair = float(json.load(f)["thread_instructions_per_pair"])
    except Exception:
        return None
    achieved = per_pair * pairs / (kernel_ms * 1e-3) / 1e12
    peak = 148 * 128 * float(sm_mhz or 1965.0) * 1e6 / 1e12
    return {"bound": "int-alu issue", "kernel": "paired_kernel", "achieved": achieved, "peak": peak, "unit": "T thread-instructions/s",
            "frac": achieved / peak, "source": "profiles/r2_traffic.json (ncu --set full capture) x pairs / CUDA-event kernel time"}


def workload_config(n_gpus, pairs):
    mbp = sum(GENOME_CONTIGS) // 1_000_000
    tag = "C3: " if (mbp, READ_LEN) == (3100, 150) else "C2: " if (mbp, READ_LEN) == (100, 100) else ""
    return {"workload": f"{tag}snap paired, {mbp} Mbp repeat-injected synthetic genome ({len(GENOME_CONTIGS)}x25 Mbp), seed 20, 2x{READ_LEN}bp WGsim pairs e={ERR_RATE:.0%}",
            "pairs_per_step_per_gpu": pairs, "read_len": READ_LEN, "options": "-d 15 -n 8 -h 16000 -H 16000 -s 50 1000 -D 2",
            "parallelism": f"reads sharded over {n_gpus} GPU(s), index replicated; 2 host threads / 2 streams per GPU keep two batches in flight", "l2": "inputs larger than L2 (index + genome >= 1.76 GB, 0.4 GB batch per step)"}


def make_genome():
    from snap_rnaseq_b200 import synth
    contigs = synth.random_contigs(GENOME_CONTIGS, seed=20)
    synth.inject_repeats(contigs, frac=0.05, seed=21)
    return contigs


def make_pairs(contigs, n, seed):
    from snap_rnaseq_b200 import synth
    sim = synth.simulate(contigs, n, READ_LEN, paired=True, err=ERR_RATE, indel_frac=0.15, seed=seed)
    return sim["batches"]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def pinned_batch(batch):
    """Copy a Batch into pinned host memory (torch) and return (torch tensors, raw pointers)."""
    import torch
    t = {k: torch.from_numpy(getattr(batch, k).copy()).pin_memory() for k in ("offsets", "bases", "quals")}
    return t


def run_ours(args):
    import torch
    import torch.distributed as dist

    import snap_rnaseq_b200 as S
    from snap_rnaseq_b200 import _abi as A
    from snap_rnaseq_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this library has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = S.lib(local)
    pairs = args.pairs
    contigs = make_genome()
    bases, offs = synth.snap_layout(contigs, 500)
    t0 = time.time()
    h = L.build_index(bases, offs, list(contigs), seed_len=20, device=local)
    t_index = time.time() - t0
    b0, b1 = make_pairs(contigs, pairs, seed=1000 + rank)
    params = A.paired_defaults()

    # ---- kernels on HBM-resident batches (value) ----
    # Two sessions (two streams, two resident batches) driven by two host threads, the way the reference's worker
    # threads would drive the C ABI: the tail and the fallback launches of one batch overlap the head of the next.
    import threading
    batches = [(b0, b1), make_pairs(contigs, pairs, seed=2000 + rank)]
    sessions = []
    for (x0, x1) in batches:
        s_ = S.Session(L, h, pairs, max(128, READ_LEN + 8))
        s_.upload(0, x0)
        s_.upload(1, x1)
        s_.sync()
        sessions.append(s_)
    sess = sessions[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(which, n_steps, acc):
        for _ in range(n_steps):
            sessions[which].run_paired(params)   # returns after its stream has drained (counters are read back)
            acc["launches"] += sessions[which].last_run()[1]
            acc["main_ms"].append(sessions[which].main_kernel_ms())

    def run_both(n_steps):
        accs = [{"launches": 0, "main_ms": []}, {"launches": 0, "main_ms": []}]
        split = [(n_steps + 1) // 2, n_steps // 2]
        th = [threading.Thread(target=run_steps, args=(i, split[i], accs[i])) for i in range(2) if split[i]]
        for t in th:
            t.start()
        for t in th:
            t.join()
        return accs

    run_both(max(args.warmup, 2))
    L.stats_reset(h)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    accs = run_both(args.steps)
    for s_ in sessions:
        s_.sync()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    clocks = sampler.finish()
    launches = sum(a["launches"] for a in accs)
    main_ms = [m for a in accs for m in a["main_ms"]]
    stats = L.stats(h)
    out = np.zeros(pairs, A.PAIRED_RESULT)
    sess.download_paired(out)

    # ---- end to end through the C ABI from pinned host memory (e2e): two host threads, each with its own batch ----
    res_pins, rbs = [], []
    for (x0, x1) in batches:
        pin0, pin1 = pinned_batch(x0), pinned_batch(x1)
        res_pin = torch.empty(pairs * A.PAIRED_RESULT.itemsize, dtype=torch.uint8).pin_memory()
        res_pins.append(res_pin)
        rbs.append((pin0, pin1, A.ReadBatch(pairs, C.cast(pin0["offsets"].data_ptr(), C.POINTER(C.c_uint32)), C.cast(pin0["bases"].data_ptr(), C.POINTER(C.c_uint8)),
                                            C.cast(pin0["quals"].data_ptr(), C.POINTER(C.c_uint8))),
                    A.ReadBatch(pairs, C.cast(pin1["offsets"].data_ptr(), C.POINTER(C.c_uint32)), C.cast(pin1["bases"].data_ptr(), C.POINTER(C.c_uint8)),
                                C.cast(pin1["quals"].data_ptr(), C.POINTER(C.c_uint8)))))

    def e2e_steps_fn(which, n_steps):
        _, _, r0, r1 = rbs[which]
        for _ in range(n_steps):
            rc = L.lib.snapb200_paired_batch(h, C.byref(params), C.byref(r0), C.byref(r1), C.c_void_p(res_pins[which].data_ptr()))
            if rc != 0:
                raise RuntimeError(L.lib.snapb200_last_error())

    def e2e_both(n_steps):
        split = [(n_steps + 1) // 2, n_steps // 2]
        th = [threading.Thread(target=e2e_steps_fn, args=(i, split[i])) for i in range(2) if split[i]]
        for t in th:
            t.start()
        for t in th:
            t.join()

    e2e_both(2)
    barrier()
    t1 = time.perf_counter()
    e2e_steps = max(8, args.steps + (args.steps & 1))  # both host threads get the same number of batches; enough of them that the
    # un-overlapped first upload and last download do not dominate
    e2e_both(e2e_steps)
    torch.cuda.synchronize()
    dt_e2e = time.perf_counter() - t1
    barrier()
    e2e_res = np.frombuffer(res_pins[0].numpy(), dtype=A.PAIRED_RESULT)
    same = all(np.array_equal(e2e_res[f], out[f]) for f in ("location", "mapq", "status", "score", "direction"))

    # ---- reductions: max time over ranks, summed stats (the AlignerStats all-reduce over NCCL) ----
    tv = torch.tensor([dt, dt_e2e], dtype=torch.float64, device="cuda")
    sv = torch.from_numpy(stats.astype(np.int64)).cuda()
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        dist.all_reduce(sv, op=dist.ReduceOp.SUM)
    dt_max, dt_e2e_max = [float(x) for x in tv.cpu()]
    tot = sv.cpu().numpy()

    if rank == 0:
        reads_per_step = 2 * pairs * world
        value = reads_per_step * args.steps / dt_max
        e2e_value = reads_per_step * e2e_steps / dt_e2e_max
        h2d = int(b0.bases.size + b0.quals.size + b0.offsets.size * 4 + b1.bases.size + b1.quals.size + b1.offsets.size * 4)
        d2h = int(pairs * A.PAIRED_RESULT.itemsize)
        # algorithmic bytes of the dominant kernel per launch (counters are this rank's, per step)
        steps = args.steps
        per = lambda i: float(stats[i]) / steps
        n_lv = per(9)   # locations scored per step (device counter; both resident batches contribute steps)
        alg_bytes = 12.0 * per(12) + 4.0 * per(13) + n_lv * (READ_LEN + 31) + 2.0 * (2 * READ_LEN) * pairs + 56.0 * pairs
        k_ms = float(np.mean(main_ms))
        peak, peak_src = measured_peaks()
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u32 integer + f64 probabilities", "data": "synthetic", "config": workload_config(world, pairs),
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "matches_resident_run": bool(same)},
            "roofline": {"bound": "hbm", "kernel": "paired_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "traffic": measured_traffic(pairs), "kernel_ms": k_ms, "algorithmic_bytes_per_launch": alg_bytes,
                         "per_pair": {"table_probes": per(12) / pairs, "hit_words": per(13) / pairs, "lv_locations": n_lv / pairs,
                                      "lookups": float(out["n_lookups"].mean())}},
            "roofline_int_alu": int_alu_roofline(pairs, k_ms, (clocks or {}).get("sm_mhz")),
            "index_build_s": t_index,
            "stats_allreduce": {"total_reads": int(tot[0]), "single_hits": int(tot[2]), "multi_hits": int(tot[3]), "not_found": int(tot[4]),
                                "aligned_as_pairs": int(tot[6]), "locations_scored": int(tot[9]), "lookups": int(tot[8])},
            "aligned_fraction": float((out["status"] != 0).mean()),
        }
        try:
            line["probe_stage"] = probe_stage(L, h, bases.size)
        except Exception as e:  # diagnostics only
            line["probe_stage"] = {"error": str(e)}
        if world == 1:
            try:
                line["io_edges"] = io_edges_stage(L, h, b0, b1, out)
            except Exception as e:  # diagnostics only: the stages either side of the hot path
                line["io_edges"] = {"error": str(e)}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(L, h, b0, b1, params, out)
        if world == 1 and not args.no_dropin:
            # The reference's own command line with and without the extension (BASELINE.json configs[3], RNA mode): what a user of
            # `snap-rna paired` sees.  Runs after this process has released the GPU-resident index.
            for s_ in sessions:
                s_.close()
            sessions.clear()
            L.close_index(h)
            h = None
            try:
                sys.path.insert(0, os.path.join(ROOT, "scripts"))
                from dropin_bench import run_dropin
                line["dropin"], _ = run_dropin(pairs=args.dropin_pairs, mbp=40, reps=2)
            except Exception as e:  # diagnostics: the headline numbers above stand without it
                line["dropin"] = {"error": str(e)[-500:]}
        emit(line)
    for s_ in sessions:
        s_.close()
    if h is not None:
        L.close_index(h)
    if world > 1:
        dist.destroy_process_group()


def io_edges_stage(L, h, b0, b1, results):
    """The stages either side of the hot path on this step's batch: FASTQ text of both mates ->
    snapb200_fastq_parse -> read arrays, and this step's alignments -> snapb200_sam_batch -> SAM text.  Kernel time (CUDA events, no
    copies), end-to-end time of the C-ABI call from pinned host buffers, and the algorithmic bytes (text + arrays, each counted once)
    against the measured HBM copy peak.  Parity and the reference's own reader / writer timed beside: scripts/io_bench.py,
    profiles/README.md (session 3)."""
    import torch
    from snap_rnaseq_b200 import _abi as A
    from snap_rnaseq_b200 import synth

    def pinned(n, dtype):
        return torch.empty(n, dtype=dtype).pin_memory().numpy()

    pairs = b0.n
    texts = []
    for e, b in enumerate((b0, b1)):
        t = synth.fastq_fixed(b, e)
        p = pinned(t.size, torch.uint8)
        np.copyto(p, t)
        texts.append(p)
    nb = texts[0].size
    bufs = [(pinned(pairs + 1, torch.int32).view(np.uint32), pinned(pairs + 1, torch.int32).view(np.uint32), pinned(nb, torch.uint8),
             pinned(nb, torch.uint8), pinned(nb, torch.uint8), pinned(pairs, torch.int16).view(np.uint16), pinned(pairs, torch.int16).view(np.uint16))
            for _ in range(2)]
    for _ in range(3):  # the third repetition counts
        reads, k_fq, t_fq = [], 0.0, 0.0
        for t, bf in zip(texts, bufs):
            t0 = time.perf_counter()
            r, used = L.fastq_parse(t, 3, bufs=bf)
            t_fq += time.perf_counter() - t0
            k_fq += L.io_last_kernel_ms()[0]
            if used != t.size or r.n != pairs:
                raise RuntimeError("fastq_parse did not consume the text")
            reads.append(r)
    same_reads = bool(np.array_equal(reads[0].bases[:b0.bases.size], b0.bases) and np.array_equal(reads[1].quals[:b1.quals.size], b1.quals))
    aln = []
    for e in range(2):
        a = np.zeros(pairs, A.SAM_ALIGNMENT)
        for f in ("location", "mapq", "status", "direction"):
            a[f] = results[f][:, e]
        aln.append(a)
    buf = pinned(2 * pairs * (2 * READ_LEN + 220), torch.uint8)
    for _ in range(3):
        t0 = time.perf_counter()
        sam, lo = L.sam(h, reads[0], reads[1], aln[0], aln[1], False, None, out=buf)
        t_sam = time.perf_counter() - t0
        k_sam = L.io_last_kernel_ms()[1]
    peak, _ = measured_peaks()
    arrays = sum(int(r.offsets[-1]) * 2 + int(r.id_offsets[-1]) + r.n * 12 for r in reads)
    fq_bytes = sum(t.size for t in texts) + arrays
    mapped = int((aln[0]["status"] != 0).sum() + (aln[1]["status"] != 0).sum())
    sam_bytes = arrays + 2 * pairs * 12 + mapped * (READ_LEN + 80) + int(lo[-1])
    n_lines = int(np.count_nonzero(np.frombuffer(sam, np.uint8) == 10)) if len(sam) < (1 << 31) else None
    return {
        "fastq_parse": {"reads_per_s_kernels": 2 * pairs / (k_fq * 1e-3), "reads_per_s_e2e": 2 * pairs / t_fq, "kernel_ms": k_fq, "e2e_ms": t_fq * 1e3,
                        "text_bytes": int(sum(t.size for t in texts)), "arrays_match_the_batch": same_reads,
                        "roofline": {"bound": "hbm", "achieved": fq_bytes / (k_fq * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": fq_bytes / (k_fq * 1e-3) / 1e9 / peak}},
        "sam_text": {"reads_per_s_kernels": 2 * pairs / (k_sam * 1e-3), "reads_per_s_e2e": 2 * pairs / t_sam, "kernel_ms": k_sam, "e2e_ms": t_sam * 1e3,
                     "sam_bytes": int(lo[-1]), "lines": n_lines,
                     "roofline": {"bound": "hbm", "achieved": sam_bytes / (k_sam * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                  "frac": sam_bytes / (k_sam * 1e-3) / 1e9 / peak}},
        "parity": "tests/test_io_edges.py, scripts/io_bench.py (2 M reads / 752 MB of SAM text identical to the reference's FASTQReader and SAMFormat)",
    }


def probe_stage(L, h, n_bases):
    """Stage 2 (index probes) in isolation against the random-32-byte-sector gather ceiling measured in the same run
    (MEASURED_PEAKS.json only has the sequential copy peak).  See scripts/probe_roofline.py / profiles/README.md."""
    info = L.index_info(h)
    table_bytes = int(info.hash_table_entries) * 12
    rng = np.random.default_rng(3)
    n = 1 << 24
    pos = rng.integers(500, n_bases - 600, size=n, dtype=np.uint32)
    ms, slots, counts, hits, gms = C.c_float(), C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_float()
    L._check(L.lib.snapb200_probe_bench(h, C.c_uint32(n), pos.ctypes.data_as(C.c_void_p), C.c_uint32(5), C.byref(ms), C.byref(slots),
                                        C.byref(counts), C.byref(hits)), "probe_bench")
    L._check(L.lib.snapb200_gather_bench(C.c_int(L.device), C.c_uint64(table_bytes), C.c_uint32(n), C.c_uint32(5), C.byref(gms)), "gather_bench")
    sectors = slots.value + counts.value
    ach = sectors / (ms.value * 1e-3)
    peak = n / (gms.value * 1e-3)
    return {"kernel": "probe_bench_kernel (lookup_seed, one lane per seed)", "lookups": n, "ms": ms.value, "lookups_per_s": n / (ms.value * 1e-3),
            "table_slots_per_lookup": slots.value / n, "bound": "hbm random 32B sectors", "achieved": ach, "peak": peak, "unit": "sectors/s",
            "frac": ach / peak, "peak_source": f"random-sector gather over {table_bytes >> 20} MiB measured in this run (snapb200_gather_bench)"}


def cpu_baseline(L, h, b0, b1, params, gpu_out):
    """The compiled reference (oracle/_ref) on the host cores, over a bounded sample of the same batch.  The index
    it loads is the device-built one written in the reference's file format (lookup-equivalent; tests check that)."""
    import tempfile

    from oracle import oracle as O
    from snap_rnaseq_b200 import _abi as A
    n = min(CPU_SAMPLE_PAIRS, b0.n)
    s0, s1 = b0.slice(0, n), b1.slice(0, n)
    cores = os.cpu_count() or 1
    with tempfile.TemporaryDirectory(dir=scratch_dir(sum(GENOME_CONTIGS) * 18)) as tmp:
        d = os.path.join(tmp, "idx")
        L.save_index(h, d)
        if O.have_ref():
            impl, kind = O.ref(threads=cores), "reference"
        else:
            impl, kind, cores = O.port(), "port", 1
        hc = impl.load_index(d)
        if kind == "reference":  # per-thread aligner objects constructed outside the timed region, as in --impl reference
            lib = impl.lib
            lib.ref_paired_pool_create.restype = C.c_void_p
            pool = C.c_void_p(lib.ref_paired_pool_create(hc, C.byref(params), C.c_int(cores)))
            res = np.zeros(n, A.PAIRED_RESULT)
            t0 = time.perf_counter()
            lib.ref_paired_pool_run(pool, s0.byref(), s1.byref(), res.ctypes.data_as(C.c_void_p))
            dt = time.perf_counter() - t0
            lib.ref_paired_pool_destroy(pool)
        else:
            t0 = time.perf_counter()
            res = impl.paired(hc, params, s0, s1)
            dt = time.perf_counter() - t0
    fields = ("location", "mapq", "status", "score", "direction", "p_all", "p_best")  # the two FP64 probabilities bit for bit as well
    agree = all(np.array_equal(res[f], gpu_out[f][:n]) for f in fields)
    same = np.ones((n, 2), bool)  # per read: location, strand, edit distance, MAPQ and status all identical
    for f in fields:
        eq = res[f] == gpu_out[f][:n]
        same &= eq if eq.ndim == 2 else eq[:, None]
    return {"value": 2 * n / dt, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"first {n} pairs of the rank-0 batch, {cores} threads, aligner calls only (no I/O)",
            "bit_exact_vs_gpu_on_sample": bool(agree), "reads_bit_exact_pct": float(100.0 * same.mean()), "reads_compared": int(2 * n)}


def scratch_dir(need_bytes):
    """A RAM-backed scratch directory when it has room for `need_bytes`, else the regular temp dir."""
    import shutil
    for d in ("/dev/shm", None):
        try:
            if d is None or (os.path.isdir(d) and shutil.disk_usage(d).free > need_bytes * 1.2):
                return d
        except OSError:
            pass
    return None


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref, compiled from /root/reference):
    its GenomeIndex loader, its ChimericPairedEndAligner over IntersectingPairedEndAligner + BaseAligner, all host threads.
    Rank 0 only.  The index files: at 100 Mbp the reference's own indexer builds them (14 s); at 3.1 Gbp it needs ~70 GB and
    tens of minutes, so there the lookup-equivalent index built on the GPU is written in the reference's
    file format (tests/test_cuda_parity.py::test_index_build_equivalence) and the reference loads that.  Nothing of this
    repo is on the timed path."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tempfile

    from oracle import oracle as O
    from snap_rnaseq_b200 import _abi as A
    from snap_rnaseq_b200 import synth
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cores = os.cpu_count() or 1
    if not O.have_ref():
        emit({"impl": "reference", "unavailable": "oracle/_ref not present on this box"})
        return
    contigs = make_genome()
    n = args.pairs                       # the repo arm's pairs per step (same `config`); each step is one pass over them
    b0, b1 = make_pairs(contigs, n, seed=1000)
    params = A.paired_defaults()
    mbp = sum(GENOME_CONTIGS) // 1_000_000
    index_how = "reference indexer (snap-rna index -s 20)"
    with tempfile.TemporaryDirectory(dir=scratch_dir(mbp * 18_000_000)) as tmp:
        d = os.path.join(tmp, "idx")
        if mbp <= 200:
            fa = os.path.join(tmp, "g.fa")
            synth.write_fasta(fa, contigs)
            O.ref_build_index(fa, d, seed_len=20, threads=cores)
        else:
            import snap_rnaseq_b200 as S
            L = S.lib(int(os.environ.get("LOCAL_RANK", "0")))
            bases, offs = synth.snap_layout(contigs, 500)
            h = L.build_index(bases, offs, list(contigs), seed_len=20)
            L.save_index(h, d)
            L.close_index(h)
            del bases
            index_how = "lookup-equivalent index built on the GPU, saved in the reference's file format (the reference's indexer needs ~70 GB and tens of minutes at 3.1 Gbp; --config c2 uses the reference's own indexer)"
        impl = O.ref(threads=cores)
        hc = impl.load_index(d)
        # one set of aligner objects per thread, constructed once and kept across steps, as a worker thread of the reference keeps
        # them for a whole run (SNAPLib/PairedAligner.cpp:459-527): construction is outside the timed region
        lib = impl.lib
        lib.ref_paired_pool_create.restype = C.c_void_p
        pool = C.c_void_p(lib.ref_paired_pool_create(hc, C.byref(params), C.c_int(cores)))
        res = np.zeros(n, A.PAIRED_RESULT)

        def step(x0, x1, out):
            rc = lib.ref_paired_pool_run(pool, x0.byref(), x1.byref(), out.ctypes.data_as(C.c_void_p))
            if rc != 0:
                raise RuntimeError("ref_paired_pool_run failed")

        w = min(n, 50_000)
        for _ in range(args.warmup):     # W untimed warm-up steps on a slice (page cache, allocator, branch predictors)
            step(b0.slice(0, w), b1.slice(0, w), res[:w])
        steps = args.steps
        t0 = time.perf_counter()
        for _ in range(steps):
            step(b0, b1, res)
        dt = time.perf_counter() - t0
        lib.ref_paired_pool_destroy(pool)
    value = 2 * n * steps / dt
    cfg = workload_config(world, n)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/u32 integer + f64 probabilities", "data": "synthetic", "config": cfg, "index": index_how,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                             "sample": f"{n} pairs per step x {steps} steps, {cores} threads with their aligner objects kept across steps, "
                                       "ChimericPairedEndAligner::align only (no I/O); warm-up steps run on the first 50000 pairs"},
            "aligned_fraction": float((res["status"] != 0).mean()),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print (NCCL's version banner goes to stdout) is sent to stderr; the one JSON line is
    written to the real stdout by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropin", action="store_true", help="skip the `dropin` key (the reference's command line with / without the extension)")
    ap.add_argument("--dropin-pairs", type=int, default=300_000)
    ap.add_argument("--config", default="c3", choices=["c2", "c3", "custom"],
                    help="c3 (default) = BASELINE.json configs[2]: 3.1 Gbp genome, 2x150 bp, 1 %% error; c2 = configs[1]: 100 Mbp, 2x100 bp, 2 %%")
    ap.add_argument("--genome-mbp", type=int, default=None, help="with --config custom: synthetic genome size (contigs of 25 Mbp)")
    ap.add_argument("--read-len", type=int, default=None, help="with --config custom: read length")
    ap.add_argument("--err", type=float, default=None, help="with --config custom: per-base mutation rate of the simulated reads")
    args = ap.parse_args()
    preset = {"c3": (3100, 150, 0.01), "c2": (100, 100, 0.02), "custom": (100, 100, 0.02)}[args.config]
    args.genome_mbp = args.genome_mbp or preset[0]
    args.read_len = args.read_len or preset[1]
    args.err = args.err if args.err is not None else preset[2]
    global READ_LEN, ERR_RATE
    READ_LEN = args.read_len
    ERR_RATE = args.err
    global GENOME_CONTIGS
    GENOME_CONTIGS = [25_000_000] * max(1, args.genome_mbp // 25)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
