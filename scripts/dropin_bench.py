"""Whole-command-line comparison on BASELINE.json configs[3] (RNA-seq mode): the compiled reference `snap-rna paired` against
`snap-rna-b200 paired` (same host code, GpuAlignerExtension + libsnapb200.so), same indices, same FASTQ, -t <host threads>.
Prints one JSON line: wall clock of both, reads/s, and whether the sorted SAM records are byte-identical.
usage: dropin_bench.py [pairs] [genome_mbp] [threads]"""
import hashlib, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from snap_rnaseq_b200 import synth
REF = os.path.join(ROOT, "oracle", "_ref", "snap-rna")
B200 = os.path.join(ROOT, "oracle", "_ref", "snap-rna-b200")
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
mbp = int(sys.argv[2]) if len(sys.argv) > 2 else 40
threads = int(sys.argv[3]) if len(sys.argv) > 3 else (os.cpu_count() or 8)
d = tempfile.mkdtemp(prefix="dropin_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


def run(cmd):
    t = time.perf_counter()
    r = subprocess.run(cmd, cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise SystemExit(" ".join(cmd) + "\n" + r.stdout[-3000:])
    return time.perf_counter() - t, r.stdout


contigs = {"chrDecoy": synth.random_contigs([2000], seed=99)["chr1"]}
contigs.update(synth.random_contigs([mbp * 500_000] * 2, seed=20))
synth.inject_repeats({k: v for k, v in contigs.items() if k != "chrDecoy"}, frac=0.04, seed=21)
synth.write_fasta(os.path.join(d, "g.fa"), contigs)
synth.make_gtf(os.path.join(d, "a.gtf"), contigs)
t_index, _ = run([REF, "index", "g.fa", "gidx", "-s", "20", f"-t{threads}"])
run([REF, "transcriptome", "a.gtf", "g.fa", "tidx", f"-t{threads}", "-s", "20"])
r0, r1 = synth.simulate_rna(contigs, os.path.join(d, "a.gtf"), pairs, 100, seed=8)
synth.write_fastq_plain(os.path.join(d, "x1.fq"), r0, mate=0)
synth.write_fastq_plain(os.path.join(d, "x2.fq"), r1, mate=1)
res = {}
for tag, exe in (("reference", REF), ("b200", B200)):
    best = None
    for rep in range(1 if os.environ.get("SNAPB200_SHIM_TIMING") else 2):  # second run: page cache and (b200) a warm GPU context
        t, out = run([exe, "paired", "gidx", "tidx", "a.gtf", "x1.fq", "x2.fq", "-o", tag + ".sam", "-t", str(threads)])
        best = t if best is None else min(best, t)
        for l in out.split("\n"):
            if "[snapb200 shim]" in l:
                sys.stderr.write(l + "\n")
    recs = sorted(l for l in open(os.path.join(d, tag + ".sam")) if not l.startswith("@"))
    stats = [l for l in out.split("\n") if l.strip().startswith("16000")][-1:]
    # the run's own throughput figure: the stats line's "Reads/s (at: <ms of the alignment phase>)" (AlignerContext.cpp:372-393)
    align_ms = float(stats[0].split("(at:")[1].split(")")[0]) if stats and "(at:" in stats[0] else None
    side = {}
    for f in sorted(os.listdir(d)):
        if f.startswith(tag + ".") and not f.endswith(".sam"):
            side[f[len(tag) + 1:]] = hashlib.sha1(open(os.path.join(d, f), "rb").read()).hexdigest()
    res[tag] = {"wall_s": best, "reads_per_s_wall": 2 * pairs / best, "records": len(recs),
                "sha1_sorted_records": hashlib.sha1("".join(recs).encode()).hexdigest(), "stats_line": stats,
                "align_phase_s": align_ms / 1e3 if align_ms else None, "reads_per_s_align_phase": 2 * pairs / (align_ms / 1e3) if align_ms else None,
                "outside_align_phase_s": best - align_ms / 1e3 if align_ms else None, "side_files_sha1": side}
same_side = {k: res["reference"]["side_files_sha1"].get(k) == v for k, v in res["b200"]["side_files_sha1"].items()}
print(json.dumps({"config": f"C4 RNA-seq mode: {mbp} Mbp genome + GTF transcriptome, {pairs} 2x100 bp pairs (50 % spliced fragments, 1 % chimeric), -t {threads}",
                  "reference": res["reference"], "b200": res["b200"], "speedup_wall": res["reference"]["wall_s"] / res["b200"]["wall_s"],
                  "speedup_align_phase": (res["reference"]["align_phase_s"] / res["b200"]["align_phase_s"]) if res["b200"]["align_phase_s"] else None,
                  "sam_identical": res["reference"]["sha1_sorted_records"] == res["b200"]["sha1_sorted_records"],
                  "statistics_files_identical": same_side, "reference_index_build_s": t_index,
                  "note": "outside_align_phase_s is index loading plus the reference's own GTF epilogue (GTFReader::AnalyzeReadIntervals / WriteReadCounts, "
                          "AlignerContext.cpp:126-127), unchanged host code that both binaries run"}))
import shutil
shutil.rmtree(d, ignore_errors=True)
