"""Attribute ncu warp-stall samples / executed instructions of one kernel to source lines.

usage: ncu_hot_lines.py <report.ncu-rep> <kernel-symbol-substring> [top]
Joins `ncu --page source --csv` (SASS view: per-instruction samples) with `nvdisasm -g` line info of the in-tree
cubin, instruction by instruction (the two listings have the same order).
"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, sym = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "snap_rnaseq_b200", "libsnapb200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout.split("\n")
# walk the function: remember the current (file,line) for every instruction
lines, cur, inside = [], ("?", 0), False
for l in dis:
    if l.startswith(".text.") and sym in l:
        inside = True
        continue
    if inside and l.startswith(".text.") and sym not in l:
        break
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.split("\n")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
S, I, T = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
inst = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
print(f"{len(inst)} SASS instructions in report, {len(lines)} in disassembly")
agg = collections.defaultdict(lambda: [0, 0, 0])
for k, r in enumerate(inst):
    key = lines[k] if k < len(lines) else ("?", 0)
    a = agg[key]
    a[0] += int(r[S]); a[1] += int(r[I]); a[2] += int(r[T])
ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
src_cache = {}
def src(f, n):
    for d in ("snap_rnaseq_b200/csrc",):
        p = os.path.join(root, d, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().split("\n")
            L = src_cache[p]
            return L[n - 1].strip() if 0 < n <= len(L) else ""
    return ""
print(f"total samples {ts}, warp instructions {ti}")
for (f, n), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{100*a[0]/ts:5.1f}% smp {100*a[1]/ti:5.1f}% inst  thr/inst {a[2]/max(a[1],1):4.1f}  {f}:{n}  {src(f,n)[:100]}")
