"""Static SASS size of one kernel attributed to source lines (nvdisasm -g line info).  usage: sass_lines.py <kernel-substring> [top]"""
import os, re, subprocess, sys, tempfile, collections
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sym = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "snap_rnaseq_b200", "libsnapb200.so")], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout.split("\n")
inside, cur = False, ("?", 0)
agg = collections.Counter()
for l in dis:
    m = re.match(r"\s+\.section\s+\.text\.(\S+?),", l)
    if m:
        inside = sym in m.group(1); continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        agg[cur] += 1
tot = sum(agg.values())
print("total", tot, "instructions", tot * 16 / 1024, "KB")
byfile = collections.Counter()
for (f, n), c in agg.items():
    byfile[f] += c
print(", ".join(f"{f} {c}" for f, c in byfile.most_common()))
src = {}
def line(f, n):
    p = os.path.join(root, "snap_rnaseq_b200", "csrc", f)
    if os.path.exists(p):
        if p not in src: src[p] = open(p).read().split("\n")
        return src[p][n - 1].strip() if 0 < n <= len(src[p]) else ""
    return ""
for (f, n), c in agg.most_common(top):
    print(f"{c:5d}  {f}:{n}  {line(f, n)[:120]}")
