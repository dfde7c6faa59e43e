"""SASS instruction count (and bytes) per function / label of the in-tree libsnapb200.so.  usage: sass_size.py [kernel-substring]"""
import os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "snap_rnaseq_b200", "libsnapb200.so")], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout.split("\n")
want = sys.argv[1] if len(sys.argv) > 1 else ""
sec, fn, counts = None, None, {}
for l in dis:
    m = re.match(r"\s+\.section\s+\.text\.(\S+?),", l)
    if m:
        sec = m.group(1); fn = sec; continue
    m = re.match(r"(\$?[\w$.]+):\s*$", l)   # labels: device functions inside a section start with their mangled name
    if m and sec and not m.group(1).startswith(".L") and not m.group(1).startswith(".text"):
        fn = m.group(1)
    if sec and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        counts[(sec, fn)] = counts.get((sec, fn), 0) + 1
tot = {}
for (s, f), c in counts.items():
    tot[s] = tot.get(s, 0) + c
for s, c in sorted(tot.items(), key=lambda x: -x[1]):
    if want not in s or c < 50:
        continue
    print(f"{c:7d} instr {c*16/1024:7.1f} KB  {s[:80]}")
    for (s2, f), c2 in sorted(counts.items(), key=lambda x: -x[1]):
        if s2 == s and f != s:
            print(f"        {c2:7d} {c2*16/1024:7.1f} KB  {f[:90]}")
