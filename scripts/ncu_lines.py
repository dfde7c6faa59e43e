"""Per-source-line warp-stall samples of one ncu report (captured with --import-source on, built with -lineinfo).

usage: ncu_lines.py <report.ncu-rep> [top]
Reads `ncu --page source --print-source cuda,sass --csv`, whose rows with a line number carry the metrics of all SASS
instructions ncu attributes to that source line (inlining included), and prints the lines with the most samples and
their split into the main stall reasons.
"""
import csv, subprocess, sys, os, collections
rep = os.path.abspath(sys.argv[1])
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True, cwd="/tmp").stdout
rows = list(csv.reader(out.split("\n")))
cur_file, hdr, items = "?", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; col = {}
        for i, h in enumerate(hdr):
            col.setdefault(h, i)
        continue
    if hdr is None or len(r) < len(hdr) or r[0] in ("", "Function Name") or not r[0].isdigit():
        continue
    def g(name):  # counted from the right: a source line with quotes in it can split into extra fields
        v = r[len(r) - (len(hdr) - col[name])]
        try:
            return int(v)
        except ValueError:
            return 0
    items.append((cur_file, int(r[0]), r[1].strip(), g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed"),
                  g("stall_no_inst"), g("stall_long_sb"), g("stall_wait"), g("stall_short_sb"), g("stall_branch_resolving")))
ts = sum(i[3] for i in items) or 1
ti = sum(i[4] for i in items) or 1
print(f"total samples {ts}, warp instructions {ti}; columns: %samples %instructions threads/inst | no_inst long_sb wait short_sb branch (as % of the line's samples)")
byfile = collections.Counter()
for i in items:
    byfile[i[0]] += i[3]
print("by file:", ", ".join(f"{f} {100*v/ts:.1f}%" for f, v in byfile.most_common()))
for f, n, src, s, ie, te, ni, lsb, w, ssb, br in sorted(items, key=lambda x: -x[3])[:top]:
    d = max(s, 1)
    print(f"{100*s/ts:5.1f}% {100*ie/ti:5.1f}% {te/max(ie,1):4.1f} | {100*ni/d:3.0f} {100*lsb/d:3.0f} {100*w/d:3.0f} {100*ssb/d:3.0f} {100*br/d:3.0f} | {f}:{n}  {src[:110]}")
