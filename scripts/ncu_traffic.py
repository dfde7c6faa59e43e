"""profiles/<tag>_traffic.json from an `ncu --set full` report of one paired_kernel launch: the DRAM bytes and the instruction counts
bench.py scales to its own launch for `roofline.traffic` and `roofline_int_alu` (it cannot run ncu inside the timed run).
usage: ncu_traffic.py <report.ncu-rep> <pairs in the captured launch> <out.json> [description]"""
import csv, json, subprocess, sys
import os
rep, pairs, out = os.path.abspath(sys.argv[1]), int(sys.argv[2]), os.path.abspath(sys.argv[3])
desc = sys.argv[4] if len(sys.argv) > 4 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, cwd="/tmp").stdout
rows = list(csv.reader(raw.split("\n")))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3}.get(u, 1.0)
    return v * scale
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
wi = get("smsp__inst_executed.sum")
tpi = get("smsp__thread_inst_executed_per_inst_executed.ratio")
d = {"kernel": "paired_kernel", "config": desc, "source": f"{rep} (ncu --set full --clock-control none)", "dram_bytes_read": rd, "dram_bytes_write": wr, "pairs": pairs,
     "dram_bytes_per_pair": (rd + wr) / pairs, "warp_instructions": wi, "threads_per_instruction": tpi, "thread_instructions_per_pair": wi * tpi / pairs,
     "gpu_time_ms": get("gpu__time_duration.sum"),
     "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"), "icache_hit_pct": get("sm__icc_request_hit_rate.pct"),
     "l1_hit_pct": get("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": get("lts__t_sector_hit_rate.pct"),
     "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active"), "registers": get("launch__registers_per_thread"),
     "shared_mem_per_block": get("launch__shared_mem_per_block_dynamic")}
json.dump(d, open(out, "w"))
print(json.dumps(d))
