"""Stage-2 (index probe) roofline in isolation: lookups/s and 32-byte sectors/s of the probe kernel on the C2 index,
against a uniformly random 32-byte-sector gather over a buffer of the same size (the honest ceiling for this access
pattern; MEASURED_PEAKS.json only has the sequential copy peak)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snap_rnaseq_b200 as S
from snap_rnaseq_b200 import synth, _abi as A
import bench
L = S.lib(0)
contigs = bench.make_genome()
bases, offs = synth.snap_layout(contigs, 500)
h = L.build_index(bases, offs, list(contigs), seed_len=20)
info = L.index_info(h)
table_bytes = info.hash_table_entries * 12
rng = np.random.default_rng(3)
n = 1 << 24
pos = rng.integers(500, bases.size - 600, size=n, dtype=np.uint32)
ms, slots, counts, hits = C.c_float(), C.c_uint64(), C.c_uint64(), C.c_uint64()
L._check(L.lib.snapb200_probe_bench(h, C.c_uint32(n), pos.ctypes.data_as(C.c_void_p), C.c_uint32(5), C.byref(ms), C.byref(slots), C.byref(counts), C.byref(hits)), "probe_bench")
gms = C.c_float()
L._check(L.lib.snapb200_gather_bench(C.c_int(0), C.c_uint64(table_bytes), C.c_uint32(n), C.c_uint32(5), C.byref(gms)), "gather_bench")
gms_big = C.c_float()
L._check(L.lib.snapb200_gather_bench(C.c_int(0), C.c_uint64(48 << 30), C.c_uint32(n), C.c_uint32(5), C.byref(gms_big)), "gather_bench")
# sectors an ideal implementation must touch: 1 per table slot examined (12-byte entries, ~1.3 sectors when straddling is
# counted: 12/32 of entries straddle a 32-byte boundary... counted as 1 here) + 1 per overflow count word; seeds arrive packed
alg_sectors = slots.value + counts.value
out = {
    "index_table_bytes": int(table_bytes), "n_lookups": n, "probe_kernel_ms": ms.value,
    "lookups_per_s": n / (ms.value * 1e-3), "table_slots_per_lookup": slots.value / n, "count_words_per_lookup": counts.value / n,
    "hits_per_lookup": hits.value / n,
    "algorithmic_sectors_per_s": alg_sectors / (ms.value * 1e-3), "algorithmic_GBps_at_32B": alg_sectors * 32 / (ms.value * 1e-3) / 1e9,
    "gather_same_footprint_ms": gms.value, "gather_same_footprint_sectors_per_s": n / (gms.value * 1e-3),
    "gather_48GB_ms": gms_big.value, "gather_48GB_sectors_per_s": n / (gms_big.value * 1e-3),
}
out["frac_of_random_sector_peak_same_footprint"] = out["algorithmic_sectors_per_s"] / out["gather_same_footprint_sectors_per_s"]
out["frac_of_random_sector_peak_48GB"] = out["algorithmic_sectors_per_s"] / out["gather_48GB_sectors_per_s"]
print(json.dumps(out, indent=1))
