import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from graphtyper_b200 import engine
ref, sites, gts, rs, regions, graphs, batches = bench.make_workload(0)
ctx = engine.Context(0)
ids = list(range(len(graphs)))
for k, g in zip(ids, graphs):
    ctx.region_begin(k, g); ctx.pool_begin(k, 1)
ctx.submit_multi(ids, batches)
out = (C.c_uint64 * 24)()
ctx.lib.gtb_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
ctx.lib.gtb_debug_counters(ctx.h, out)
names = ["refs", "vars", "paths", "locs", "labels", "cand_vars", "cands", "keys", "tap", "pool", "read_len", "probe_flag"]
print("fast->slow reasons:", {n: int(out[i]) for i, n in enumerate(names) if out[i]})
print("slow overflow:", {n: int(out[12 + i]) for i, n in enumerate(names) if out[12 + i]})
print(ctx.last_kernel_timing())
