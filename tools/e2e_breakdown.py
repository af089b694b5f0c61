"""Host-side breakdown of one end-to-end step (run on the GPU box): where the milliseconds outside the kernels go."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from graphtyper_b200 import engine

ref, sites, gts, rs, regions, graphs, batches = bench.make_workload(0)
ctx = engine.Context(0)
batches, _arena = engine.pin_batches(batches)
ids = list(range(len(graphs)))
for k, g in zip(ids, graphs):
    ctx.region_begin(k, g); ctx.pool_begin(k, 1)
bufs = [ctx.alloc_accumulators(k) for k in ids]
T = {"reset": [], "submit": [], "finish": [], "h2d_ev": [], "kern_ev": []}
for it in range(12):
    t0 = time.perf_counter(); ctx.pool_reset_multi(ids)
    t1 = time.perf_counter(); ctx.submit_multi(ids, batches)
    t2 = time.perf_counter(); ctx.pool_finish_multi(ids, out=bufs)
    t3 = time.perf_counter()
    h2d, al, sc, d2h = ctx.last_timing()
    if it >= 2:
        T["reset"].append(t1 - t0); T["submit"].append(t2 - t1); T["finish"].append(t3 - t2)
        T["h2d_ev"].append(h2d / 1e3); T["kern_ev"].append((al + sc) / 1e3)
for k, v in T.items():
    print(f"{k:10s} {np.mean(v)*1e3:8.3f} ms")
print("submit - h2d - kernels (host staging + launch + sync) =",
      (np.mean(T["submit"]) - np.mean(T["h2d_ev"]) - np.mean(T["kern_ev"])) * 1e3, "ms")
