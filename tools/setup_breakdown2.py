import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.cuda.set_device(0)
import bench
from graphtyper_b200 import engine
ref, sites, gts, rs, regions, graphs, batches = bench.make_workload(0)
ctx = engine.Context(0)
batches, arena = engine.pin_batches(batches)
ids = list(range(len(graphs)))
for rep in range(4):
    t0 = time.perf_counter(); ctx.region_begin_multi(ids, graphs); t1 = time.perf_counter()
    for k in ids: ctx.pool_begin(k, 1)
    t2 = time.perf_counter()
    for k in ids: ctx.region_end(k)
    t3 = time.perf_counter()
    print(f"rep {rep}: region_begin_multi {1e3*(t1-t0):.1f} ms  pool_begin x20 {1e3*(t2-t1):.1f} ms  region_end x20 {1e3*(t3-t2):.1f} ms")
