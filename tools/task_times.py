"""Per-task timeline of chain_general_kernel (profiling aid): GTB_TASK_TIMES=<file> makes the library dump start/end globaltimer
values of every chain_general_kernel task of chunk 0; this script runs the bench workload once with 1 chunk and summarises."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = "/tmp/task_times.bin"
os.environ["GTB_TASK_TIMES"] = out
import bench
from graphtyper_b200 import engine
ref, sites, gts, rs, regions, graphs, batches = bench.make_workload(0)
ctx = engine.Context(0)
ids = list(range(len(graphs)))
ctx.region_begin_multi(ids, graphs)
for k in ids:
    ctx.pool_begin(k, 1)
ctx.set_chunks(1)
for _ in range(3):
    ctx.pool_reset_multi(ids)
    ctx.submit_multi(ids, batches)
print(ctx.last_kernel_timing())
raw = open(out, "rb").read()
n = int(np.frombuffer(raw[:8], np.uint64)[0])
tt = np.frombuffer(raw[8:8 + 16 * n], np.uint64).reshape(n, 2).astype(np.int64)
t0 = tt[:, 0].min()
start, end = (tt[:, 0] - t0) / 1e3, (tt[:, 1] - t0) / 1e3
dur = end - start
print(f"tasks {n}; kernel span {end.max():.1f} us; task duration us: mean {dur.mean():.1f} median {np.median(dur):.1f} "
      f"p90 {np.percentile(dur, 90):.1f} p99 {np.percentile(dur, 99):.1f} max {dur.max():.1f}")
print("start time us: p50 %.1f p90 %.1f max %.1f" % (np.median(start), np.percentile(start, 90), start.max()))
lanes = int(os.environ.get("GTB_GEN_LANES", 8))
w = n // lanes
wd = dur[:w * lanes].reshape(w, lanes)
print("per warp: max-task mean %.1f, mean-task mean %.1f; warp end p50 %.1f p90 %.1f p99 %.1f max %.1f" % (
    wd.max(1).mean(), wd.mean(1).mean(), *np.percentile(end[:w * lanes].reshape(w, lanes).max(1), [50, 90, 99, 100])))
hist, edges = np.histogram(end, bins=12)
print("end-time histogram:", list(zip(np.round(edges[:-1]).astype(int), hist)))
blk = n // 128
bs = start[:blk * 128].reshape(blk, 128).min(1)
print("block start us: p50 %.1f p75 %.1f p90 %.1f max %.1f" % tuple(np.percentile(bs, [50, 75, 90, 100])))
