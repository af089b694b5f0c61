"""Throughput of the bench workload when several host threads drive one context each on the same GPU (profiling aid).

Each thread replays (device-resident) or submits (host buffers, pinned) the whole 2e5-record step on its OWN context; steps
are issued back to back, so one step's tail (general / slow tiers, second score pass, D2H) overlaps the next step's front.
Prints steps/s and reads/s for 1..N threads."""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from graphtyper_b200 import engine

ref, sites, gts, rs, regions, graphs, batches = bench.make_workload(0)
n_reads = sum(len(b) for b in batches)
ids = list(range(len(graphs)))
max_threads = int(os.environ.get("PP_THREADS", 3))
steps = int(os.environ.get("PP_STEPS", 20))
batches, arena = engine.pin_batches(batches)
ctxs = []
for t in range(max_threads):
    c = engine.Context(0)
    c.region_begin_multi(ids, graphs)
    for k in ids:
        c.pool_begin(k, 1)
    c.set_chunks(int(os.environ.get("PP_CHUNKS", 1)))
    c.submit_multi(ids, batches)
    ctxs.append(c)
acc = [[c.alloc_accumulators(k) for k in ids] for c in ctxs]


def run(nt, mode):
    bar = threading.Barrier(nt + 1)

    def work(t):
        c = ctxs[t]
        bar.wait()
        for _ in range(steps):
            c.pool_reset_multi(ids)
            if mode == "replay":
                c.replay()
            elif mode == "replay+finish":  # kernels + D2H, no H2D
                c.replay()
                c.pool_finish_multi(ids, out=acc[t])
            elif mode == "submit":         # H2D + kernels, no D2H
                c.submit_multi(ids, batches)
            else:
                c.submit_multi(ids, batches)
                c.pool_finish_multi(ids, out=acc[t])
        bar.wait()

    th = [threading.Thread(target=work, args=(t,)) for t in range(nt)]
    for x in th:
        x.start()
    bar.wait()
    t0 = time.perf_counter()
    bar.wait()
    dt = time.perf_counter() - t0
    for x in th:
        x.join()
    return dt / (steps * nt)


only = int(os.environ.get("PP_ONLY", 0))
for mode in os.environ.get("PP_MODES", "replay,e2e").split(","):
    for nt in ([only] if only else range(1, max_threads + 1)):
        run(nt, mode)
        ms = run(nt, mode) * 1e3
        print(f"{mode:6s} threads {nt}: {ms:.3f} ms/step  {n_reads / ms / 1e3:.1f} M reads/s")
