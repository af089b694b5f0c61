#!/usr/bin/env python
"""Throughput of the discovery re-alignment kernel (gtb_sw_align_batch) next to the compiled paw on the host cores.

  python tools/sw_bench.py [--pairs 200000] [--reps 5]

Pairs are shaped like realign_to_indels' input (reads <= 151 bp against 400..520 bp haplotype windows).  Prints one
JSON line: device-resident kernel rate (CUDA events, replay on resident inputs), end-to-end rate through the C ABI with
host buffers, cell-update rate, and the CPU baseline (paw_probe --time on all host threads, bounded sample)."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphtyper_b200 import engine, synth  # noqa: E402


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=200000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--min-query", type=int, default=20, help="shortest read (151 = every read full length)")
    a = ap.parse_args()
    base = 20000
    q0, d0 = synth.make_sw_pairs(base, seed=3, min_db=400, max_db=520, min_query=a.min_query)
    rep = (a.pairs + base - 1) // base
    q, d = (q0 * rep)[:a.pairs], (d0 * rep)[:a.pairs]
    qb, qo = engine.pack_sequences(q)
    db, do = engine.pack_sequences(d)
    cells = float(np.sum(np.diff(qo).astype(np.int64) * np.diff(do).astype(np.int64)))
    ctx = engine.Context(0)
    ctx.sw_align_packed(qb, qo, db, do)  # warm-up (allocations)
    e2e = []
    for _ in range(a.reps):
        t0 = time.perf_counter()
        out = ctx.sw_align_packed(qb, qo, db, do)
        e2e.append(time.perf_counter() - t0)
    tim = ctx.sw_last_timing()
    ker = []
    for _ in range(a.reps):
        ctx.sw_replay()
        ker.append(ctx.sw_last_timing()["kernel_ms"])
    k_ms = float(np.median(ker))
    line = {"metric": "sw_pairs_per_s", "pairs": a.pairs, "mean_query": float(np.diff(qo).mean()),
            "mean_window": float(np.diff(do).mean()), "kernel_ms": k_ms, "value": a.pairs / (k_ms * 1e-3),
            "gcups": cells / (k_ms * 1e-3) / 1e9, "e2e": a.pairs / float(np.median(e2e)),
            "h2d_ms": tim["h2d_ms"], "d2h_ms": tim["d2h_ms"], "checksum": int(out[:, 0].astype(np.int64).sum())}
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "paw_probe")
    if os.path.exists(exe) and a.cpu_sample > 0:
        n = min(a.cpu_sample, a.pairs)
        with tempfile.NamedTemporaryFile("wb", suffix=".tsv", delete=False) as f:
            f.write(b"".join(x + b"\t" + y + b"\n" for x, y in zip(q[:n], d[:n])))
        cores = os.cpu_count() or 1
        try:
            txt = subprocess.run([exe, f.name, "--time", str(cores)], capture_output=True, text=True, check=True).stdout
            tok = txt.split()
            line["cpu_baseline"] = {"value": n / float(tok[3]), "unit": "pairs/s", "cores": cores, "kind": "reference",
                                    "sample": f"{n} pairs, compiled paw (AVX-512 dispatch), {cores} threads"}
        finally:
            os.unlink(f.name)
    print(json.dumps(line))


if __name__ == "__main__":
    main()
