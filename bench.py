#!/usr/bin/env python
"""bench.py -- reads aligned + genotyped per second on the BASELINE config-2 workload.

Workload (BASELINE.json configs[1]): 1 Mb synthetic region, 10 000 SNP/indel sites, 1 sample, 30x paired 150-bp
reads (2*10^5 records), processed exactly as `graphtyper genotype --vcf` chops it: 20 regions of 50 kb (+1 kb pads),
each with its own graph + k-mer index.  One "step" = one pass of the hot path over all 2*10^5 records
(all 20 regions in ONE region-batched launch sequence: probe, chain, slow, score kernels).

  value      device-resident throughput: inputs already in HBM, kernels timed with CUDA events on the launching
             stream (library-side events, gtb_last_timing), max over ranks.
  e2e        the same records through the C-ABI call a user makes (gtb_submit_reads_multi + gtb_pool_finish) with
             HOST buffers: pinned staging + H2D + kernels + accumulator D2H inside the timed region.
  roofline   probe_kernel (the memory-bound index-probe kernel) against the measured HBM copy bandwidth (MEASURED_PEAKS.json) using the algorithmic
             6 468 B/read of SURVEY.md 8(d).
  cpu_baseline  the compiled reference (`oracle/_ref/bin/graphtyper genotype`, kind "reference") or the oracle
             port timed on this box's host cores on a bounded sample of the same workload.

`--impl reference` times the reference's own CPU implementation only (rank 0).
N > 1 (torchrun): every rank processes its own 30x read set of the same 1 Mb graph (weak scaling: reads sharded
by batch), then ONE NCCL all-reduce of the widened per-variant accumulators per region.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES_PER_READ = 6468  # 100 read + 388 probes x 16 + 72 labels + 48 graph bases + 40 accumulate
LENGTH = 1_000_000
N_SITES = 10_000
REGION = 50_000
# dram__bytes_read.sum + dram__bytes_write.sum of ONE probe_kernel launch on this workload, from the committed ncu capture
# profiles/r1c_ncu_summary.txt (223.9 MB + 13.2 MB; cold L2).  Well below the algorithmic 1.29 GB: "absent" -- the fate of 95+
# of the 96 Hamming-1 neighbours of a seed -- is answered from the presence filter in shared memory, and the tables are L2-resident.
PROBE_DRAM_BYTES_PER_LAUNCH = 237_096_704

def env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_workload(seed_offset: int = 0, length: int = LENGTH, n_sites: int = N_SITES):
    from graphtyper_b200 import abi, graph_build, synth
    ref = synth.make_reference(length, 11)
    sites = synth.make_sites(ref, n_sites, 12)
    gts = synth.make_genotypes(n_sites, 1, 13 + seed_offset)
    rs = synth.simulate_reads(ref, sites, gts[0], "SAMP1", 111 + seed_offset)
    regions = synth.split_regions(length, REGION)
    graphs, batches = [], []
    for (b, e) in regions:
        graphs.append(graph_build.build_region_graph(ref, sites, b, e))
        idx = synth.reads_for_region(rs, b, e)
        batches.append(abi.batch_from_readsets([rs], [idx]))
    return ref, sites, gts, rs, regions, graphs, batches


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs() -> tuple:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ reference arm
def write_reference_inputs(tmp: str, ref, sites, rs, regions, n_regions: int):
    from graphtyper_b200 import synth
    import oracle
    fa = os.path.join(tmp, "ref.fa")
    synth.write_fasta(fa, ref)
    vcf = os.path.join(tmp, "sites.vcf")
    synth.write_vcf(vcf, sites, "chr1", len(ref))
    subprocess.run([oracle.ref_binary("bgzip"), "-f", vcf], check=True)
    subprocess.run([oracle.ref_binary("tabix"), "-f", "-p", "vcf", vcf + ".gz"], check=True)
    jobs = []
    for (b, e) in regions[:n_regions]:
        idx = synth.reads_for_region(rs, b, e)
        sam = os.path.join(tmp, f"r{b}.sam")
        synth.write_sam(sam, rs.subset(idx), "chr1", len(ref))
        jobs.append((b, e, sam, len(idx)))
    return fa, vcf + ".gz", jobs


def run_reference_step(fa, vcfgz, jobs, tmp, threads: int) -> float:
    """One pass of `graphtyper genotype --vcf --no_bamshrink` over the sample regions, `threads` regions at a time
    (the reference cannot use more threads than samples inside one region: src/main.cpp:410-414)."""
    import oracle
    exe = oracle.ref_binary("graphtyper")
    t0 = time.perf_counter()
    running = []
    it = iter(jobs)
    k = 0
    while True:
        while len(running) < threads:
            j = next(it, None)
            if j is None:
                break
            b, e, sam, _ = j
            out = os.path.join(tmp, f"out{k}")
            k += 1
            env = dict(os.environ, TMPDIR=tmp)
            running.append(subprocess.Popen([exe, "genotype", fa, f"--sam={sam}", f"--region=chr1:{b}-{e}",
                                             f"--vcf={vcfgz}", "--no_bamshrink", "--threads=1", f"--output={out}"],
                                            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env))
        if not running:
            break
        running[0].wait()
        if running[0].returncode != 0:
            raise RuntimeError("reference graphtyper failed")
        running.pop(0)
    return time.perf_counter() - t0


def oracle_port_step(graphs, batches, n_regions: int) -> tuple:
    import oracle
    O = oracle.Oracle()
    t0 = time.perf_counter()
    n = 0
    for g, b in list(zip(graphs, batches))[:n_regions]:
        h = O.index_build(g)
        r = O.pool_run(g, h, 1, b, tap=False)
        O.result_free(r)
        O.index_free(h)
        n += len(b)
    return n, time.perf_counter() - t0


def cpu_baseline(ref, sites, rs, regions, graphs, batches, steps: int = 1, warmup: int = 0, sample_regions: int = 20):
    import oracle
    cores = os.cpu_count() or 1
    if oracle.ref_binary("graphtyper") and oracle.ref_binary("bgzip"):
        tmp = tempfile.mkdtemp(prefix="gtb_ref_")
        try:
            fa, vcfgz, jobs = write_reference_inputs(tmp, ref, sites, rs, regions, sample_regions)
            n_reads = sum(j[3] for j in jobs)
            threads = max(1, min(cores, len(jobs)))
            for _ in range(warmup):
                run_reference_step(fa, vcfgz, jobs, tmp, threads)
            times = [run_reference_step(fa, vcfgz, jobs, tmp, threads) for _ in range(max(1, steps))]
            t = float(np.mean(times))
            return {"value": n_reads / t, "unit": "reads/s", "cores": threads, "kind": "reference",
                    "sample": f"{len(jobs)} x 50 kb regions ({n_reads} records), `graphtyper genotype --vcf --no_bamshrink` "
                              f"CLI wall incl. graph+index build and VCF write, {threads} region processes in parallel",
                    "seconds_per_step": t, "n_reads": n_reads}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    n, t = oracle_port_step(graphs, batches, sample_regions)
    return {"value": n / t, "unit": "reads/s", "cores": 1, "kind": "port",
            "sample": f"{sample_regions} regions ({n} records) through oracle/gtb_oracle.cpp incl. index build",
            "seconds_per_step": t, "n_reads": n}


# ------------------------------------------------------------------------------------------------ main
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    metric = "reads/sec genotyped (150bp, 1Mb/10k-var graph)"
    config = {"workload": "configs[1]: 1 Mb synthetic region, 10k SNP/indel graph, 1 sample 30x 150 bp reads "
                          "(2e5 records, 20 x 50 kb regions, region-batched)",
              "reads_per_gpu": None, "regions": LENGTH // REGION, "l2": "flushed between timed iterations (256 MiB write)",
              "value_path": "device-resident replay, one launch sequence (prep, probe, chain, slow, huge, score)",
              "e2e_path": "gtb_pool_reset_multi + gtb_submit_reads_multi (4 concurrent chunks) + gtb_pool_finish_multi, pinned host buffers"}

    if args.impl == "reference":
        if rank != 0:
            return
        ref, sites, gts, rs, regions, graphs, batches = make_workload(0)
        cb = cpu_baseline(ref, sites, rs, regions, graphs, batches, steps=args.steps, warmup=min(args.warmup, 1))
        config["reads_per_gpu"] = cb["n_reads"]
        out = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "reads/s", "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_step"] * 1e3,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
               "config": config, "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
               "e2e": {"value": cb["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    from graphtyper_b200 import engine

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    ref, sites, gts, rs, regions, graphs, batches = make_workload(rank)
    n_reads = sum(len(b) for b in batches)
    config["reads_per_gpu"] = n_reads
    ctx = engine.Context(device=local_rank)
    # inputs live in page-locked host memory (gtb_host_alloc), as a production caller would fill them
    batches, pinned_arena = engine.pin_batches(batches)
    ids = list(range(len(graphs)))
    # region setup = graph upload + index build on the device (enumerate, radix sorts, table) for all 20 regions.
    # First call pays one-time CUDA module/pinned-pool initialisation, so the steady-state figure is the second call.
    setup_times = []
    for rep in range(5):
        if rep:
            for k in ids:
                ctx.region_end(k)
        t0 = time.perf_counter()
        ctx.region_begin_multi(ids, graphs)
        for k in ids:
            ctx.pool_begin(k, 1)
        setup_times.append(time.perf_counter() - t0)
    t_region = float(np.median(setup_times[1:]))
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.from_numpy(ctx.nccl_unique_id()))
        dist.broadcast(uid, 0)
        ctx.nccl_init(world, rank, uid.cpu().numpy())

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        flush.fill_(1)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reset():
        ctx.pool_reset_multi(ids)

    acc_bufs = [ctx.alloc_accumulators(k) for k in ids]

    h2d_bytes = sum(b.nbytes_h2d() for b in batches)

    # ---- e2e: host buffers through the public C-ABI call, H2D + kernels + accumulator D2H in the timed region
    def e2e_step():
        reset()
        st = ctx.submit_multi(ids, batches)
        if world > 1:
            ctx.allreduce_multi(ids)  # the single NCCL reduce of per-variant counts (all regions in one group)
        accs = ctx.pool_finish_multi(ids, out=acc_bufs)
        return st, accs

    for _ in range(max(3, args.warmup)):
        st, accs = e2e_step()
    d2h_bytes = sum(sum(v.nbytes for v in a.as_dict().values()) for a in accs)
    barrier()
    e2e_times = []
    for _ in range(args.steps):
        flush_l2()
        barrier()
        t = time.perf_counter()
        e2e_step()
        torch.cuda.synchronize()
        e2e_times.append(time.perf_counter() - t)
    e2e_t = float(np.mean(e2e_times))

    # ---- device-resident: replay the batch already in HBM, CUDA-event time of the kernels.  With nothing to copy there
    #      is nothing to overlap, so the batch is resident as ONE launch sequence (the e2e steps above use the library's
    #      automatic 2-chunk copy/compute pipeline).
    replay_chunks = env_int("GTB_BENCH_REPLAY_CHUNKS", 0)  # 0 = the library's automatic choice (4 concurrent chunks)
    ctx.set_chunks(replay_chunks)
    reset()
    ctx.submit_multi(ids, batches)
    ctx.set_chunks(0)
    for _ in range(max(3, args.warmup)):
        reset()
        ctx.replay()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev_ms, align_ms, score_ms, wall = [], [], [], []
    kt = {"prep_kernels": [], "probe_kernel": [], "chain_kernel": [], "slow_kernel": [], "score_kernel": []}
    n_slow = 0
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        flush_l2()
        reset()
        barrier()
        t = time.perf_counter()
        st = ctx.replay()
        wall.append(time.perf_counter() - t)
        _, a, s, d = ctx.last_timing()
        align_ms.append(a)
        score_ms.append(s)
        ev_ms.append(a + s)
        k = ctx.last_kernel_timing()
        for nm in kt:
            kt[nm].append(k[nm])
        n_slow = k["n_slow_tasks"]
        chain_t = ctx.last_chain_timing()
    barrier()
    t_all = time.perf_counter() - t_all0
    clocks = sampler.stop()
    step_ms = float(np.mean(ev_ms))
    if world > 1:
        tt = torch.tensor([step_ms, e2e_t], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_ms, e2e_t = float(tt[0]), float(tt[1])
        tot = torch.tensor([n_reads], device=dev, dtype=torch.float64)
        dist.all_reduce(tot)
        total_reads = float(tot[0])
    else:
        total_reads = float(n_reads)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        a_ms = float(np.mean(kt["probe_kernel"]))
        achieved = n_reads * ALGO_BYTES_PER_READ / (a_ms * 1e-3) / 1e9
        cb = None
        if not args.no_cpu_baseline and world == 1:
            try:
                cb = cpu_baseline(ref, sites, rs, regions, graphs, batches, steps=1, warmup=0)
            except Exception as ex:  # keep the GPU line even if the CPU arm breaks
                cb = {"value": None, "unit": "reads/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
        out = {
            "metric": metric, "value": total_reads / (step_ms * 1e-3), "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
            "e2e": {"value": total_reads / e2e_t, "unit": "reads/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_t * 1e3},
            "gpu_launches": int(st.kernel_launches) * args.steps,
            "kernels_ms": {**{nm: float(np.mean(v)) for nm, v in kt.items()},
                           "replay_call_wall_ms": float(np.mean(wall)) * 1e3, "slow_tasks": n_slow, "chain_tiers": chain_t},
            "region_setup_s": t_region,
            "e2e_incl_region_setup": {"value": total_reads / (e2e_t + t_region), "unit": "reads/s",
                                      "note": "index build + graph upload + H2D + kernels + D2H for the whole 1 Mb"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": PROBE_DRAM_BYTES_PER_LAUNCH, "kernel": "probe_kernel", "peak_source": peak_src,
                         "algorithmic_bytes_per_read": ALGO_BYTES_PER_READ},
            "clocks": clocks,
        }
        if cb is not None:
            out["cpu_baseline"] = {k: cb.get(k) for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
