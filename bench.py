#!/usr/bin/env python
"""bench.py -- reads aligned + genotyped per second on the BASELINE config-2 workload.

Workload (BASELINE.json configs[1]): 1 Mb synthetic region, 10 000 SNP/indel sites, 1 sample, 30x paired 150-bp
reads (2*10^5 records), processed exactly as `graphtyper genotype --vcf` chops it: 20 regions of 50 kb (+1 kb pads),
each with its own graph + k-mer index.  One "step" = one pass of the hot path over all 2*10^5 records of one sample
(all 20 regions in ONE region-batched launch sequence: prep, probe, chain, chain_general, slow, huge, score kernels).

Steps are issued the way the reference issues pools: one pool per worker thread (paw::Station, caller.cpp:272-391), all
threads sharing one index.  Here: GTB_BENCH_THREADS (default 4) host threads, one context each, the regions shared through
gtb_region_attach, every context with its own sample (read set); the K timed steps are dealt round-robin to the threads
and run back to back, so one step's tail (general / slow tiers, second score pass, D2H) overlaps the next step's front.

  value      whole-job device-resident throughput: every context's batch already in HBM, K steps (reset + replay of
             the launch sequence) bracketed by barrier + synchronize, CUDA events around the bracket, max over ranks.
  e2e        the same K steps through the C-ABI call a user makes (gtb_pool_reset_multi + gtb_submit_reads_multi +
             gtb_pool_finish_multi) with HOST buffers: pinned staging + H2D + kernels + accumulator D2H in the timed region.
  latency    ONE step alone (L2 flushed before it): device time and per-kernel CUDA-event times -- what the roofline
             table is computed from.
  roofline   step level, as SURVEY.md 8(d) defines it: reads/s x 6 468 algorithmic B/read over the measured HBM copy
             bandwidth (MEASURED_PEAKS.json); plus the dominant kernel on its own bytes and the per-kernel table.
  parity_checked  (N = 1) records of the VCFs the reference CLI wrote for the 20 regions whose GT:AD:MD:DP:GQ:PL column
             equals what the accumulators of the GPU run give (gtb_calls_from_accumulators) -- all of them, or the run fails.
  cpu_baseline  the compiled reference (`oracle/_ref/bin/graphtyper genotype`, kind "reference") timed on this box's
             host cores: (a) CLI wall and (b) the align + accumulate span from its own --vverbose timestamps.

`--impl reference` times the reference's own CPU implementation only (rank 0).
N > 1 (torchrun): sample-sharded weak scaling (SURVEY.md 8e primary axis) -- every rank genotypes its own samples of the same
1 Mb graph, accumulators and calls stay local, and the job ends with ONE NCCL reduce of the per-variant summaries
(VarStats::add_stats: sum + max, gtb_allreduce_varstats).  Before the timed region a verification step shards ONE sample's
reads over the ranks, all-reduces the accumulators and compares them with the single-GPU result (sharded_parity_checked).
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import re
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
# per kernel (same table): probe = read in + probes + label fetch, chain = tail extension + path records, score = accumulate
ALGO_BYTES_PER_KERNEL = {"probe_kernel": 100 + 6208 + 72, "chain_kernel": 48 + 72, "score_kernel": 40}
LENGTH = 1_000_000
N_SITES = 10_000
REGION = 50_000
KERNEL_TABLE = os.path.join(ROOT, "profiles", "r2_kernel_table.json")  # ncu figures per kernel (committed capture)


def env_int(name: str, default: int) -> int:
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_workload(seed_offset: int = 0, length: int = LENGTH, n_sites: int = N_SITES, with_graphs: bool = True):
    from graphtyper_b200 import abi, graph_build, synth
    ref = synth.make_reference(length, 11)
    sites = synth.make_sites(ref, n_sites, 12)
    gts = synth.make_genotypes(n_sites, 1, 13 + seed_offset)
    rs = synth.simulate_reads(ref, sites, gts[0], "SAMP1", 111 + seed_offset)
    regions = synth.split_regions(length, REGION)
    graphs, batches = [], []
    for (b, e) in regions:
        if with_graphs:
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50",
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


_TS = re.compile(r"^\[(\d+)-(\d+)-(\d+) (\d+):(\d+):(\d+)\.(\d+)\] <\w+> (\S+?):(\d+) ")


def align_span_seconds(log_path: str):
    """Span (b) of SURVEY.md 8(d) from the reference's own --vverbose log: `Got N haplotypes` (vcf_writer.cpp:70) ->
    `Num of duplicated records` (hts_parallel_reader.cpp:711) = read + align + accumulate of the pool."""
    t_begin = t_end = None
    with open(log_path) as f:
        for line in f:
            m = _TS.match(line)
            if not m:
                continue
            t = int(m.group(4)) * 3600 + int(m.group(5)) * 60 + int(m.group(6)) + int(m.group(7)) / 1000.0
            if m.group(8) == "vcf_writer.cpp" and m.group(9) == "70":
                t_begin = t
            elif m.group(8) == "hts_parallel_reader.cpp" and m.group(9) == "711":
                t_end = t
    if t_begin is None or t_end is None:
        return None
    return (t_end - t_begin) % 86400.0


def run_reference_step(fa, vcfgz, jobs, tmp, threads: int, keep: str = None):
    """One pass of `graphtyper genotype --vcf --no_bamshrink` over the sample regions, `threads` regions at a time
    (the reference cannot use more threads than samples inside one region: src/main.cpp:410-414).  Returns (wall seconds,
    summed align+accumulate span seconds or None); with `keep`, the per-region VCFs are copied there as r<begin>.vcf.gz."""
    import oracle
    exe = oracle.ref_binary("graphtyper")
    t0 = time.perf_counter()
    running = []
    it = iter(jobs)
    k = 0
    done = []
    while True:
        while len(running) < threads:
            j = next(it, None)
            if j is None:
                break
            b, e, sam, _ = j
            out = os.path.join(tmp, f"out{k}")
            log = os.path.join(tmp, f"log{k}.txt")
            k += 1
            env = dict(os.environ, TMPDIR=tmp)
            p = subprocess.Popen([exe, "genotype", fa, f"--sam={sam}", f"--region=chr1:{b}-{e}", f"--vcf={vcfgz}",
                                  "--no_bamshrink", "--threads=1", f"--output={out}", "--vverbose", f"--log={log}"],
                                 stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
            running.append((p, b, out, log))
        if not running:
            break
        p, b, out, log = running.pop(0)
        p.wait()
        if p.returncode != 0:
            raise RuntimeError("reference graphtyper failed")
        done.append((b, out, log))
    wall = time.perf_counter() - t0
    spans = [align_span_seconds(log) for _, _, log in done if os.path.exists(log)]
    span = float(sum(spans)) if spans and all(s is not None for s in spans) else None
    if keep:
        os.makedirs(keep, exist_ok=True)
        for b, out, _ in done:
            for v in glob.glob(os.path.join(out, "chr1", "*.vcf.gz")):
                shutil.copy(v, os.path.join(keep, f"r{b}.vcf.gz"))
    for _, out, _ in done:
        shutil.rmtree(out, ignore_errors=True)
    return wall, span


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


def cpu_baseline(ref, sites, rs, regions, graphs, batches, steps: int = 1, warmup: int = 0, sample_regions: int = 20,
                 keep_vcfs: str = None):
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
            res = [run_reference_step(fa, vcfgz, jobs, tmp, threads, keep=keep_vcfs if i == 0 else None)
                   for i in range(max(1, steps))]
            t = float(np.mean([r[0] for r in res]))
            spans = [r[1] for r in res if r[1] is not None]
            out = {"value": n_reads / t, "unit": "reads/s", "cores": threads, "kind": "reference",
                   "sample": f"{len(jobs)} x 50 kb regions ({n_reads} records), `graphtyper genotype --vcf --no_bamshrink` "
                             f"CLI wall incl. process start, SAM parse, graph + index build and VCF write, {threads} region "
                             f"processes in parallel",
                   "seconds_per_step": t, "n_reads": n_reads}
            if spans:
                s = float(np.mean(spans))  # core-seconds: the regions' spans summed
                out["align_accumulate_span"] = {
                    "core_seconds_per_step": s, "reads_per_s_per_core": n_reads / s,
                    "reads_per_s_all_cores": n_reads / s * threads,
                    "note": "span (b) of SURVEY 8(d): read + align + accumulate only, from the reference's own --vverbose "
                            "timestamps (vcf_writer.cpp:70 -> hts_parallel_reader.cpp:711), summed over the regions; the "
                            "all-cores figure assumes the region processes scale perfectly"}
            return out
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    n, t = oracle_port_step(graphs, batches, sample_regions)
    return {"value": n / t, "unit": "reads/s", "cores": 1, "kind": "port",
            "sample": f"{sample_regions} regions ({n} records) through oracle/gtb_oracle.cpp incl. index build",
            "seconds_per_step": t, "n_reads": n}


# ------------------------------------------------------------------------------------------------ helpers of the GPU arm
def check_against_reference_vcfs(ctx, accs, vcf_dir, regions, ref, sites):
    """GT:AD:MD:DP:GQ:PL of every record the reference CLI wrote for the 20 regions against the calls the accumulators give."""
    import oracle
    from graphtyper_b200 import vcf_check
    binned = oracle.binned_pl()
    if binned is None:
        return None
    n_total, bad_total, n_regions = 0, [], 0
    for (b, e), acc in zip(regions, accs):
        path = os.path.join(vcf_dir, f"r{b}.vcf.gz")
        if not os.path.exists(path):
            continue
        ph, gt, gq = ctx.calls(acc)
        n, bad = vcf_check.check_region(path, acc, ph, gt, gq, binned, ref, sites)
        n_total += n
        n_regions += 1
        bad_total += [f"region {b}: {x}" for x in bad]
    return {"records": n_total, "regions": n_regions, "mismatches": len(bad_total), "first_mismatches": bad_total[:5]}


def sw_key(ctx):
    """Discovery re-alignment kernel (N1): pairs/s, GCUPS and the share of the integer-issue peak, next to compiled paw."""
    from graphtyper_b200 import engine, synth
    n_pairs = 100000
    q0, d0 = synth.make_sw_pairs(20000, seed=3, min_db=400, max_db=520, min_query=140)  # reads of 140-151 bp, as realign_to_indels sees
    q, d = (q0 * 5)[:n_pairs], (d0 * 5)[:n_pairs]
    qb, qo = engine.pack_sequences(q)
    db, do = engine.pack_sequences(d)
    cells = float(np.sum(np.diff(qo).astype(np.int64) * np.diff(do).astype(np.int64)))
    ctx.sw_align_packed(qb, qo, db, do)
    e2e = []
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.sw_align_packed(qb, qo, db, do)
        e2e.append(time.perf_counter() - t0)
    ker = []
    for _ in range(5):
        ctx.sw_replay()
        ker.append(ctx.sw_last_timing()["kernel_ms"])
    k_ms = float(np.median(ker))
    cups = cells / (k_ms * 1e-3)
    table = json.load(open(KERNEL_TABLE)) if os.path.exists(KERNEL_TABLE) else {}
    tipc = table.get("sw_kernel", {}).get("thread_inst_per_cell")
    out = {"pairs": n_pairs, "kernel_ms": k_ms, "pairs_per_s": n_pairs / (k_ms * 1e-3), "gcups": cups / 1e9,
           "e2e_pairs_per_s": n_pairs / float(np.median(e2e)),
           "int_issue_peak_lane_ops_per_s": 148 * 128 * 1.965e9,
           "thread_inst_per_cell": tipc,
           "frac_of_int_issue_peak": (cups * tipc / (148 * 128 * 1.965e9)) if tipc else None,
           "note": "one warp per (read, window) pair, DPX VIADDMNMX cells; thread_inst_per_cell from the committed ncu capture"}
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "paw_probe")
    if os.path.exists(exe):
        n = 10000
        with tempfile.NamedTemporaryFile("wb", suffix=".tsv", delete=False) as f:
            f.write(b"".join(x + b"\t" + y + b"\n" for x, y in zip(q[:n], d[:n])))
        cores = os.cpu_count() or 1
        try:
            tok = subprocess.run([exe, f.name, "--time", str(cores)], capture_output=True, text=True, check=True).stdout.split()
            out["cpu_baseline"] = {"value": n / float(tok[3]), "unit": "pairs/s", "cores": cores, "kind": "reference",
                                   "sample": f"{n} pairs, compiled paw (AVX-512 dispatch), {cores} threads"}
        except Exception as ex:
            out["cpu_baseline"] = {"kind": "unavailable", "sample": repr(ex)}
        finally:
            os.unlink(f.name)
    return out


def bgzf_key(owner, ids, graphs, regions, rs, n_regions=6, n_threads=4):
    """N3: the BGZF entry (gtb_submit_bgzf: compressed BAM bytes in, genotyped pool out) on the first regions of the workload,
    one pool per call as the drop-in reader issues them, several pool threads in flight; next to the record entry on the same
    records (same accumulators required) and to zlib inflating the same bytes on one host core."""
    from graphtyper_b200 import abi, bgzf, engine, synth
    ids = ids[:n_regions]
    bams = [abi.bam_batch_from_readsets([rs], [synth.reads_for_region(rs, b, e)]) for (b, e) in regions[:n_regions]]
    n = sum(len(b) for b in bams)
    hdr = bgzf.bam_header([("chr1", 250000000)])
    pools, queries, comp_bytes, raws = [], [], 0, []
    for bam in bams:
        raw, blocks = bgzf.bgzf_compress_records(hdr, bgzf.records_from_batch(bam))  # blocks cut at record boundaries, as htslib writes
        u = bgzf.voffset_of(blocks, len(hdr))
        pools.append(bgzf.HostBgzfFiles([(raw, [(u, len(raw) << 16, True)], 0, 0)], [[b[0] for b in blocks] + [len(raw) - 28, len(raw)]]))
        pos = bam.core["pos"].astype(np.int64)
        queries.append(bgzf.query(int(bam.core["tid"][0]), int(pos.min()), int(pos.max()) + 1))
        comp_bytes += len(raw)
        raws.append(raw)
    t0 = time.perf_counter()
    inflated = sum(len(bgzf.inflate_file(r)) for r in raws)
    t_zlib = time.perf_counter() - t0
    ctxs = []
    for t in range(n_threads):
        c = engine.Context(device=owner.device)
        for k in ids:
            c.region_attach(k, owner, k)
            c.pool_begin(k, 1)
        ctxs.append(c)
    try:
        c0 = ctxs[0]
        want = []
        for k in ids:  # the record entry on the same records: the accumulators the BGZF entry must reproduce
            c0.submit_bam(k, bams[k])
            want.append({a: b.copy() for a, b in c0.pool_finish(k).as_dict().items()})
            c0.pool_reset(k)
        lat = []
        same = True
        for k in ids:
            t0 = time.perf_counter()
            c0.submit_bgzf(k, pools[k], queries[k])
            lat.append(time.perf_counter() - t0)
            got = c0.pool_finish(k).as_dict()
            same = same and all(np.array_equal(got[a], want[k][a]) for a in want[k])
            c0.pool_reset(k)
        stitched = c0.debug_bgzf_stitched()

        def pool(t, s):
            k = s % len(ids)
            ctxs[t].pool_reset(k)
            ctxs[t].submit_bgzf(k, pools[k], queries[k])

        reps = 4 * len(ids)
        run_pipelined(n_threads, reps, pool)
        per = run_pipelined(n_threads, reps, pool) / reps
    finally:
        for c in ctxs:
            c.close()
    return {"regions": len(ids), "records": n, "compressed_bytes": comp_bytes, "inflated_bytes": inflated,
            "accumulators_equal_record_entry": bool(same), "files_with_parallel_record_walk": stitched,
            "ms_per_pool_one_at_a_time": float(np.median(lat)) * 1e3, "pool_threads": n_threads, "ms_per_pool_pooled": per * 1e3,
            "reads_per_s": n / len(ids) / per, "compressed_GBps": comp_bytes / len(ids) / per / 1e9,
            "cpu_baseline": {"kind": "port", "cores": 1, "value": inflated / t_zlib / 1e9, "unit": "GB/s inflated",
                             "sample": "zlib inflate of the same BGZF bytes alone (no record parsing, no merge), one host core"},
            "gpu_inflated_GBps": inflated / len(ids) / per / 1e9,
            "note": "one warp per BGZF block (a serial Huffman decode per block: one pool alone waits ~2.4 ms for its slowest "
                    "block, pools in flight overlap); compressed bytes -> accumulators, nothing but 32 status bytes comes back "
                    "before the pool is finished"}


def cli_end_to_end(ref, sites, rs, regions):
    """end-to-end CLI wall.  ONE process genotypes all 20 regions of the sample from one indexed BAM
    (`graphtyper genotype REF --sam=all.bam --region_file=... --vcf=...`, bamshrink included) -- once with the stock
    binary, once with the drop-in binary oracle/_ref/bin/graphtyper_gtb (the same reference with index_graph and
    parallel_reader_genotype_only replaced by libgtb200, integration/gtb_pool_reader.cpp); the VCFs must be identical."""
    import gzip
    import oracle
    from graphtyper_b200 import synth
    exe = {k: oracle.ref_binary(k) for k in ("graphtyper", "graphtyper_gtb", "bgzip", "tabix", "sam2bam")}
    if not all(exe.values()):
        return {"unavailable": "needs " + ", ".join(k for k, v in exe.items() if not v) + " under oracle/_ref/bin"}
    tmp = tempfile.mkdtemp(prefix="gtb_cli_")
    try:
        fa = os.path.join(tmp, "ref.fa")
        synth.write_fasta(fa, ref)
        vcf = os.path.join(tmp, "sites.vcf")
        synth.write_vcf(vcf, sites, "chr1", len(ref))
        subprocess.run([exe["bgzip"], "-f", vcf], check=True)
        subprocess.run([exe["tabix"], "-f", "-p", "vcf", vcf + ".gz"], check=True)
        sam = os.path.join(tmp, "all.sam")
        synth.write_sam(sam, rs, "chr1", len(ref))
        bam = os.path.join(tmp, "all.bam")
        subprocess.run([exe["sam2bam"], sam, bam], check=True)
        os.unlink(sam)
        rf = os.path.join(tmp, "regions.txt")
        with open(rf, "w") as f:
            f.write("".join(f"chr1:{b}-{e}\n" for b, e in regions))
        res, text = {}, {}
        for name in ("graphtyper", "graphtyper_gtb", "graphtyper_gtb+files"):
            times = []
            for rep in range(4):  # later runs have the page cache and (GPU) the driver warm
                out = os.path.join(tmp, f"out_{name}")
                shutil.rmtree(out, ignore_errors=True)
                t0 = time.perf_counter()
                env = dict(os.environ, TMPDIR=tmp)
                if name.endswith("+files"):
                    env["GTB200_VCF_FILES"] = "1"  # pool results through the reference's cereal + gzip files again
                log = os.path.join(tmp, f"log_{name}.txt")
                if os.path.exists(log):
                    os.unlink(log)
                r = subprocess.run([exe[name.split("+")[0]], "genotype", fa, f"--sam={bam}", f"--region_file={rf}",
                                    f"--vcf={vcf}.gz", "--threads=1", f"--output={out}", "--verbose", f"--log={log}"],
                                   capture_output=True, text=True, env=env)
                times.append(time.perf_counter() - t0)
                if r.returncode != 0:
                    return {"error": f"{name} failed: {r.stderr[-400:]}"}
            res[name] = min(times)
            res[name + " median"] = float(np.median(times))
            # steady state per region from the tool's own log: time between consecutive "Finished! Output written" lines
            stamps = []
            with open(os.path.join(tmp, f"log_{name}.txt")) as f:  # (the log of the last run)
                for line in f:
                    m = _TS.match(line.replace("<info> ", "<info> x:0 ", 1)) if "Finished! Output written" in line else None
                    if m:
                        stamps.append(int(m.group(4)) * 3600 + int(m.group(5)) * 60 + int(m.group(6)) + int(m.group(7)) / 1000.0)
            if len(stamps) > 2:
                res[name + " per_region"] = float(np.median(np.diff(stamps)))
            text[name] = {}
            for v in sorted(glob.glob(os.path.join(tmp, f"out_{name}", "chr1", "*.vcf.gz"))):
                with gzip.open(v, "rt") as fh:
                    text[name][os.path.basename(v)] = fh.read()
        same = text["graphtyper"] == text["graphtyper_gtb"] and len(text["graphtyper"]) == len(regions)
        n_rec = sum(sum(1 for ln in t.splitlines() if ln and not ln.startswith("#")) for t in text["graphtyper"].values())
        same = same and text["graphtyper_gtb+files"] == text["graphtyper"]
        return {"reference_cli_s": res["graphtyper"], "dropin_cli_s": res["graphtyper_gtb"],
                "dropin_cli_s_pool_results_through_files": res["graphtyper_gtb+files"],
                "medians_s": {k: v for k, v in res.items() if k.endswith("median")},
                "steady_state_s_per_region": {k.split(" ")[0]: v for k, v in res.items() if k.endswith("per_region")},
                "speedup": res["graphtyper"] / res["graphtyper_gtb"], "vcf_files": len(text["graphtyper"]),
                "vcf_records": n_rec, "vcfs_identical": bool(same), "reads": int(len(rs)),
                "note": "one process, --threads=1 (one sample), 20 regions from one indexed BAM incl. bamshrink, graph "
                        "construction, VCF merge + BGZF; best of 4 runs each; the drop-in runs include ~0.6 s of CUDA context creation, steady_state_s_per_region (median time between the regions' 'Finished!' log lines) does not"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_pipelined(n_threads: int, steps: int, fn, at_end=None):
    """steps calls of fn(thread, step) dealt round-robin to n_threads host threads (then at_end(thread), if given); returns
    wall seconds from the common start to the last thread's end (the callers bracket it with barrier + synchronize and CUDA
    events)."""
    bar = threading.Barrier(n_threads + 1)
    errors = []

    def work(t):
        try:
            bar.wait()
            for s in range(t, steps, n_threads):
                fn(t, s)
            if at_end is not None:
                at_end(t)
        except Exception as ex:
            errors.append(ex)
        finally:
            bar.wait()

    th = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    for x in th:
        x.start()
    bar.wait()
    t0 = time.perf_counter()
    bar.wait()
    dt = time.perf_counter() - t0
    for x in th:
        x.join()
    if errors:
        raise errors[0]
    return dt


def flat_accumulators(accs) -> np.ndarray:
    return np.concatenate([np.concatenate([v.astype(np.int64).ravel() for k, v in sorted(a.as_dict().items())
                                           if k != "saturated"]) for a in accs])


# ------------------------------------------------------------------------------------------------ main
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    metric = "reads/sec genotyped (150bp, 1Mb/10k-var graph)"
    config = {"workload": "configs[1]: 1 Mb synthetic region, 10k SNP/indel graph, 1 sample 30x 150 bp reads "
                          "(2e5 records per step, 20 x 50 kb regions, region-batched)",
              "reads_per_step": None, "regions": LENGTH // REGION}

    if args.impl == "reference":
        if rank != 0:
            return
        ref, sites, gts, rs, regions, graphs, batches = make_workload(0)
        cb = cpu_baseline(ref, sites, rs, regions, graphs, batches, steps=args.steps, warmup=min(args.warmup, 1))
        config["reads_per_step"] = cb["n_reads"]
        config["path"] = ("oracle/_ref/bin/graphtyper (the unmodified reference compiled here), one `genotype --vcf` process per "
                          "50 kb region, as many in parallel as host cores")
        base = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "align_accumulate_span") if k in cb}
        out = {"impl": "reference", "metric": metric, "value": cb["value"], "unit": "reads/s", "n_gpus": args.gpus,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["seconds_per_step"] * 1e3,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
               "config": config, "cpu_baseline": base,
               "e2e": {"value": cb["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    from graphtyper_b200 import abi, engine

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    host_cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if world > 1 and hasattr(os, "sched_setaffinity") and not os.environ.get("GTB_BENCH_NO_PIN"):
        # one slice of the node's cores per rank: the ranks' pool and staging threads stop migrating over each other
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        os.environ["GTB_HOST_THREADS"] = str(len(mine))
    # pool threads per rank: 4, fewer when the ranks of the node have to share the host cores
    n_threads = max(1, env_int("GTB_BENCH_THREADS", max(2, min(4, host_cores // max(1, world) - 1))))
    steps, warmup = max(1, args.steps), max(3, args.warmup)

    # ---- one read set (sample) per context; the graph is the same for all
    ref, sites, gts, rs, regions, graphs, batches0 = make_workload(rank * 16)
    workloads = [batches0] + [make_workload(rank * 16 + t, with_graphs=False)[6] for t in range(1, n_threads)]
    n_reads = sum(len(b) for b in batches0)
    n_reads_ctx = [sum(len(b) for b in w) for w in workloads]
    config["reads_per_step"] = n_reads
    ids = list(range(len(graphs)))

    owner = engine.Context(device=local_rank)
    # region setup = graph upload + index build on the device (enumerate, radix sorts, table) for all 20 regions.
    # The first call pays one-time CUDA module / pinned-pool initialisation, so the steady-state figure is the median of the rest.
    setup_times = []
    for rep in range(5):
        if rep:
            for k in ids:
                owner.region_end(k)
        t0 = time.perf_counter()
        owner.region_begin_multi(ids, graphs)
        for k in ids:
            owner.pool_begin(k, 1)
        setup_times.append(time.perf_counter() - t0)
    t_region = float(np.median(setup_times[1:]))
    ctxs = [owner]
    for t in range(1, n_threads):
        c = engine.Context(device=local_rank)
        for k in ids:
            c.region_attach(k, owner, k)  # shares graph + index; only the pool state is per context
            c.pool_begin(k, 1)
        ctxs.append(c)
    pinned = [engine.pin_batches(w) for w in workloads]  # page-locked host memory (gtb_host_alloc), as a production caller fills it
    batches = [p[0] for p in pinned]
    acc_bufs = [[c.alloc_accumulators(k) for k in ids] for c in ctxs]
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.from_numpy(owner.nccl_unique_id()))
        dist.broadcast(uid, 0)
        owner.nccl_init(world, rank, uid.cpu().numpy())

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        flush.fill_(1)
        torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    h2d_bytes = sum(b.nbytes_h2d() for b in batches[0])

    # with several pool threads the overlap of copies and kernels comes from the other pools: one launch sequence per submit;
    # a lone pool thread keeps the library's automatic 4-chunk pipeline
    e2e_chunks = env_int("GTB_BENCH_E2E_CHUNKS", 1 if n_threads > 1 else 0)
    for c in ctxs:
        c.set_chunks(e2e_chunks)

    def e2e_step(t, s=0):
        c = ctxs[t]
        c.pool_reset_multi(ids)
        st = c.submit_multi(ids, batches[t])
        return st, c.pool_finish_multi(ids, out=acc_bufs[t])

    # ---- parity, visible to the driver (N = 1): the GPU run's calls against the VCFs of the reference CLI on the same records
    parity = None
    cb = None
    st, accs = e2e_step(0)
    d2h_bytes = sum(sum(v.nbytes for v in a.as_dict().values()) for a in accs)
    launches_per_step = int(st.kernel_launches)
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        vdir = tempfile.mkdtemp(prefix="gtb_ref_vcfs_")
        try:
            cb = cpu_baseline(ref, sites, rs, regions, graphs, batches0, steps=1, warmup=0, keep_vcfs=vdir)
            if cb.get("kind") == "reference":
                parity = check_against_reference_vcfs(owner, accs, vdir, regions, ref, sites)
        except Exception as ex:  # keep the GPU line even if the CPU arm breaks
            cb = {"value": None, "unit": "reads/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
        finally:
            shutil.rmtree(vdir, ignore_errors=True)
        if parity is not None and (parity["mismatches"] or parity["records"] == 0):
            print(json.dumps({"error": "GPU calls differ from the reference CLI's VCF records", "parity": parity}))
            sys.exit(1)

    # ---- N > 1, verification: ONE sample's reads sharded over the ranks (mates stay together), accumulators all-reduced,
    #      compared with the single-GPU result of the same records (rank 0's sample, broadcast as the expected arrays)
    sharded = None
    if world > 1:
        expect = flat_accumulators(accs) if rank == 0 else None
        n_exp = torch.tensor([expect.size if rank == 0 else 0], device=dev, dtype=torch.int64)
        dist.broadcast(n_exp, 0)
        exp_t = torch.from_numpy(expect).to(dev) if rank == 0 else torch.zeros(int(n_exp[0]), dtype=torch.int64, device=dev)
        dist.broadcast(exp_t, 0)
        base_batches = make_workload(0, with_graphs=False)[6] if rank != 0 else workloads[0]
        shard = [abi.shard_batch(b, world)[rank] for b in base_batches]
        owner.pool_reset_multi(ids)
        owner.submit_multi(ids, shard)
        owner.allreduce_multi(ids)
        got = flat_accumulators(owner.pool_finish_multi(ids))
        same = got.size == int(n_exp[0]) and bool(np.array_equal(got, exp_t.cpu().numpy()))
        flag = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        sharded = {"values_compared": int(got.size), "ranks": world, "equal_on_all_ranks": bool(int(flag[0]))}
        if not sharded["equal_on_all_ranks"]:
            if rank == 0:
                print(json.dumps({"error": "read-sharded accumulators differ from the single-GPU result", "sharded": sharded}))
            sys.exit(1)

    # ---- e2e: host buffers through the public C-ABI calls, K steps back to back over n_threads pool threads
    # sample-sharded job end (N > 1): every pool thread summarises its own pools when it is done, as the reference's pool
    # threads do (Variant::scan_calls inside parallel_reader_genotype_only, hts_parallel_reader.cpp:1018-1022); then ONE
    # reduce: the rank's pools are merged while they are packed, NCCL sum / max over the ranks
    per_ctx = [None] * len(ctxs)

    def summarise(t):
        per_ctx[t] = ctxs[t].scan_calls_multi(acc_bufs[t])

    def final_reduce():
        t0 = time.perf_counter()
        for t in range(len(ctxs)):
            if per_ctx[t] is None:
                summarise(t)
        t1 = time.perf_counter()
        owner.allreduce_varstats_multi(per_ctx)
        t2 = time.perf_counter()
        V, A, R = per_ctx[0]
        out = {"leftover_scan_calls_ms": (t1 - t0) * 1e3, "merge_and_nccl_ms": (t2 - t1) * 1e3,
               "checksum": int(V.sum() % (1 << 62)), "bytes_reduced": int(V.nbytes + A.nbytes + R.nbytes)}
        for t in range(len(ctxs)):
            per_ctx[t] = None
        return out

    for s in range(warmup):
        e2e_step(s % n_threads)
    if world > 1:
        final_reduce()  # warm-up of the reduce as well (NCCL sets its channels up on the first collective)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    wall_e2e = run_pipelined(n_threads, steps, e2e_step, at_end=summarise if world > 1 else None)
    t_reduce = None
    if world > 1:
        tr0 = time.perf_counter()
        t_reduce = final_reduce()
        t_reduce["final_reduce_ms"] = (time.perf_counter() - tr0) * 1e3
        wall_e2e += time.perf_counter() - tr0
    torch.cuda.synchronize()
    ev1.record()
    torch.cuda.synchronize()
    e2e_t = max(wall_e2e, ev0.elapsed_time(ev1) * 1e-3) / steps
    barrier()

    # one step alone through the same calls (what a single pool thread sees)
    single_e2e = []
    ctxs[0].set_chunks(0)  # a lone pool thread: the library's own copy / compute pipeline
    for _ in range(3):
        e2e_step(0)  # (its chunk buffers are allocated on first use)
    for _ in range(5):
        flush_l2()
        t0 = time.perf_counter()
        e2e_step(0)
        single_e2e.append(time.perf_counter() - t0)

    # ---- device-resident: every context replays its batch already in HBM
    for t, c in enumerate(ctxs):
        c.set_chunks(env_int("GTB_BENCH_REPLAY_CHUNKS", 1))  # nothing to copy: one launch sequence per step
        c.pool_reset_multi(ids)
        c.submit_multi(ids, batches[t])
        c.set_chunks(0)

    def replay_step(t, s=0):
        ctxs[t].pool_reset_multi(ids)
        return ctxs[t].replay()

    for s in range(warmup):
        replay_step(s % n_threads)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0.record()
    wall_dev = run_pipelined(n_threads, steps, replay_step)
    torch.cuda.synchronize()
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    step_t = max(wall_dev, ev0.elapsed_time(ev1) * 1e-3) / steps
    barrier()

    # ---- latency of ONE step alone + per-kernel CUDA-event times (L2 flushed before each)
    kt = {"prep_kernels": [], "probe_kernel": [], "chain_kernel": [], "slow_kernel": [], "score_kernel": []}
    lat_ms = []
    chain_t, n_slow = None, 0
    for _ in range(5):
        flush_l2()
        owner.pool_reset_multi(ids)
        owner.replay()
        _, a, s_, _ = owner.last_timing()
        lat_ms.append(a + s_)
        k = owner.last_kernel_timing()
        for nm in kt:
            kt[nm].append(k[nm])
        n_slow = k["n_slow_tasks"]
        chain_t = owner.last_chain_timing()
    kernels_ms = {nm: float(np.mean(v)) for nm, v in kt.items()}
    kernels_ms["chain_kernel"] = chain_t["chain_kernel"]
    kernels_ms["chain_general_kernel"] = chain_t["chain_general_kernel"]

    total_reads_per_step = float(np.mean(n_reads_ctx))
    if world > 1:
        tt = torch.tensor([step_t, e2e_t], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        step_t, e2e_t = float(tt[0]), float(tt[1])
        tot = torch.tensor([total_reads_per_step], device=dev, dtype=torch.float64)
        dist.all_reduce(tot)
        total_reads_per_step = float(tot[0])

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        value = total_reads_per_step / step_t
        e2e_value = total_reads_per_step / e2e_t
        table = json.load(open(KERNEL_TABLE)) if os.path.exists(KERNEL_TABLE) else {}
        per_kernel = {}
        for nm, ms in kernels_ms.items():
            row = {"ms": ms}
            if nm in ALGO_BYTES_PER_KERNEL and ms > 0:
                row["algorithmic_GBps"] = n_reads * ALGO_BYTES_PER_KERNEL[nm] / (ms * 1e-3) / 1e9
                row["frac_of_hbm_peak"] = row["algorithmic_GBps"] / peak
            row.update(table.get(nm, {}))
            per_kernel[nm] = row
        # dominant = the longest of the kernels every read goes through (chain_general / slow serve ~3 % of the reads beside
        # the first score pass: long by the clock, but a latency floor, not where the bytes move)
        dominant = max(ALGO_BYTES_PER_KERNEL, key=lambda nm: kernels_ms[nm])
        dom = per_kernel[dominant]
        achieved = value / world * ALGO_BYTES_PER_READ / 1e9  # per GPU
        config.update({
            "pool_threads": n_threads,
            "l2": "no flush between the pipelined steps: consecutive steps on a GPU belong to different contexts with different "
                  "read sets, and one step touches ~170 MB (20 MB reads, 64 MB seed / label records, 21 MB path records, "
                  "k-mer tables and labels of 20 regions) against 126 MB of L2; the single-step `latency` figures are taken "
                  "with a 256 MiB L2 flush before every step",
            "value_path": "device-resident: reset + replay of the resident launch sequence (prep, probe, chain, chain_general, "
                          "slow, huge, score x2), K steps over %d pool threads" % n_threads,
            "e2e_path": "gtb_pool_reset_multi + gtb_submit_reads_multi (%s) + gtb_pool_finish_multi, pinned "
                        "host buffers, K steps over %d pool threads sharing the regions (gtb_region_attach)"
                        % ("one launch sequence per submit" if e2e_chunks == 1 else "the library's concurrent chunks", n_threads)
                        + ("; the job ends with ONE NCCL reduce of the per-variant summaries" if world > 1 else ""),
            "host_cores": host_cores})
        out = {
            "metric": metric, "value": value, "unit": "reads/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": step_t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_t * 1e3},
            "gpu_launches": launches_per_step * steps,
            "latency": {"device_ms_one_step_alone": float(np.mean(lat_ms)),
                        "e2e_ms_one_step_alone": float(np.mean(single_e2e)) * 1e3,
                        "reads_per_s_one_step_alone": n_reads / (float(np.mean(lat_ms)) * 1e-3),
                        "note": "one pool thread, L2 flushed before the step; the tiers after chain_kernel overlap, so the "
                                "kernel times below do not add up to it"},
            "kernels_ms": {**kernels_ms, "slow_tasks": n_slow, "chain_tiers": chain_t},
            "region_setup_s": t_region,
            "e2e_incl_region_setup": {"value": total_reads_per_step / (e2e_t + t_region), "unit": "reads/s",
                                      "note": "index build + graph upload of the 20 regions charged to every step"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "definition": "step level, SURVEY 8(d): reads/s per GPU x 6 468 algorithmic B/read; the path is "
                                       "latency / issue bound, not HBM bound (tables L2-resident, absent keys answered from "
                                       "shared memory)",
                         "traffic": dom.get("dram_bytes"), "kernel": dominant,
                         "kernel_achieved": dom.get("algorithmic_GBps"), "kernel_frac": dom.get("frac_of_hbm_peak"),
                         "per_kernel": per_kernel, "peak_source": peak_src, "algorithmic_bytes_per_read": ALGO_BYTES_PER_READ},
            "clocks": clocks,
        }
        if parity is not None:
            out["parity_checked"] = parity["records"]
            out["parity"] = parity
        if sharded is not None:
            out["sharded_parity_checked"] = sharded["values_compared"]
            out["sharded_parity"] = sharded
        if t_reduce is not None:
            out["multi_gpu"] = {**t_reduce, "axis": "samples (SURVEY 8e primary): accumulators and calls stay on their GPU"}
        if cb is not None:
            out["cpu_baseline"] = {k: cb.get(k) for k in ("value", "unit", "cores", "kind", "sample", "align_accumulate_span")
                                   if k in cb}
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["sw"] = sw_key(owner)
            except Exception as ex:
                out["sw"] = {"error": repr(ex)}
            try:
                out["bgzf"] = bgzf_key(owner, ids, graphs, regions, rs)
            except Exception as ex:
                out["bgzf"] = {"error": repr(ex)}
            if out["bgzf"].get("accumulators_equal_record_entry") is False:
                print(json.dumps({"error": "the BGZF entry's accumulators differ from the record entry's", "bgzf": out["bgzf"]}))
                sys.exit(1)
            try:
                out["cli"] = cli_end_to_end(ref, sites, rs, regions)
            except Exception as ex:
                out["cli"] = {"error": repr(ex)}
            if out["cli"].get("vcfs_identical") is False:
                print(json.dumps({"error": "the drop-in binary's VCFs differ from the reference's", "cli": out["cli"]}))
                sys.exit(1)
        print(json.dumps(out))
    for c in ctxs[1:]:
        c.close()
    owner.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
