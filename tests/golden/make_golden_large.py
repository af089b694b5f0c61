#!/usr/bin/env python
"""Generates tests/golden/large/*.gtba.gz: golden vectors of the two large-shape fixtures (tests/large_fixtures.py) from
the compiled, UNMODIFIED reference (oracle/_ref/bin/gt_probe --light).  Only runnable where /root/reference was compiled.

  python tests/golden/make_golden_large.py [--only=pool50|sv1m]

Per fixture: <name>.graph.gtba.gz (the reference's graph), <name>.kept.gtba.gz (which records of every sample's file the
reference's pool loop processed), <name>.accum.gtba.gz (its accumulators after the pool).  Before anything is written the
regenerated batch is checked against the reference's record stream, and the oracle's accumulators against the reference's."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import compare  # noqa: E402
import large_fixtures as lf  # noqa: E402
import oracle  # noqa: E402
from graphtyper_b200 import abi, gtba, synth  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def run(name: str) -> None:
    t0 = time.time()
    ds = lf.build(name)
    L = ds["length"]
    tmp = tempfile.mkdtemp(prefix="gtb_large_")
    try:
        fa = os.path.join(tmp, "ref.fa")
        synth.write_fasta(fa, ds["ref"])
        vcf = os.path.join(tmp, "sites.vcf")
        if ds["is_sv"]:
            synth.write_sv_vcf(vcf, ds["sites"], "chr1", L)
        else:
            synth.write_vcf(vcf, ds["sites"], "chr1", L)
        subprocess.run([os.path.join(BIN, "bgzip"), "-f", vcf], check=True)
        subprocess.run([os.path.join(BIN, "tabix"), "-f", "-p", "vcf", vcf + ".gz"], check=True)
        sams = []
        for k, rs in enumerate(ds["readsets"]):
            sam = os.path.join(tmp, f"s{k:03d}.sam")
            synth.write_sam(sam, rs, "chr1", L)
            sams.append(sam)
        print(f"{name}: inputs written ({sum(len(r) for r in ds['readsets'])} records, {time.time() - t0:.0f} s)", flush=True)
        pre = os.path.join(tmp, name)
        cmd = [os.path.join(BIN, "gt_probe"), "--ref", fa, "--vcf", vcf + ".gz", "--region", ds["region"], "--light",
               "--sams", ",".join(sams), "--out", pre] + (["--sv"] if ds["is_sv"] else [])
        subprocess.run(cmd, check=True, stderr=subprocess.DEVNULL)
        print(f"{name}: reference pool loop done ({time.time() - t0:.0f} s)", flush=True)
        st = gtba.load(pre + ".stream.gtba")
        ns = len(ds["readsets"])
        # which records of which file were processed
        kept = []
        for k, rs in enumerate(ds["readsets"]):
            m = st["file"] == k
            key = st["name_id"][m].astype(np.int64) * 2 + ((st["flag"][m] & 64) != 0)
            mine = rs.name_id.astype(np.int64) * 2 + ((rs.flag & 64) != 0)
            assert len(np.unique(mine)) == len(mine), "record identity (pair id, first-in-pair) is not unique"
            idx = np.nonzero(np.isin(mine, key))[0]
            assert len(idx) == int(m.sum()), (k, len(idx), int(m.sum()))
            kept.append(idx)
        b = abi.batch_from_readsets(ds["readsets"], region_idx=kept)
        # The k-way merge orders records by (position, bases); records that tie (the same read sequence at the same position
        # in several samples) leave the reference's heap in a heap-dependent order.  They are equal_pos_seq duplicates of
        # each other, so nothing downstream depends on their order: the stream is compared up to that order -- same records
        # per sample, same number of duplicate decisions, same AS-XS values -- and the decisive check is the next one
        # (the oracle, run on the regenerated batch, must reproduce the reference's accumulators).
        assert len(b) == len(st["flag"]), (len(b), len(st["flag"]))
        assert np.array_equal(np.bincount(b.sample, minlength=ns), np.bincount(st["sample"], minlength=ns))
        assert np.array_equal(np.sort(b.flag), np.sort(st["flag"])), "flags differ from the reference's stream"
        assert int((b.dup_of >= 0).sum()) == int((st["isdup"] != 0).sum()), "duplicate decisions differ from the reference loop"
        assert np.array_equal(np.sort(b.score_diff), np.sort(st["score_diff"])), "AS-XS differs from the reference's update_paths"
        g = abi.HostGraph.from_gtba(gtba.load(pre + ".graph.gtba"))
        golden = compare.probe_accum(gtba.load(pre + ".accum.gtba"))
        O = oracle.Oracle()
        h = O.index_build(g)
        acc = O.result_accum(O.pool_run(g, h, ns, b, tap=False), ns)
        compare.compare_accum(golden, {k: v for k, v in acc.as_dict().items() if k != "saturated"}, name + " oracle")
        print(f"{name}: regenerated batch == reference stream, oracle accumulators == reference ({time.time() - t0:.0f} s)",
              flush=True)
        os.makedirs(lf.LARGE_DIR, exist_ok=True)
        keptd = {"n_records": np.array([len(r) for r in ds["readsets"]], np.int64)}
        for k in range(ns):
            m = np.zeros(len(ds["readsets"][k]), np.uint8)
            m[kept[k]] = 1
            keptd[f"kept_{k}"] = np.packbits(m)
        gtba.save(os.path.join(lf.LARGE_DIR, f"{name}.kept.gtba"), keptd)
        shutil.copy(pre + ".graph.gtba", os.path.join(lf.LARGE_DIR, f"{name}.graph.gtba"))
        shutil.copy(pre + ".accum.gtba", os.path.join(lf.LARGE_DIR, f"{name}.accum.gtba"))
        for s in ("kept", "graph", "accum"):
            subprocess.run(["gzip", "-9", "-n", "-f", os.path.join(lf.LARGE_DIR, f"{name}.{s}.gtba")], check=True)
        print(f"{name}: wrote", {s: os.path.getsize(os.path.join(lf.LARGE_DIR, f"{name}.{s}.gtba.gz"))
                                 for s in ("kept", "graph", "accum")}, flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    for nm in lf.CONFIGS:
        if not only or nm in only:
            run(nm)
