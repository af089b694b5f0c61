#!/usr/bin/env python
"""Generates the golden vectors of the parity tests by running the compiled, UNMODIFIED reference
(oracle/_ref/bin/gt_probe, built by oracle/ref_build/Makefile) on seeded synthetic inputs.

  python tests/golden/make_golden.py          # small fixtures -> tests/golden/      (committed)
  python tests/golden/make_golden.py --big    # larger fixtures -> tests/golden/big/*.gtba.gz (committed, gzip -9)

Only runnable where /root/reference was compiled (this container); the GPU box uses the produced files.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from graphtyper_b200 import synth  # noqa: E402

BIN = os.path.join(ROOT, "oracle", "_ref", "bin")

SMALL = {
    # name: (make_dataset kwargs, region_size)
    "tiny": (dict(length=3000, n_sites=7, n_samples=1, seed=5, coverage=30), 50000),
    "mini_stress": (dict(length=8000, n_sites=160, n_samples=3, seed=21, coverage=12, err=0.01, n_rate=0.002,
                         lowmapq_rate=0.1, unpaired_rate=0.05, improper_rate=0.08, flip_rate=0.5), 50000),
    "mini_r100": (dict(length=6000, n_sites=40, n_samples=1, seed=31, coverage=20, err=0.004, read_len=100), 50000),
    # multi-allelic bubbles, overlapping / adjacent records, GT_ID / GT_ANTI_HAPLOTYPE events
    "mini_complex": (dict(length=7000, n_sites=90, n_samples=2, seed=91, coverage=14, err=0.004, complex_sites=True), 50000),
}
# structural-variant fixtures (genotype_sv flow, is_sv_graph = true): kwargs of the SV generator below
SMALL_SV = {
    "mini_sv": dict(length=16000, n_sites=3, n_samples=2, seed=61, coverage=14, max_size=400, orphan_rate=0.06),
}
BIG_SV = {
    "sv30k": dict(length=30000, n_sites=6, n_samples=2, seed=71, coverage=20, max_size=600, orphan_rate=0.06),
    "sv60k": dict(length=60000, n_sites=14, n_samples=3, seed=81, coverage=15, max_size=1000, orphan_rate=0.04),
}
BIG = {
    "r60k": (dict(length=60000, n_sites=600, n_samples=1, seed=11), 50000),
    "stress": (dict(length=30000, n_sites=600, n_samples=3, seed=21, coverage=20, err=0.01, n_rate=0.002,
                    lowmapq_rate=0.1, unpaired_rate=0.05, improper_rate=0.08, flip_rate=0.5), 50000),
    "r100": (dict(length=20000, n_sites=150, n_samples=1, seed=31, coverage=25, err=0.004, read_len=100), 50000),
    "dense": (dict(length=20000, n_sites=1200, n_samples=2, seed=41, coverage=20, err=0.006, n_rate=0.001), 50000),
    "complex": (dict(length=40000, n_sites=700, n_samples=3, seed=93, coverage=18, err=0.005, n_rate=0.0005,
                     complex_sites=True), 50000),
}


def run(name: str, kwargs: dict, region_size: int, out_dir: str) -> None:
    tmp = tempfile.mkdtemp(prefix="gtb_golden_")
    try:
        ds = synth.make_dataset(**kwargs)
        man = synth.write_dataset(ds, tmp, region_size=region_size)
        subprocess.run([os.path.join(BIN, "bgzip"), "-f", "-k", man["vcf"]], check=True)
        subprocess.run([os.path.join(BIN, "tabix"), "-f", "-p", "vcf", man["vcf"] + ".gz"], check=True)
        for k, reg in enumerate(man["regions"]):
            pre = os.path.join(out_dir, f"{name}.r{k}")
            subprocess.run([os.path.join(BIN, "gt_probe"), "--ref", man["fasta"], "--vcf", man["vcf"] + ".gz",
                            "--region", f"{man['contig']}:{reg['begin']}-{reg['end']}",
                            "--sams", ",".join(reg["sams"]), "--out", pre], check=True)
            print("wrote", pre, {s: os.path.getsize(pre + s) for s in
                                 (".graph.gtba", ".index.gtba", ".reads.gtba", ".accum.gtba")})
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_sv(name: str, kw: dict, out_dir: str) -> None:
    """<DEL>/<INS>/<DUP> sites, reads from the carrier haplotypes, a fraction of mates dropped (they become the
    'leftover' reads genotype_sv processes at pool end); probe run with --sv (is_sv_graph = true)."""
    import numpy as np
    tmp = tempfile.mkdtemp(prefix="gtb_golden_sv_")
    try:
        L = kw["length"]
        ref = synth.make_reference(L, kw["seed"])
        sites = synth.make_sv_sites(ref, kw["n_sites"], seed=kw["seed"] + 1, max_size=kw["max_size"], p_del=0.5, p_ins=0.3)
        gts = synth.make_genotypes(len(sites), kw["n_samples"], kw["seed"] + 2)
        fa = os.path.join(tmp, "ref.fa")
        synth.write_fasta(fa, ref)
        vcf = os.path.join(tmp, "sv.vcf")
        synth.write_sv_vcf(vcf, sites, "chr1", L)
        subprocess.run([os.path.join(BIN, "bgzip"), "-f", vcf], check=True)
        subprocess.run([os.path.join(BIN, "tabix"), "-f", "-p", "vcf", vcf + ".gz"], check=True)
        rng = np.random.default_rng(kw["seed"] + 3)
        sams = []
        for k in range(kw["n_samples"]):
            rs = synth.simulate_reads(ref, sites, gts[k], f"SAMP{k + 1}", kw["seed"] + 10 + k, coverage=kw["coverage"],
                                      err=0.004, lowmapq_rate=0.05)
            keep = np.nonzero((rng.random(len(rs)) > kw["orphan_rate"]) & (rs.pos >= 0) & (rs.mpos >= 0))[0]
            sam = os.path.join(tmp, f"s{k}.sam")
            synth.write_sam(sam, rs.subset(keep), "chr1", L)
            sams.append(sam)
        pre = os.path.join(out_dir, f"{name}.r0")
        subprocess.run([os.path.join(BIN, "gt_probe"), "--ref", fa, "--vcf", vcf + ".gz", "--region", f"chr1:1-{L}", "--sv",
                        "--sams", ",".join(sams), "--out", pre], check=True, stderr=subprocess.DEVNULL)
        print("wrote", pre, {s: os.path.getsize(pre + s) for s in (".graph.gtba", ".index.gtba", ".reads.gtba", ".accum.gtba")})
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def main() -> None:
    big = "--big" in sys.argv
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    out_dir = os.path.join(ROOT, "tests", "golden", "big") if big else os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (kw, rs) in (BIG if big else SMALL).items():
        if not only or name in only:
            run(name, kw, rs, out_dir)
    for name, kw in (BIG_SV if big else SMALL_SV).items():
        if not only or name in only:
            run_sv(name, kw, out_dir)
    if big:  # the committed form is gzip-compressed (gtba.load reads either)
        import glob
        for p in glob.glob(os.path.join(out_dir, "*.gtba")):
            subprocess.run(["gzip", "-9", "-n", "-f", p], check=True)


if __name__ == "__main__":
    main()
