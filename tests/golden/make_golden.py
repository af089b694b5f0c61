#!/usr/bin/env python
"""Generates the golden vectors of the parity tests by running the compiled, UNMODIFIED reference
(oracle/_ref/bin/gt_probe, built by oracle/ref_build/Makefile) on seeded synthetic inputs.

  python tests/golden/make_golden.py          # small fixtures -> tests/golden/      (committed)
  python tests/golden/make_golden.py --big    # larger fixtures -> tests/data_local/ (git-ignored, ships to the GPU box)

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
}
BIG = {
    "r60k": (dict(length=60000, n_sites=600, n_samples=1, seed=11), 50000),
    "stress": (dict(length=30000, n_sites=600, n_samples=3, seed=21, coverage=20, err=0.01, n_rate=0.002,
                    lowmapq_rate=0.1, unpaired_rate=0.05, improper_rate=0.08, flip_rate=0.5), 50000),
    "r100": (dict(length=20000, n_sites=150, n_samples=1, seed=31, coverage=25, err=0.004, read_len=100), 50000),
    "dense": (dict(length=20000, n_sites=1200, n_samples=2, seed=41, coverage=20, err=0.006, n_rate=0.001), 50000),
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


def main() -> None:
    big = "--big" in sys.argv
    out_dir = os.path.join(ROOT, "tests", "data_local" if big else "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (kw, rs) in (BIG if big else SMALL).items():
        run(name, kw, rs, out_dir)


if __name__ == "__main__":
    main()
