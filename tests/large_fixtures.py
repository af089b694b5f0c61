"""The two large-shape fixtures (BASELINE configs 3 and 4): definitions shared by the generator
(tests/golden/make_golden_large.py, runs the compiled reference here) and the GPU test (tests/test_gpu_large.py).

The inputs are too large to commit (5*10^5 and 2.4*10^6 records), so they are REGENERATED from seeds on the test box; what
is committed (tests/golden/large/*.gtba.gz) is what only the reference can give: its graph, which records of which file its
pool loop processed (flag filter / is_good_read), and its accumulators after the pool.  The generator also checks that the
regenerated batch is the reference's record stream (same files, same order, same duplicate decisions, same AS-XS)."""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np

from graphtyper_b200 import abi, gtba, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LARGE_DIR = os.path.join(ROOT, "tests", "golden", "large")

CONFIGS: Dict[str, dict] = {
    # configs[2] shape: one 50 kb region (+1 kb pads) of a 10 variants/kb graph, a pool of 50 samples at 30x
    "pool50": dict(kind="snp", length=52000, n_sites=500, n_samples=50, seed=301, coverage=30.0, region="chr1:1001-51000"),
    # configs[3] shape: genotype_sv window of 1.2 Mb, 400 <DEL>/<INS>/<DUP> of 50..1500 bp, 10 samples at 30x
    "sv1m": dict(kind="sv", length=1_200_000, n_sites=400, n_samples=10, seed=401, coverage=30.0, max_size=1500, spacing=1000,
                 orphan_rate=0.02, region="chr1:1-1200000"),
}


def build(name: str) -> dict:
    """Reference, sites, genotypes and one ReadSet per sample, from the seeds of CONFIGS[name]."""
    kw = CONFIGS[name]
    L, seed = kw["length"], kw["seed"]
    ref = synth.make_reference(L, seed)
    if kw["kind"] == "sv":
        sites = synth.make_sv_sites(ref, kw["n_sites"], seed=seed + 1, max_size=kw["max_size"], spacing=kw["spacing"],
                                    p_del=0.5, p_ins=0.3)
    else:
        sites = synth.make_sites(ref, kw["n_sites"], seed + 1)
    gts = synth.make_genotypes(len(sites), kw["n_samples"], seed + 2)
    rng = np.random.default_rng(seed + 3)
    readsets: List[synth.ReadSet] = []
    for k in range(kw["n_samples"]):
        rs = synth.simulate_reads(ref, sites, gts[k], f"SAMP{k + 1:03d}", seed + 10 + k, coverage=kw["coverage"],
                                  err=0.003, lowmapq_rate=0.03)
        if kw["kind"] == "sv":  # orphaned mates -> the leftover reads genotype_sv processes at pool end
            keep = np.nonzero((rng.random(len(rs)) > kw["orphan_rate"]) & (rs.pos >= 0) & (rs.mpos >= 0))[0]
            rs = rs.subset(keep)
        readsets.append(rs)
    return {"ref": ref, "sites": sites, "gts": gts, "readsets": readsets, "is_sv": kw["kind"] == "sv", "region": kw["region"],
            "length": L}


def kept_indices(name: str, n_samples: int) -> List[np.ndarray]:
    d = gtba.load(os.path.join(LARGE_DIR, f"{name}.kept.gtba"))
    return [np.nonzero(np.unpackbits(d[f"kept_{k}"])[:int(d["n_records"][k])])[0] for k in range(n_samples)]


def batch(ds: dict, kept: List[np.ndarray]) -> abi.HostBatch:
    """The pool's batch in the reference's merge order, from the records its loop processed."""
    return abi.batch_from_readsets(ds["readsets"], region_idx=kept)
