"""Shared comparison helpers: probe (reference) dumps vs oracle / CUDA outputs."""
from __future__ import annotations

import numpy as np


def compare_index(ref: dict, got: dict) -> None:
    assert np.array_equal(ref["keys"], got["keys"]), "index keys differ"
    assert np.array_equal(ref["label_off"], got["label_off"]), "index label offsets differ"
    assert np.array_equal(ref["labels"], got["labels"]), "index labels (order-sensitive) differ"


def probe_paths(d: dict) -> dict:
    """Normalises the probe's path dump into the gtb_debug_paths layout."""
    n = len(d["p_start"])
    f = np.stack([d["p_start"], d["p_end"], d["p_rs"], d["p_re"], d["p_mm"], d["p_nvar"]], axis=1).reshape(-1)
    return {"gp_npaths": d["gp_npaths"], "gp_longest": d["gp_longest"], "p_fields": f.astype(np.uint32),
            "v_order": d["v_order"], "v_nnum": d["v_nnum"], "v_nums": d["v_nums"]}


def compare_paths(ref: dict, got: dict, what: str = "") -> None:
    for k in ("gp_npaths", "gp_longest", "p_fields", "v_order", "v_nnum", "v_nums"):
        a, b = np.asarray(ref[k]), np.asarray(got[k])
        if a.shape != b.shape or not np.array_equal(a, b):
            n = min(len(a), len(b))
            bad = np.nonzero(a[:n] != b[:n])[0]
            first = int(bad[0]) if len(bad) else n
            raise AssertionError(f"{what} paths.{k} differ: shapes {a.shape} vs {b.shape}, first diff at {first}: "
                                 f"{a[max(0, first - 3):first + 4]} vs {b[max(0, first - 3):first + 4]}")


def probe_seeds(d: dict) -> dict:
    """Probe seeds: per non-dup record, orientation fwd then rev; s0 (exact) and s1 (Hamming-1) separately.
    Returns the gtb_debug_seeds layout: nslots[(u*2+o)*2+h], then per-slot counts and labels in that order."""
    n_units = len(d["s0_nslots"]) // 2
    nsl0, nsl1 = d["s0_nslots"], d["s1_nslots"]
    nslots = np.zeros(n_units * 4, np.uint32)
    nslots[0::2] = nsl0
    nslots[1::2] = nsl1
    # interleave per (unit, orientation): h0 slots then h1 slots
    c0 = np.concatenate([[0], np.cumsum(nsl0)]).astype(np.int64)
    c1 = np.concatenate([[0], np.cumsum(nsl1)]).astype(np.int64)
    l0 = np.concatenate([[0], np.cumsum(d["s0_nlabels"])]).astype(np.int64)
    l1 = np.concatenate([[0], np.cumsum(d["s1_nlabels"])]).astype(np.int64)
    nl, lab = [], []
    lab0 = d["s0_labels"].reshape(-1, 3)
    lab1 = d["s1_labels"].reshape(-1, 3)
    for uo in range(n_units * 2):
        nl.append(d["s0_nlabels"][c0[uo]:c0[uo + 1]])
        lab.append(lab0[l0[c0[uo]]:l0[c0[uo + 1]]])
        nl.append(d["s1_nlabels"][c1[uo]:c1[uo + 1]])
        lab.append(lab1[l1[c1[uo]]:l1[c1[uo + 1]]])
    return {"nslots": nslots, "nlabels": np.concatenate(nl) if nl else np.zeros(0, np.uint32),
            "labels": (np.concatenate(lab) if lab else np.zeros((0, 3), np.uint32)).reshape(-1)}


def compare_seeds(ref: dict, got: dict, what: str = "") -> int:
    """Compares seed label lists for every (unit, orientation) the implementation computed (nslots > 0)."""
    rn, gn = ref["nslots"], got["nslots"]
    assert len(rn) == len(gn), f"{what} seed unit count differs {len(rn)} vs {len(gn)}"
    rc = np.concatenate([[0], np.cumsum(rn)]).astype(np.int64)
    gc = np.concatenate([[0], np.cumsum(gn)]).astype(np.int64)
    rl = np.concatenate([[0], np.cumsum(ref["nlabels"])]).astype(np.int64)
    gl = np.concatenate([[0], np.cumsum(got["nlabels"])]).astype(np.int64)
    rlab = ref["labels"].reshape(-1, 3)
    glab = got["labels"].reshape(-1, 3)
    checked = 0
    for k in range(len(rn)):
        if gn[k] == 0:
            continue
        assert rn[k] == gn[k], f"{what} seed slot count differs at {k}: {rn[k]} vs {gn[k]}"
        a = ref["nlabels"][rc[k]:rc[k + 1]]
        b = got["nlabels"][gc[k]:gc[k + 1]]
        assert np.array_equal(a, b), f"{what} seed label counts differ at list {k}: {a} vs {b}"
        la = rlab[rl[rc[k]]:rl[rc[k + 1]]]
        lb = glab[gl[gc[k]]:gl[gc[k + 1]]]
        assert np.array_equal(la, lb), f"{what} seed labels differ at list {k}"
        checked += 1
    return checked


def probe_accum(d: dict) -> dict:
    """Probe accumulators -> gtb_accumulators layout (the probe already iterates bubble-major, sample-minor)."""
    ns, nb = int(d["meta"][0]), int(d["meta"][1])
    num = d["hap_num"].astype(np.uint64)
    score_off = np.concatenate([[0], np.cumsum(num * (num + 1) // 2)]).astype(np.uint64)
    cov_off = np.concatenate([[0], np.cumsum(num)]).astype(np.uint64)
    extra = {"ref_depth": d["ref_depth"]} if "ref_depth" in d else {}
    return {**extra, "bubble_id": d["hap_id"], "n_alleles": d["hap_num"], "score_off": score_off, "cov_off": cov_off,
            "log_score": d["log_score"], "gt_coverage": d["gt_cov"], "max_log_score": d["max_log_score"],
            "ambiguous_depth": d["amb"], "ambiguous_depth_alt": d["amb_alt"], "alt_proper_pair_depth": d["alt_pp"],
            "vs_clipped_reads": d["vs_clipped_reads"], "vs_mapq_squared": d["vs_mapq_sq"],
            "pa_clipped_bp": d["pa_clipped_bp"], "pa_mapq_squared": d["pa_mapq_sq"],
            "pa_score_diff": d["pa_score_diff"], "pa_mismatches": d["pa_mismatches"], "read_strand": d["rs_counts"]}


def compare_accum(ref: dict, got: dict, what: str = "") -> None:
    for k, a in ref.items():
        b = np.asarray(got[k])
        a = np.asarray(a)
        if a.shape != b.shape or not np.array_equal(a.astype(np.uint64), b.astype(np.uint64)):
            n = min(len(a), len(b))
            bad = np.nonzero(a[:n].astype(np.uint64) != b[:n].astype(np.uint64))[0]
            raise AssertionError(f"{what} accumulators.{k} differ: shapes {a.shape} vs {b.shape}; "
                                 f"{len(bad)} mismatches, first at {bad[:5]}: {a[bad[:5]]} vs {b[bad[:5]]}")


def probe_connections(d: dict) -> np.ndarray:
    """Golden HapSample::connections -> uint32 [n, 6] rows (sample hap1 allele1 hap2 allele2 count), sorted."""
    return np.concatenate([d["conn_tuples"].reshape(-1, 5).astype(np.uint32),
                           d["conn_counts"].astype(np.uint32)[:, None]], axis=1)


def compare_connections(ref: np.ndarray, got: np.ndarray, what: str = "") -> None:
    if ref.shape != got.shape or not np.array_equal(ref, got):
        rs, gs = {tuple(r) for r in ref.tolist()}, {tuple(r) for r in got.tolist()}
        raise AssertionError(f"{what} connections differ: {len(ref)} vs {len(got)} entries; "
                             f"missing {sorted(rs - gs)[:5]} extra {sorted(gs - rs)[:5]}")
