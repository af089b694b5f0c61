#!/usr/bin/env python
"""Golden vectors of the discovery re-alignment path: seeded (read, window) pairs and the results of the compiled,
UNMODIFIED paw::pairwise_alignment (oracle/_ref/bin/paw_probe, built by oracle/ref_build/Makefile) on them.

  python tests/golden/make_golden_sw.py     ->  tests/golden/sw_pairs.gtba   (committed)

Only runnable where /root/reference was compiled (this container)."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from graphtyper_b200 import gtba, synth  # noqa: E402
from graphtyper_b200.engine import pack_sequences  # noqa: E402


def edge_pairs():
    """Hand-made corner cases: 1-base inputs, read longer than the window, all-N, all-mismatch, homopolymers
    (many equal-score ties), lower-case window."""
    A = b"ACGTTGCAAGGCTTAACCGGATCGATCGTTAGC"
    q, d = [], []

    def add(a, b):
        q.append(a)
        d.append(b)

    add(b"A", b"A")
    add(b"A", b"C")
    add(b"N", b"N")
    add(b"ACGT", b"A")
    add(A * 4, A[:20])
    add(A, A)
    add(A, A.lower())
    add(A.lower(), A)
    add(b"N" * 30, A * 3)
    add(A, b"N" * 60)
    add(b"A" * 100, b"A" * 300)
    add(b"A" * 50 + b"C" * 50, b"A" * 120 + b"C" * 120)
    add(b"AC" * 60, b"AC" * 200)
    add(b"T" * 40, b"A" * 200)
    add(A[:10] + b"GGGGGGGG" + A[10:], A * 3)
    add(A[:12] + A[20:], A * 3)
    add(b"TTTTTTTTTT" + A * 2 + b"GGGGGGGGGG", b"CC" + A * 2 + b"CC")
    add((A * 5)[:151], (b"GATTACA" * 10) + A * 5 + (b"TGCA" * 30))
    return q, d


def main() -> None:
    q, d = synth.make_sw_pairs(700, seed=101)
    eq, ed = edge_pairs()
    q, d = eq + q, ed + d
    ref = oracle.paw_reference(q, d)
    if ref is None:
        raise SystemExit("oracle/_ref/bin/paw_probe is missing: make -C oracle/ref_build")
    qb, qo = pack_sequences(q)
    db, do = pack_sequences(d)
    out = os.path.join(ROOT, "tests", "golden", "sw_pairs.gtba")
    gtba.save(out, {"q": qb, "q_off": qo, "d": db, "d_off": do, "result": ref.reshape(-1).astype(np.int32)})
    print("wrote", out, os.path.getsize(out), "bytes;", len(q), "pairs")


if __name__ == "__main__":
    main()
