"""Pins the scalar restatement of paw::pairwise_alignment (oracle/paw_oracle.cpp) against the golden results of the
compiled, unmodified paw (tests/golden/sw_pairs.gtba, made by tests/golden/make_golden_sw.py) and, where the reference
is compiled (this container), against fresh random pairs."""
import os

import numpy as np
import pytest

import oracle
from graphtyper_b200 import gtba, synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "sw_pairs.gtba")


def load_golden():
    g = gtba.load(GOLDEN)
    q = [bytes(g["q"][a:b]) for a, b in zip(g["q_off"][:-1], g["q_off"][1:])]
    d = [bytes(g["d"][a:b]) for a, b in zip(g["d_off"][:-1], g["d_off"][1:])]
    return q, d, g["result"].reshape(-1, 5)


def test_oracle_matches_golden_pairs():
    q, d, want = load_golden()
    assert len(q) >= 700
    got = oracle.PawOracle().align(q, d)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, f"{bad.size} pairs differ, first {bad[0]}: oracle {got[bad[0]]} paw {want[bad[0]]}"


def test_golden_covers_clipping_and_gaps():
    """The fixture is only a pin if it exercises the interesting branches."""
    q, d, want = load_golden()
    qlen = np.array([len(x) for x in q])
    dlen = np.array([len(x) for x in d])
    assert (want[:, 3] > 0).sum() > 20            # clip_begin
    assert (want[:, 4] < qlen).sum() > 20         # clip_end
    assert (want[:, 1] > 0).sum() > 100           # free leading window bases
    assert (want[:, 2] < dlen).sum() > 100        # free trailing window bases
    assert (want[:, 0] < qlen - 8).sum() > 100    # gaps / mismatches


@pytest.mark.skipif(oracle.ref_binary("paw_probe") is None, reason="compiled reference not present")
@pytest.mark.parametrize("seed", [7, 8])
def test_oracle_matches_compiled_paw_on_random_pairs(seed):
    q, d = synth.make_sw_pairs(1500, seed=seed)
    want = oracle.paw_reference(q, d)
    got = oracle.PawOracle().align(q, d)
    assert np.array_equal(got, want)
