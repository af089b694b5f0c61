"""GPU suite (-m gpu) of the discovery re-alignment kernel (gtb_sw_align_batch): bit-exact against the golden results
of the compiled paw (tests/golden/sw_pairs.gtba) and against the scalar oracle on fresh seeded pairs."""
import numpy as np
import pytest

import oracle
from graphtyper_b200 import engine, synth
from test_sw_oracle import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = engine.Context(device=0)
    yield c
    c.close()


def assert_same(got, want, q, d):
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (f"{bad.size}/{len(q)} pairs differ; first {bad[0]} (m={len(q[bad[0]])}, n={len(d[bad[0]])}): "
                           f"cuda {got[bad[0]]} want {want[bad[0]]}")


def test_sw_matches_compiled_paw_golden(ctx):
    q, d, want = load_golden()
    assert_same(ctx.sw_align(q, d), want, q, d)


@pytest.mark.parametrize("seed,n", [(11, 4000), (12, 20000)])
def test_sw_matches_oracle_on_random_pairs(ctx, seed, n):
    q, d = synth.make_sw_pairs(n, seed=seed)
    want = oracle.PawOracle().align(q, d)
    assert_same(ctx.sw_align(q, d), want, q, d)


def test_sw_maximum_sizes_and_ragged_batch(ctx):
    """Query up to GTB_SW_MAX_QUERY (160), window up to GTB_SW_MAX_DATABASE (2048), lengths 1..max mixed in one batch
    (a long window next to 1-base ones exercises the per-warp scratch sizing)."""
    rng = np.random.default_rng(5)
    B = np.frombuffer(b"ACGT", np.uint8)
    q, d = [], []
    for m, n in [(160, 2048), (1, 1), (1, 2048), (160, 1), (159, 2047), (5, 7), (151, 1300), (160, 160), (2, 1)]:
        w = bytes(B[rng.integers(0, 4, size=n)])
        st = int(rng.integers(0, max(1, n - m)))
        r = bytearray(w[st:st + m].ljust(m, b"A"))
        for _ in range(m // 25):
            r[int(rng.integers(0, m))] = int(B[rng.integers(0, 4)])
        q.append(bytes(r))
        d.append(w)
    want = oracle.PawOracle().align(q, d)
    assert_same(ctx.sw_align(q, d), want, q, d)
    # same pairs in reverse order: results do not depend on which warp / scratch slab handles a pair
    assert_same(ctx.sw_align(q[::-1], d[::-1]), want[::-1], q[::-1], d[::-1])


def test_sw_replay_is_idempotent_and_timed(ctx):
    q, d = synth.make_sw_pairs(3000, seed=13)
    a = ctx.sw_align(q, d)
    ctx.sw_replay()
    t = ctx.sw_last_timing()
    assert t["kernel_ms"] > 0
    assert np.array_equal(ctx.sw_align(q, d), a)


def test_sw_input_errors(ctx):
    assert ctx.sw_align([], []).shape == (0, 5)
    with pytest.raises(engine.GtbError):
        ctx.sw_align([b"A" * 161], [b"ACGT"])
    with pytest.raises(engine.GtbError):
        ctx.sw_align([b"ACGT"], [b"A" * 2049])
    with pytest.raises(engine.GtbError):
        ctx.sw_align([b""], [b"ACGT"])
