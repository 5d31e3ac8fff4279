"""CPU check of the production kernel's flow control: the geometry the host planner picks
(w2b_ring_plan_query — pure host arithmetic inside libw2b) is run through an executable model of the
kernel's sampler / loader / consumer protocol (tests/ring_model.py) under randomised schedules.
A geometry that can hang the GPU or recycle a shared-memory row too early fails here, without a GPU."""
import random

import pytest

import word2bits_b200 as w2b
from tests.ring_model import Deadlock, RingModel, random_positions

SMEM_LIMIT = 227 * 1024


def _check(plan, window, negative, seeds=(0, 1), n=40):
    for style in ("typical", "extreme", "tiny", "any"):
        for seed in seeds:
            rng = random.Random(1000 * seed + len(style))
            pos = random_positions(rng, n, window, negative, style)
            RingModel(plan, window, negative, pos, seed=seed).run()


def test_planner_invariants_over_the_supported_envelope():
    """Every (D, window, negative) the ABI accepts: either the planner declines (register kernel) or the
    geometry fits shared memory, holds a whole window of context rows, has enough landing barriers, and
    respects the lower bound on the v-ring that rules out a wait cycle inside one position."""
    n_ring = 0
    for D in (4, 8, 52, 100, 128, 200, 256, 300, 400, 512, 640, 800, 1000, 1024, 1028, 2048):
        for window in (1, 2, 5, 8, 10, 16, 32, 64):
            for negative in (0, 1, 5, 12, 24, 25, 31, 32, 40, 63):
                p = w2b.ring_plan(size=D, window=window, negative=negative)
                if not p["ring"]:
                    continue
                n_ring += 1
                nt, G, R, ncw = negative + 1, p["group"], p["rows_in_flight"], p["consumer_warps"]
                assert p["smem_bytes"] <= SMEM_LIMIT, (D, window, negative, p)
                assert p["u_rows"] >= 2 * window, (D, window, negative, p)
                assert (nt + G - 1) // G <= p["max_groups"] and 1 <= G <= 32, (D, window, negative, p)
                assert p["threads"] == 32 * (ncw + 2) <= 1024
                assert p["v_rows"] >= 2 * G or p["v_rows"] >= nt, (D, window, negative, p)
                assert p["v_rows"] >= nt or p["v_rows"] >= (2 * R - 1) * ncw + G, (D, window, negative, p)
    assert n_ring > 500
    # outside the envelope the planner declines instead of guessing
    assert not w2b.ring_plan(size=402, window=5, negative=5)["ring"]          # D % 4 != 0
    assert not w2b.ring_plan(size=2048, window=5, negative=5)["ring"]         # D > 1024
    assert not w2b.ring_plan(size=200, window=5, negative=5, reg=0.1)["ring"]  # regularisation
    assert not w2b.ring_plan(size=200, window=5, negative=5, mode=w2b.MODE_STRICT)["ring"]
    assert not w2b.ring_plan(size=200, window=5, negative=5, kernel=1)["ring"]


# BASELINE.json configurations + the shapes tests/test_gpu_parity.py runs on the GPU
NAMED = [(800, 10, 24), (400, 10, 12), (400, 10, 24), (200, 8, 24), (800, 10, 63), (1024, 3, 7), (256, 20, 40),
         (100, 5, 5), (8, 2, 1), (4, 1, 0), (64, 5, 6), (640, 64, 63), (1000, 2, 31)]


@pytest.mark.parametrize("D,window,negative", NAMED)
def test_named_geometries_are_live_and_safe(D, window, negative):
    p = w2b.ring_plan(size=D, window=window, negative=negative)
    if not p["ring"]:
        pytest.skip("register kernel for this shape")
    _check(p, window, negative, seeds=(0, 1, 2), n=60)


def test_sweep_is_live_and_safe():
    """Coarser positions, wider grid, plus the caller-controlled knobs (group, ring_rows)."""
    n = 0
    for D in (8, 200, 400, 800, 1024):
        for window in (1, 5, 10, 32):
            for negative in (0, 5, 24, 40, 63):
                for group, ring_rows in ((0, 0), (1, 0), (5, 0), (16, 0), (0, 1), (0, 27), (7, 30)):
                    p = w2b.ring_plan(size=D, window=window, negative=negative, group=group, ring_rows=ring_rows)
                    if not p["ring"]:
                        continue
                    n += 1
                    _check(p, window, negative, seeds=(n,), n=12)
    assert n > 300


def test_model_catches_a_ring_that_is_too_small():
    """The model must be able to fail: a v-ring below the planner's bound (with positions longer than the
    ring) dead-locks — this is the hang the bound exists to prevent."""
    p = w2b.ring_plan(size=800, window=10, negative=63)
    assert p["ring"] and p["v_rows"] < 64
    bad = dict(p, v_rows=p["group"] + 2)
    pos = [(20, 64)] * 6
    with pytest.raises((Deadlock, AssertionError)):
        for seed in range(20):
            RingModel(bad, 10, 63, pos, seed=seed).run()
    RingModel(p, 10, 63, pos, seed=0).run()  # the planner's own geometry runs the same positions


def test_division_free_row_index_is_exact():
    """kernel = 2 replaces (vs0 + i) % nv and i / G by multiply-high with a precomputed reciprocal; the same
    inline helpers run on the host here, over every operand the kernel can see (i <= 63, G <= 16, vs0 < nv)."""
    import ctypes as C
    from word2bits_b200._lib import lib
    s, g = C.c_int(), C.c_int()
    nvs = list(range(1, 130)) + [255, 256, 257, 1000, 4095, 4096, 14000, 60000]
    for nv in nvs:
        for G in (1, 2, 3, 5, 7, 8, 9, 13, 16):
            for vs0 in sorted({0, 1, nv // 2, nv - 2, nv - 1} & set(range(nv))):
                for i in range(0, 64):
                    assert lib.w2b_host_ring_index(vs0, i, nv, G, C.byref(s), C.byref(g)) == 0
                    assert (s.value, g.value) == ((vs0 + i) % nv, i // G), (nv, G, vs0, i)
    assert lib.w2b_host_ring_index(5, 0, 5, 1, C.byref(s), C.byref(g)) != 0  # vs0 must be < nv


@pytest.mark.parametrize("kernel", [2, 3, 4, 5])
def test_variant_geometries_are_live_and_safe(kernel):
    """The experimental variants (cfg.kernel 2: same geometry as the default; 3 / 4: two / four row units per
    consumer warp, which deepens the v-ring bound) through the same model."""
    n = 0
    for D in (8, 64, 100, 200, 256, 400, 512, 800):
        for window in (2, 10, 32):
            for negative in (0, 5, 12, 24, 40, 63):
                p = w2b.ring_plan(size=D, window=window, negative=negative, kernel=kernel)
                if not p["ring"]:
                    continue
                upw = p["units_per_warp"]
                want = {2: 1, 3: 2 if D <= 512 else 1, 4: 4 if D <= 256 else 1, 5: 1}[kernel]
                assert upw == want, (D, kernel, p)
                base = w2b.ring_plan(size=D, window=window, negative=negative)
                if base["ring"]:
                    assert p["consumer_warps"] == base["consumer_warps"] + (2 if kernel == 5 and D > 512 else 0)
                nt, G, R, nunits = negative + 1, p["group"], p["rows_in_flight"], p["consumer_warps"] * upw
                assert p["smem_bytes"] <= SMEM_LIMIT
                assert p["v_rows"] >= nt or p["v_rows"] >= (2 * R - 1) * nunits + G + upw - 1, (D, window, negative, p)
                n += 1
                _check(p, window, negative, seeds=(n,), n=16)
    assert n > 100
    # bit levels other than 0/1/2 stay on the measured kernel
    assert w2b.ring_plan(size=400, window=5, negative=5, bitlevel=5, kernel=3)["units_per_warp"] == 1


def test_model_finds_the_unit_bound():
    """Two row units per warp, positions longer than the ring: rings below (2R-1)*units + G rows dead-lock
    in the model, the planner's choice and everything above that threshold is live."""
    p = w2b.ring_plan(size=400, window=10, negative=63, kernel=3)
    assert p["ring"] and p["units_per_warp"] == 2
    nunits, G, R = p["consumer_warps"] * 2, p["group"], p["rows_in_flight"]
    need = (2 * R - 1) * nunits + G
    pos = [(20, 64)] * 6
    for nv, ok in ((need - 1, False), (need, True), (need + 3, True)):
        results = set()
        for seed in range(12):
            try:
                RingModel(dict(p, v_rows=nv), 10, 63, pos, seed=seed).run()
                results.add(True)
            except (Deadlock, AssertionError):
                results.add(False)
        assert results == {ok}, (nv, need, results)
    assert p["v_rows"] > need


@pytest.mark.parametrize("kernel", [2, 3, 4, 5])
def test_early_release_mode_is_live_and_safe(kernel):
    """Variant kernels with TrainParams::serial == 2: slots of a pass are handed back at the top of the next pass
    (after a wait_group.read 0) instead of at its commit."""
    n = 0
    for D in (64, 200, 400, 800):
        for window in (2, 10):
            for negative in (5, 24, 63):
                for group in (0, 7):
                    p = w2b.ring_plan(size=D, window=window, negative=negative, kernel=kernel, group=group)
                    if not p["ring"]:
                        continue
                    n += 1
                    for style in ("typical", "extreme", "any"):
                        rng = random.Random(n)
                        pos = random_positions(rng, 16, window, negative, style)
                        RingModel(p, window, negative, pos, seed=n, early_release=True).run()
    assert n >= 40
