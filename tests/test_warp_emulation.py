"""The warp-per-shard production kernel's own source on the CPU (tests/emu: word2bits_b200/csrc/w2b_warp.cuh compiled
with -DW2B_EMULATE, one fiber per CUDA thread; mbarrier / bulk copy / bulk reduce semantics emulated with
adversarially late completion and shuffled scheduling).  Test infrastructure, not a fallback and not a timing
model: it checks control flow, the job-queue / slot protocol, the async-proxy rules and the arithmetic against
the oracle before (or without) a GPU run; tests/test_gpu_parity.py holds the real kernel to the same bars."""
import numpy as np
import pytest

import word2bits_b200 as w2b
from oracle import pyoracle as po
from tests.emu import emu
from tests.util import zipf_corpus


@pytest.fixture(scope="module")
def tiny(tmp_path_factory):
    path = zipf_corpus(str(tmp_path_factory.mktemp("e") / "tiny.txt"), 4000, 300, seed=5, newline_every=40)
    c, o = w2b.Corpus(path, 1), po.Corpus(path, 1)
    return c, o, po.unigram_table(o.counts)


@pytest.fixture(scope="module")
def medium(tmp_path_factory):
    path = zipf_corpus(str(tmp_path_factory.mktemp("e") / "medium.txt"), 24000, 1500, seed=2)
    c, o = w2b.Corpus(path, 5), po.Corpus(path, 5)
    return c, o, po.unigram_table(o.counts)


def _run(c, table, D, W, neg, b, S, **kw):
    u, v = po.init_net(c.vocab_size, D)
    out = emu.train_epoch_warp(c, table, u, v, size=D, window=W, negative=neg, bitlevel=b, shards=S, **kw)
    assert out["done"].tolist() == [1] * S
    assert np.isfinite(u).all() and np.isfinite(v).all()
    return u, v, out


@pytest.mark.parametrize("D", [64, 200, 50, 131])
def test_sequential_mode_tracks_the_oracle(D, medium):
    """serial = 1 (position p+1 is fetched after every update of p completed), one shard, fp32: the kernel differs
    from the oracle only by reduction order / FMA and by duplicate targets inside a position — the bar the
    production kernel meets on the GPU (5e-3 on the tables, 1e-3 on the loss).  Adversarial completion order of
    the copies and reduces may only move what a duplicate target inside one position sees (its second row is
    requested before the first one's update has landed): same bars."""
    c, o, table = medium
    m = po.OracleModel(o, D, 5, 6, 0, shards=1, iters=1, table=table)
    lo = m.train_shard(0)
    u, v, out = _run(c, table, D, 5, 6, 0, 1, serial=1, async_mode=0)
    assert abs(out["loss"].sum() - lo) <= 1e-3 * abs(lo)
    assert np.abs(u - m.u).max() < 5e-3 and np.abs(v - m.v).max() < 5e-3
    assert (out["wca"], out["words"].sum()) == (m.word_count_actual, m.word_count_actual)
    assert np.float32(out["alpha"]) == np.float32(m.alpha)
    u2, v2, out2 = _run(c, table, D, 5, 6, 0, 1, serial=1, async_mode=2, seed=7)
    assert abs(out2["loss"].sum() - lo) <= 1e-3 * abs(lo)
    assert np.abs(u2 - m.u).max() < 5e-3 and np.abs(v2 - m.v).max() < 5e-3


SHAPES = [  # D, window, negative, bitlevel, shards — BASELINE shapes, wide windows, many negatives, edges
    (800, 10, 24, 1, 2), (400, 10, 12, 2, 2), (400, 10, 24, 0, 2), (200, 8, 24, 1, 3), (100, 5, 63, 1, 2),
    (64, 30, 12, 0, 2), (8, 2, 1, 2, 2), (4, 1, 0, 1, 2), (1024, 3, 7, 5, 1), (132, 64, 63, 1, 1),
    (50, 5, 6, 1, 2), (150, 5, 6, 2, 1), (6, 2, 3, 0, 2), (257, 5, 6, 1, 1), (3, 1, 1, 1, 1),  # D % 4 != 0: padded rows
    (1200, 5, 6, 1, 1), (1530, 3, 4, 0, 1), (2048, 2, 3, 2, 1),  # wider than 1024 floats (the reference publishes D = 1200)
    (64, 200, 10, 1, 1), (16, 512, 63, 1, 1),  # windows as wide as a sentence
]


@pytest.mark.parametrize("D,W,neg,b,S", SHAPES)
def test_every_geometry_trains_what_the_oracle_trains(D, W, neg, b, S, tiny):
    """Prefetching mode under shuffled scheduling on every edge geometry: terminates, trains exactly the positions /
    context rows / target rows / words of the oracle's shards, and lands near the oracle's loss (rows are read one
    update stale, which weighs heavily on a 300-word vocabulary; the GPU tests' 3000-word corpus has the real bar)."""
    c, o, table = tiny
    m = po.OracleModel(o, D, W, neg, b, shards=S, iters=1, table=table)
    lo = 0.0
    want = dict(n_pos=[], n_ctx=[], n_tgt=[])
    for s in range(S):
        loss, tr = m.train_shard(s, trace_cap=100000)
        lo += loss
        trained = [t for t in tr if t[2] > 0]
        want["n_pos"].append(len(trained)); want["n_ctx"].append(sum(t[2] for t in trained))
        want["n_tgt"].append(sum(len(t[3]) for t in trained))
    for slots in (0, 3):  # the planner's ring, and the minimum the protocol allows
        u, v, out = _run(c, table, D, W, neg, b, S, serial=0, async_mode=2, seed=11, slots=slots)
        for k in want:
            assert out[k].tolist() == want[k], (k, slots)
        assert out["wca"] == m.word_count_actual
        assert abs(out["loss"].sum() - lo) <= 0.06 * abs(lo), (slots, out["loss"].sum(), lo)


@pytest.mark.parametrize("b,D", [(0, 64), (2, 200), (0, 50)])
def test_regularised_training_tracks_the_oracle(b, D, medium):
    """-reg != 0 on the production kernel (decay of every touched row in its scatter, regularisation terms in the
    reported loss): one shard, sequential mode, against the oracle with the same reg."""
    c, o, table = medium
    reg = 0.002
    m = po.OracleModel(o, D, 5, 6, b, shards=1, iters=1, table=table, reg=reg)
    lo = m.train_shard(0)
    u, v, out = _run(c, table, D, 5, 6, b, 1, serial=1, async_mode=2, seed=3, reg=reg)
    assert abs(out["loss"].sum() - lo) <= 2e-3 * abs(lo), (out["loss"].sum(), lo)
    if b == 0:
        assert np.abs(u - m.u).max() < 5e-3 and np.abs(v - m.v).max() < 5e-3
    else:
        assert np.corrcoef(u.ravel(), m.u.ravel())[0, 1] > 0.9 and np.corrcoef(v.ravel(), m.v.ravel())[0, 1] > 0.9
    m0 = po.OracleModel(o, D, 5, 6, b, shards=1, iters=1, table=table)
    m0.train_shard(0)
    assert np.abs(m0.u - m.u).max() > 1e-3  # the regulariser really moved the weights: the comparison above is not vacuous


@pytest.mark.parametrize("D,neg", [(400, 1), (1185, 0), (50, 1)])
def test_fp32_without_duplicate_targets_is_exact_to_rounding(D, neg, tiny):
    """What separates the production kernel from one reference thread in sequential mode is (a) reduction order / FMA
    and (b) duplicate targets inside a position reading the same old row (DESIGN: deviation 1).  With at most one
    negative a position cannot hold a duplicate (a draw equal to the centre is skipped, :458), so at bitlevel 0 only
    (a) is left: the epoch loss agrees to 1e-6 relative and the tables to 1e-4 (measured 1e-9 .. 1e-8 and <= 5e-6)."""
    c, o, table = tiny
    m = po.OracleModel(o, D, 4, neg, 0, shards=1, iters=1, table=table)
    lo = m.train_shard(0)
    u, v, out = _run(c, table, D, 4, neg, 0, 1, serial=1, async_mode=2, seed=9)
    assert abs(out["loss"].sum() - lo) <= 1e-6 * abs(lo), (out["loss"].sum(), lo)
    assert np.abs(u - m.u).max() < 1e-4 and np.abs(v - m.v).max() < 1e-4


@pytest.mark.parametrize("D,W,neg,b", [(32, 512, 63, 1), (47, 200, 5, 2), (20, 300, 8, 1)])
def test_sentence_wide_windows_track_the_oracle(D, W, neg, b, tmp_path):
    """Windows as wide as a 1000-word sentence (hundreds of context rows per position, job queue of thousands of
    entries, sentence buffer outside shared memory): one shard, sequential mode, loss and counters against the oracle.
    (The oracle itself is pinned to the reference at such windows by tests/test_oracle_vs_ref.py; until round 2 its
    context buffer held 130 ids, which is why wide windows used to be compared on counters only.)"""
    path = zipf_corpus(str(tmp_path / "long.txt"), 3000, 300, seed=11, newline_every=0)  # three 1000-word sentences
    c, o = w2b.Corpus(path, 1), po.Corpus(path, 1)
    table = po.unigram_table(o.counts)
    m = po.OracleModel(o, D, W, neg, b, shards=1, iters=1, table=table)
    lo, tr = m.train_shard(0, trace_cap=10000)
    assert max(t[2] for t in tr) > 2 * 64  # more context rows in one position than any window <= 64 can give
    u, v, out = _run(c, table, D, W, neg, b, 1, serial=1, async_mode=2, seed=5)
    assert out["n_ctx"].sum() == sum(t[2] for t in tr) and out["n_tgt"].sum() == sum(len(t[3]) for t in tr if t[2] > 0)
    # (63 negatives out of a 300-word vocabulary: most positions hold duplicate targets, which both read the old row
    # in this kernel — documented deviation 1 — hence the wider bar for that case)
    assert abs(out["loss"].sum() - lo) <= (1e-2 if neg > 32 else 2e-3) * abs(lo), (out["loss"].sum(), lo)
    bar = 0.95 if neg > 32 else 0.99
    assert np.corrcoef(u.ravel(), m.u.ravel())[0, 1] > bar and np.corrcoef(v.ravel(), m.v.ravel())[0, 1] > bar


def test_sampler_trace_equals_oracle(tiny):
    c, o, table = tiny
    for neg, shard in ((40, 0), (40, 2), (5, 1)):
        tr = emu.train_epoch_warp(c, table, *po.init_net(c.vocab_size, 200), size=200, window=8, negative=neg, bitlevel=1,
                                  shards=3, trace_shard=shard, trace_cap=8000)["trace"]
        m = po.OracleModel(o, 4, 8, neg, 1, shards=3, table=table)
        _, want = m.train_shard(shard, trace_cap=8000)
        assert tr == want


def test_emulator_checks_the_async_proxy_rules(tiny):
    """Negative controls of the protocol checkers on this kernel: with fence.proxy.async dropped, and with
    wait_group.read returning early, the run is reported instead of passing."""
    c, o, table = tiny
    for fault, msg in ((1, "fence.proxy.async"), (2, "not been confirmed read")):
        u, v = po.init_net(c.vocab_size, 200)
        with pytest.raises(emu.EmuError, match=msg):
            emu.train_epoch_warp(c, table, u, v, size=200, window=8, negative=24, bitlevel=1, shards=1,
                                 async_mode=2, seed=3, fault=fault)


def test_division_free_average_is_ieee_division():
    """context_avg = sum / cw (:449) is computed as a*r followed by one Newton step on the exact remainder
    (w2b::div_by_count, the kernel's own function, run here on the host): identical bits to float32 division for
    every context count 1..128 — sums of quantized levels (many exactly zero), random values, tiny and large ones."""
    import ctypes as C
    L = emu.lib()
    L.emu_div_by_count.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(1)
    n = 200000
    for cw in list(range(1, 129)):
        k = rng.integers(-cw, cw + 1, n // 4)
        levels = (k.astype(np.float32) * np.float32(0.33333334)).astype(np.float32)
        vals = np.concatenate([levels, (rng.standard_normal(n // 4) * cw).astype(np.float32),
                               (rng.integers(-1000, 1001, n // 4) * 0.25).astype(np.float32),
                               (rng.standard_normal(n // 4) * 1e-3).astype(np.float32)]).astype(np.float32)
        out = np.empty_like(vals)
        L.emu_div_by_count(vals.ctypes.data, len(vals), cw, out.ctypes.data)
        want = (vals / np.float32(cw)).astype(np.float32)
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), cw


def test_plan_fits_an_sm():
    """Planner invariants over the shapes it accepts: K >= 3 slots, the job queue holds two positions, and the
    warps the register allocation is sized for fit the SM's 228 KB with 1 KB reserved per CTA."""
    n = 0
    for D in list(range(4, 1025, 4)) + [1, 2, 3, 50, 150, 1021, 1200, 1536, 1900, 2048]:
        for W, neg in ((1, 0), (5, 5), (10, 24), (64, 63), (8, 24), (30, 12)):
            p = w2b.warp_plan(size=D, window=W, negative=neg)
            assert p["warp"] == 1, (D, W, neg)
            assert 3 <= p["slots"] <= 32
            assert p["queue_entries"] >= 2 * (2 * W + neg + 2) and p["queue_entries"] & (p["queue_entries"] - 1) == 0
            assert p["warps_per_sm"] % 4 == 0
            assert p["warps_per_sm"] * (p["smem_bytes"] + 1024) <= 228 * 1024
            n += 1
    assert n == (256 + 10) * 6
    assert w2b.warp_plan(size=2052, window=5, negative=5)["warp"] == 0   # wider than the instantiated kernels
    assert w2b.warp_plan(size=6, window=5, negative=5)["smem_bytes"] == w2b.warp_plan(size=8, window=5, negative=5)["smem_bytes"]  # rows padded to 16 bytes
    assert w2b.warp_plan(size=64, window=5, negative=5, reg=0.1)["warps_per_sm"] == 20  # -reg: own instantiations
