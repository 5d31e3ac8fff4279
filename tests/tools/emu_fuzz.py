"""TEST INFRASTRUCTURE: random geometries through the emulated production kernel (tests/emu) — protocol checks
(dead-lock, carve-up bounds, async-proxy rules, mbarrier phases observed) on every run, loss against the sequential
oracle in serial mode.   python tests/tools/emu_fuzz.py [seconds] [seed]"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import word2bits_b200 as w2b
from oracle import pyoracle as po
from tests.emu import emu
from tests.util import zipf_corpus

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
tmp = tempfile.mkdtemp()
corpora = []
for i, (n, v, nl) in enumerate(((3000, 200, 40), (5000, 500, 0), (2500, 60, 7))):
    path = zipf_corpus(os.path.join(tmp, "c%d.txt" % i), n, v, seed=10 + i, newline_every=nl)
    c, o = w2b.Corpus(path, 1), po.Corpus(path, 1)
    corpora.append((c, o, po.unigram_table(o.counts)))
t0, runs, skipped = time.time(), 0, 0
worst = []
while time.time() - t0 < budget:
    c, o, table = corpora[rng.integers(len(corpora))]
    D = int(rng.choice([rng.integers(1, 40), rng.integers(40, 300), rng.integers(300, 2049)], p=[0.4, 0.4, 0.2]))
    W = int(rng.choice([rng.integers(1, 12), rng.integers(12, 513)], p=[0.8, 0.2]))
    neg = int(rng.choice([rng.integers(0, 8), rng.integers(8, 64)], p=[0.6, 0.4]))
    b = int(rng.choice([0, 1, 2, 3, 5]))
    S = int(rng.integers(1, 4))
    serial = int(rng.integers(0, 2))
    slots = int(rng.choice([0, 3, 4, 5, 9, 16]))
    reg = float(rng.choice([0.0, 0.0, 1e-3]))
    amode = int(rng.integers(0, 3))
    seed = int(rng.integers(1, 1 << 30))
    cfg = dict(size=D, window=W, negative=neg, bitlevel=b, shards=S, serial=serial, slots=slots, reg=reg,
               async_mode=amode, seed=seed)
    if D > 600 and (W > 20 or neg > 30):  # keep single runs short
        skipped += 1
        continue
    try:
        plan = w2b.warp_plan(size=D, window=W, negative=neg, bitlevel=b, vocab_size=c.vocab_size, slots=slots, reg=reg)
    except Exception as ex:
        print("plan refused", cfg, ex); continue
    if not plan["warp"]:
        skipped += 1
        continue
    u, v = po.init_net(c.vocab_size, D)
    try:
        out = emu.train_epoch_warp(c, table, u, v, **cfg)
    except emu.EmuError as ex:
        if "no emulated instantiation" in str(ex):  # the emulator builds a subset of the row widths
            skipped += 1
            continue
        print("FAIL", cfg, plan, ex, flush=True)
        sys.exit(1)
    ok = out["done"].tolist() == [1] * S and np.isfinite(u).all() and np.isfinite(v).all()
    # (very wide windows at fp32 diverge chaotically on these tiny corpora — rounding-level differences blow up — so
    # the loss is compared for windows <= 64 only, as in tests/test_gpu_parity.py)
    if ok and serial and S == 1 and b in (0, 1, 2, 5) and W <= 64:
        m = po.OracleModel(o, D, W, neg, b, shards=1, iters=1, table=table, reg=reg)
        lo, tr = m.train_shard(0, trace_cap=20000)
        lg = out["loss"].sum()
        # duplicate targets inside a position both read the old row in this kernel (DESIGN: deviation 1); on these
        # tiny vocabularies most positions have some, and with 1-/2-bit rows the drift is visible in the loss
        dup = float(np.mean([len(set(t[3])) < len(t[3]) for t in tr if t[2] > 0] or [0.0]))
        gap = abs(lg - lo) / (abs(lo) + 1e-9)
        worst.append((gap, dup, dict(cfg)))
        if gap > 0.10:  # (observed: up to ~5 % at D > 1000 on the 60-word vocabulary, 1e-4 .. 3e-3 without duplicates)
            print("LOSS", cfg, lg, lo, flush=True)
            ok = False
        if out["words"].sum() != m.word_count_actual:
            print("COUNT", cfg, out["words"].sum(), m.word_count_actual, flush=True)
            ok = False
    if not ok:
        print("FAIL", cfg, plan, flush=True)
        sys.exit(1)
    runs += 1
print("emu fuzz: %d runs ok, %d skipped in %.0f s" % (runs, skipped, time.time() - t0))
for gap, dup, cfg in sorted(worst, key=lambda x: -x[0])[:5]:
    print("  loss gap %.2e (positions with duplicate targets %.2f): %s" % (gap, dup, cfg))
