"""A few short training steps of the production kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_step.py
Small on purpose (the sanitizer slows a kernel by two orders of magnitude): 20k-word vocabulary, 64 shards, row
widths that exercise full, padded (D % 4 != 0) and multi-group rows, with and without -reg."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import word2bits_b200 as w2b
from tools.quick_perf import synth

V, N, S = 20000, 400_000, 64
ids, cn = synth(V, N)
for D, bits, reg, prefetch in ((200, 1, 0.0, 0), (103, 2, 0.0, 1), (800, 1, 0.0, 0), (64, 1, 1e-4, 0)):
    t = w2b.Trainer(None, vocab_size=V + 1, size=D, window=5, negative=8, bitlevel=bits, iter=1, threads=S, reg=reg,
                    prefetch=prefetch)
    t.set_vocab_counts(cn, N)
    t.set_corpus(ids, np.arange(S, dtype=np.int64) * (N // S), np.full(S, -1, np.int32), True)
    for _ in range(2):
        st = t.train_step(150)
    print("D=%d bits=%d reg=%g prefetch=%d: %d positions, loss/position %.4f, plan %s" % (
        D, bits, reg, prefetch, st["positions"], st["loss"] / max(st["positions"], 1),
        w2b.warp_plan(size=D, window=5, negative=8, bitlevel=bits, reg=reg, vocab_size=V + 1)), flush=True)
    t.close()
print("sanitize_step done")
