"""One short training step for ncu: python tools/prof_step.py [D] [neg] [bits] [kernel] [words]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import word2bits_b200 as w2b
from tools.quick_perf import synth

D = int(sys.argv[1]) if len(sys.argv) > 1 else 800
neg = int(sys.argv[2]) if len(sys.argv) > 2 else 24
b = int(sys.argv[3]) if len(sys.argv) > 3 else 1
kernel = int(sys.argv[4]) if len(sys.argv) > 4 else 0
words = int(sys.argv[5]) if len(sys.argv) > 5 else 1500
V, N = 400000, 24_000_000
ids, cn = synth(V, N)
t = w2b.Trainer(None, vocab_size=V + 1, size=D, window=10 if D != 200 else 8, negative=neg, bitlevel=b, iter=1, kernel=kernel)
S = t.threads
t.set_vocab_counts(cn, N)
t.set_corpus(ids, np.arange(S, dtype=np.int64) * (N // S), np.full(S, -1, np.int32), True)
for _ in range(3):
    st = t.train_step(words)
    print(st["positions"], st["kernel_ms"], st["positions"] / st["kernel_ms"] / 1e3, "Mpos/s")
