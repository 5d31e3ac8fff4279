import sys, os, time, faulthandler, signal
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.register(signal.SIGUSR1, all_threads=True)
import numpy as np
import word2bits_b200 as w2b
from tests.util import zipf_corpus
path = zipf_corpus("/tmp/medium.txt", 60000, 3000, seed=2)
c = w2b.Corpus(path, 5)
D, W, neg, b = [int(x) for x in sys.argv[1:5]]
kernel = int(sys.argv[5]) if len(sys.argv) > 5 else 0
serial = int(sys.argv[6]) if len(sys.argv) > 6 else 0
t = w2b.Trainer(c, size=D, window=W, negative=neg, bitlevel=b, threads=6, iter=1, kernel=kernel, ring_serial=serial)
t.epoch_begin()
t0 = time.time()
for i in range(100000):
    st = t.train_step(2000)
    print(i, st["words"], st["positions"], st["shards_done"], "%.1f ms" % st["kernel_ms"], flush=True)
    if st["shards_done"] == 6:
        break
print("done", time.time() - t0)
