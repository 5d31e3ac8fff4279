"""Quick throughput probe on one GPU (not the bench): synthetic Zipf ids, a few timed steps."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import word2bits_b200 as w2b


def synth(V, N, seed=42):
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, V + 1); cdf = np.cumsum(p); cdf /= cdf[-1]
    ids = np.searchsorted(cdf, rng.random(N)).astype(np.int32) + 1
    cn = np.bincount(ids, minlength=V + 1).astype(np.int64)
    order = np.argsort(-cn[1:], kind="stable")
    remap = np.zeros(V + 1, np.int32); remap[order + 1] = np.arange(1, V + 1)
    ids = remap[ids]
    cn = np.concatenate([[0], cn[1:][order]])
    cn = np.maximum(cn, 1); cn[0] = 0
    return ids, cn


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 800
    neg = int(sys.argv[3]) if len(sys.argv) > 3 else 24
    b = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    N = int(sys.argv[5]) if len(sys.argv) > 5 else 30_000_000
    groups = [int(x) for x in sys.argv[6].split(",")] if len(sys.argv) > 6 else [0]
    threads_list = [int(x) for x in sys.argv[7].split(",")] if len(sys.argv) > 7 else [0]
    plain = 0  # (argv[8] was the removed plain_store knob; kept so that later positions do not shift)
    kernel = int(sys.argv[9]) if len(sys.argv) > 9 else 0
    slots = int(sys.argv[10]) if len(sys.argv) > 10 else 0
    ids, cn = synth(V, N)
    for group in groups:
      for threads in threads_list:
        t = w2b.Trainer(None, vocab_size=V + 1, size=D, window=10, negative=neg, bitlevel=b, iter=1,
                        threads=threads or None, group=group, kernel=kernel, slots=slots)
        S = t.threads
        t.set_vocab_counts(cn, int(N))
        start = (np.arange(S, dtype=np.int64) * (N // S))
        t.set_corpus(ids, start, np.full(S, -1, np.int32), True)
        t.train_step(2000)  # warm-up
        tot_w = tot_p = 0; ms = 0.0; rows = 0
        t0 = time.time()
        for _ in range(3):
            st = t.train_step(8000)
            tot_w += st["words"]; tot_p += st["positions"]; ms += st["kernel_ms"]; rows += st["context_rows"] + st["target_rows"]
        wall = time.time() - t0
        gb = rows * D * 4 * 2 / 1e9
        print("kernel=%d rows=%d V=%d D=%d neg=%d b=%d group=%d plain=%d shards=%d: %.2f M words/s, %.2f M pos/s (kernel), alg %.0f GB/s, wall %.2fs kernel %.1f ms loss/pos ~ alpha=%.4f" % (
            kernel, slots, V, D, neg, b, group, plain, S, tot_w / ms / 1e3, tot_p / ms / 1e3, gb / (ms / 1e3), wall, ms, st["alpha"]), flush=True)
        t.close()


if __name__ == "__main__":
    main()
