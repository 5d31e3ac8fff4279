"""Generates tests/golden/*.npz + golden_corpus.txt from the UNMODIFIED reference
(oracle/_ref/libw2b_ref_strict.so, built by oracle/Makefile from /root/reference).
Run in the dev container:  python tests/golden/make_golden.py
The fixtures pin the oracle (tests/test_oracle_golden.py) where oracle/_ref is absent."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402
from tests.util import bits, zipf_corpus  # noqa: E402

CASES = [  # D, W, neg, bits, shards, min_count, sample, reg, iters
    (8, 3, 4, 1, 1, 1, 1e-3, 0.0, 1),
    (8, 3, 4, 2, 2, 1, 1e-3, 0.0, 1),
    (8, 3, 4, 0, 1, 1, 1e-3, 0.0, 2),
    (8, 3, 4, 5, 3, 2, 1e-2, 0.0, 1),
    (8, 3, 4, 1, 2, 1, 1e-3, 0.01, 1),
    (200, 8, 24, 1, 1, 1, 1e-3, 0.0, 1),   # BASELINE.json configs[0] shape (bitlevel 1, size 200, window 8, negative 24, 1 thread)
]


def main():
    ref = po.Ref("strict")
    corpus = os.path.join(HERE, "golden_corpus.txt")
    zipf_corpus(corpus, 12500, 30, seed=1, newline_every=15)
    out = {}
    xs = np.array([0.0, -0.0, 1e-30, -1e-30, .25, .5, np.nextafter(np.float32(.5), np.float32(1)), .75, 1.0,
                   -1.0, 3.7, -3.7, 1 / 32, .0624, .0625, .09375, .49999, -.5, -.50001, .124, .126]
                  + list(np.random.default_rng(0).uniform(-1.5, 1.5, 200)), np.float32)
    out["q_x"] = xs
    out["q_bits"] = np.array([[bits(ref.quantize(x, b)) for x in xs] for b in range(9)], np.uint32)
    ref.configure(corpus, 8, 3, 4, 1, min_count=1)
    out["exptable"] = ref.exptable()
    for mc in (1, 5):
        ref.configure(corpus, 8, 3, 4, 1, min_count=mc)
        ref.learn_vocab()
        out["vocab_words_mc%d" % mc] = np.array(ref.words())
        out["vocab_counts_mc%d" % mc] = ref.counts()
        out["train_words_mc%d" % mc] = np.int64(ref.train_words)
        out["file_size"] = np.int64(ref.file_size)
    ref.configure(corpus, 8, 3, 4, 1, min_count=1)
    ref.learn_vocab()
    ref.init_net()
    ref.init_unigram()
    out["init_u"] = ref.u().copy()
    out["init_v"] = ref.v().copy()
    t = ref.table()
    # compact, lossless form of the 1e8-entry table: first slot of every word
    starts = np.concatenate([[0], np.nonzero(np.diff(t))[0] + 1, [po.TABLE_SIZE]]).astype(np.int64)
    assert len(starts) == ref.V + 1 and np.array_equal(t[starts[:-1]], np.arange(ref.V))
    out["table_starts"] = starts
    for k, (D, W, neg, b, shards, mc, sample, reg, iters) in enumerate(CASES):
        ref.configure(corpus, D, W, neg, b, threads=shards, iters=iters, min_count=mc, sample=sample, reg=reg)
        ref.learn_vocab()
        ref.init_net()
        ref.init_unigram()
        losses = []
        for _ in range(iters):
            for sid in range(shards):
                losses.append(ref.train_thread(sid))
        out["case%d_cfg" % k] = np.array([D, W, neg, b, shards, mc, iters], np.int64)
        out["case%d_fcfg" % k] = np.array([sample, reg], np.float32)
        out["case%d_u" % k] = ref.u().copy()
        out["case%d_v" % k] = ref.v().copy()
        out["case%d_loss" % k] = np.array(losses, np.float64)
        out["case%d_alpha" % k] = np.float32(ref.alpha)
        out["case%d_wca" % k] = np.int64(ref.word_count_actual)
    np.savez_compressed(os.path.join(HERE, "reference_strict.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_strict.npz"), os.path.getsize(corpus), "byte corpus")


if __name__ == "__main__":
    main()
