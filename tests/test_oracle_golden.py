"""Oracle vs the committed golden vectors (generated from the unmodified reference by
tests/golden/make_golden.py).  Runs anywhere — no /root/reference, no oracle/_ref, no GPU."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from tests.util import bits

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CORPUS = os.path.join(G, "golden_corpus.txt")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "reference_strict.npz"))


def test_quantize(gold):
    for b in range(9):
        got = np.array([bits(np.float32(po.lib().w2bo_quantize(float(x), b))) for x in gold["q_x"]], np.uint32)
        assert np.array_equal(got, gold["q_bits"][b]), b


def test_lcg_known_states():
    # first states from seed 0
    L = po.lib()
    assert L.w2bo_lcg(0) == 11
    assert L.w2bo_lcg(11) == 277363943098
    assert L.w2bo_lcg(277363943098) == 2389171320405252413


def test_exptable(gold):
    assert np.array_equal(bits(po.exptable()), bits(gold["exptable"]))


@pytest.mark.parametrize("mc", [1, 5])
def test_vocab(gold, mc):
    c = po.Corpus(CORPUS, mc)
    assert c.words() == list(gold["vocab_words_mc%d" % mc])
    assert np.array_equal(c.counts, gold["vocab_counts_mc%d" % mc])
    assert c.train_words == int(gold["train_words_mc%d" % mc])
    assert c.file_size == int(gold["file_size"])


def test_init_and_table(gold):
    c = po.Corpus(CORPUS, 1)
    u, v = po.init_net(c.vocab_size, 8)
    assert np.array_equal(bits(u), bits(gold["init_u"])) and np.array_equal(bits(v), bits(gold["init_v"]))
    assert np.array_equal(po.unigram_bounds(c.counts), gold["table_starts"])
    t = po.unigram_table(c.counts)
    s = gold["table_starts"]
    assert np.array_equal(t[s[:-1]], np.arange(c.vocab_size)) and np.array_equal(t[s[1:] - 1], np.arange(c.vocab_size))


@pytest.mark.parametrize("k", range(6))
def test_trajectories(gold, k):
    D, W, neg, b, shards, mc, iters = [int(x) for x in gold["case%d_cfg" % k]]
    sample, reg = [float(x) for x in gold["case%d_fcfg" % k]]
    c = po.Corpus(CORPUS, mc)
    m = po.OracleModel(c, D, W, neg, b, shards=shards, iters=iters, sample=sample, reg=reg)
    losses = []
    for _ in range(iters):
        for sid in range(shards):
            losses.append(m.train_shard(sid))
    assert np.array_equal(bits(m.u), bits(gold["case%d_u" % k]))
    assert np.array_equal(bits(m.v), bits(gold["case%d_v" % k]))
    assert np.array_equal(np.array(losses), gold["case%d_loss" % k])
    assert bits(np.float32(m.alpha)) == bits(gold["case%d_alpha" % k])
    assert m.word_count_actual == int(gold["case%d_wca" % k])
