"""Shared helpers for the tests: deterministic synthetic corpora."""
import os

import numpy as np


def zipf_corpus(path, n_tokens, vocab, seed=0, newline_every=0, tail=True, exponent=1.0):
    """Space-separated tokens w<r> with p(r) ~ 1/r^exponent; optional newline every k tokens
    (k may produce empty lines when 1); optional coverage tail so every id occurs once."""
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, vocab + 1) ** exponent
    p /= p.sum()
    ids = rng.choice(vocab, size=n_tokens, p=p)
    if tail:
        ids = np.concatenate([ids, rng.permutation(vocab)])
    parts = []
    for i, t in enumerate(ids):
        parts.append("w%d" % (t + 1))
        if newline_every and (i + 1) % newline_every == 0:
            parts.append("\n")
            if rng.random() < 0.2:
                parts.append("\n")  # empty sentence
        else:
            parts.append(" ")
    with open(path, "w") as f:
        f.write("".join(parts))
    return path


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)
