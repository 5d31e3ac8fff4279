"""A/B of the ring kernel variants on one GPU (CUDA events, inputs resident; not the bench):
    python tools/variant_sweep.py [--out gpurun_out/variants.md] [--shapes c2,c3,c4,d200,d100] [--kernels 0,2,3,4,5]
For every BASELINE-shaped configuration and every cfg.kernel value: positions/s, algorithmic GB/s, fraction of the
measured HBM peak, shards (CTAs) used, and the loss per position as a sanity check that the variant trains the same
thing.  Suggested first GPU call of a round:
    W2B_TEST_EXPERIMENTAL=1 python -m pytest -m gpu tests/test_gpu_variant.py -x -q && python tools/variant_sweep.py"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import word2bits_b200 as w2b
from tools.quick_perf import synth

SHAPES = {  # name: (D, negative, bitlevel, window)
    "c2": (800, 24, 1, 10), "c3": (400, 12, 2, 10), "c4": (400, 24, 0, 10), "d200": (200, 24, 1, 8), "d100": (100, 5, 1, 5),
}


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/variants.md")
    ap.add_argument("--shapes", default="c2,c3,c4,d200,d100")
    ap.add_argument("--kernels", default="0,2,3,4,5")
    ap.add_argument("--groups", default="0,7", help="landing-group sizes to try with the variants (0 = default 13)")
    ap.add_argument("--vocab", type=int, default=400000)
    ap.add_argument("--tokens", type=int, default=40_000_000)
    ap.add_argument("--words", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    ids, cn = synth(a.vocab, a.tokens)
    pk = peak()
    rows = ["| shape | kernel | release | lanes/row | consumer warps | shards | positions/s | algorithmic GB/s | of %.0f GB/s | loss/position |" % pk,
            "|---|---|---|---|---|---|---|---|---|---|"]
    for name in a.shapes.split(","):
        D, neg, b, W = SHAPES[name]
        for kernel in [int(k) for k in a.kernels.split(",")]:
            plan = w2b.ring_plan(size=D, window=W, negative=neg, bitlevel=b, kernel=kernel, vocab_size=a.vocab + 1)
            lpr = 32 // max(plan["units_per_warp"], 1)
            base = w2b.ring_plan(size=D, window=W, negative=neg, bitlevel=b, kernel=2, vocab_size=a.vocab + 1)
            if kernel >= 3 and (lpr, plan["consumer_warps"]) == (32, base["consumer_warps"]):
                continue  # variant does not apply to this width: it would repeat kernel 2
            combos = [(0, 0)] if kernel < 2 else [(r, g) for r in (0, 2) for g in [int(x) for x in a.groups.split(",")]
                                                  if not (r == 0 and g)]  # other group sizes only with early release
            for release, group in combos:  # ring_serial 2 = early slot release (variants only)
                t = w2b.Trainer(None, vocab_size=a.vocab + 1, size=D, window=W, negative=neg, bitlevel=b, iter=1,
                                threads=None, kernel=kernel, ring_serial=release, group=group)
                S = t.threads
                t.set_vocab_counts(cn, int(a.tokens))
                t.set_corpus(ids, np.arange(S, dtype=np.int64) * (a.tokens // S), np.full(S, -1, np.int32), True)
                t.train_step(2000)
                pos = rows_ = 0
                ms = loss = 0.0
                for _ in range(a.steps):
                    st = t.train_step(a.words)
                    pos += st["positions"]; rows_ += st["context_rows"] + st["target_rows"]; ms += st["kernel_ms"]; loss += st["loss"]
                t.close()
                gbs = rows_ * D * 4 * 2 / 1e9 / (ms / 1e3)
                line = "| %s D=%d neg=%d b=%d | %d | %s | %d | %d | %d | %.2f M | %.0f | %.3f | %.4f |" % (
                    name, D, neg, b, kernel, ("early, G=%d" % (group or 13)) if release == 2 else "at commit", lpr, plan["consumer_warps"], S,
                    pos / ms / 1e3, gbs, gbs / pk, loss / max(pos, 1))
                print(line, flush=True)
                rows.append(line)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        f.write("# ring kernel variants, tools/variant_sweep.py (CUDA events; %d words per shard per step, %d steps)\n\n" % (a.words, a.steps))
        f.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
