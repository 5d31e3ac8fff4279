#!/bin/bash
# Round-2 GPU call 11: padded rows (D % 4 != 0) and -reg on the production kernel: GPU suite + sweep.
set -u
mkdir -p gpurun_out
: > gpurun_out/call11_summary.txt
timeout 2400 python -m pytest tests -m gpu -q -rA > gpurun_out/gpu_suite11.log 2>&1
echo "pytest -m gpu: exit $?" | tee -a gpurun_out/call11_summary.txt
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/gpu_suite11.log | tail -12
timeout 600 python tools/warp_sweep.py --shapes c2,c3,c4,d200 --configs 0:0:0:1 --out gpurun_out/warp_sweep11.md > gpurun_out/warp_sweep11.log 2>&1
echo "sweep: exit $?" | tee -a gpurun_out/call11_summary.txt
cat gpurun_out/warp_sweep11.md
python - <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import word2bits_b200 as w2b
from tools.quick_perf import synth
ids, cn = synth(400000, 20_000_000)
for D, reg in ((50, 0.0), (150, 0.0), (250, 0.0), (300, 0.0), (800, 0.001), (200, 0.001)):
    t = w2b.Trainer(None, vocab_size=400001, size=D, window=8, negative=24, bitlevel=1, iter=1, threads=None, reg=reg)
    S = t.threads
    t.set_vocab_counts(cn, 20_000_000)
    t.set_corpus(ids, np.arange(S, dtype=np.int64) * (20_000_000 // S), np.full(S, -1, np.int32), True)
    t.train_step(500)
    st = t.train_step(1500)
    gbs = (st["context_rows"] + st["target_rows"]) * D * 8 / 1e9 / (st["kernel_ms"] / 1e3)
    print("D=%d reg=%g shards=%d: %.1f M positions/s, %.0f GB/s algorithmic (%.2f of 6577)" % (D, reg, S, st["positions"] / st["kernel_ms"] / 1e3, gbs, gbs / 6577.4), flush=True)
    t.close()
PY
