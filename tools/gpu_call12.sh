#!/bin/bash
# Round-2 GPU call 12: rows wider than 1024 floats on the production kernel.
set -u
mkdir -p gpurun_out
: > gpurun_out/call12_summary.txt
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "odd_shapes or single_step or strict or checkpoint or full_size_shape_properties" > gpurun_out/gpu_suite12.log 2>&1
echo "pytest subset: exit $?" | tee -a gpurun_out/call12_summary.txt
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/gpu_suite12.log | tail -12
python - <<'PY'
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import word2bits_b200 as w2b
from tools.quick_perf import synth
ids, cn = synth(400000, 20_000_000)
for D, b in ((1000, 1), (1200, 1), (1200, 2), (1536, 0), (2048, 1)):
    t = w2b.Trainer(None, vocab_size=400001, size=D, window=10, negative=24, bitlevel=b, iter=1, threads=None)
    S = t.threads
    t.set_vocab_counts(cn, 20_000_000)
    t.set_corpus(ids, np.arange(S, dtype=np.int64) * (20_000_000 // S), np.full(S, -1, np.int32), True)
    t.train_step(500)
    st = t.train_step(2500)
    gbs = (st["context_rows"] + st["target_rows"]) * D * 8 / 1e9 / (st["kernel_ms"] / 1e3)
    print("D=%d b=%d shards=%d: %.1f M positions/s, %.0f GB/s algorithmic (%.2f of 6577)" % (D, b, S, st["positions"] / st["kernel_ms"] / 1e3, gbs, gbs / 6577.4), flush=True)
    t.close()
PY
