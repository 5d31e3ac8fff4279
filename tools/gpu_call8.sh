#!/bin/bash
# Round-2 GPU call 8: sampler state in shared memory (register diet) — warps per SM A/B, whole GPU suite.
set -u
mkdir -p gpurun_out
: > gpurun_out/call8_summary.txt
for dense in 0 1; do
  W2B_WARP_DENSE=$dense timeout 600 python tools/warp_sweep.py --shapes c2,c3,c4,d200,d100 --configs 0:0:0:1 --out gpurun_out/warp_sweep_dense$dense.md > gpurun_out/warp_sweep_dense$dense.log 2>&1
  echo "sweep dense=$dense: exit $?" | tee -a gpurun_out/call8_summary.txt
  cat gpurun_out/warp_sweep_dense$dense.md
done
W2B_WARP_DENSE=1 W2B_WARP_RED=1 timeout 600 python tools/warp_sweep.py --shapes c2,c3,d200 --configs 0:0:0:1 --out gpurun_out/warp_sweep_dense1_red1.md > /dev/null 2>&1
cat gpurun_out/warp_sweep_dense1_red1.md
W2B_WARP_DENSE=1 timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA > gpurun_out/parity8.log 2>&1
echo "GPU parity suite (dense): exit $?" | tee -a gpurun_out/call8_summary.txt
grep -E "passed|failed|full-size L3|^FAILED|^ERROR" gpurun_out/parity8.log | tail -12
