#!/bin/bash
# Round-2 first GPU call (run through gpurun from the repo root):
#   gpurun --timeout 1800 -- 'bash tools/first_gpu_call.sh'
# 1. the opt-in tests of the round-1 ring-kernel variants (cfg.kernel 2..5), each pytest under its own timeout;
# 2. the A/B sweep of the variants against the measured default on the BASELINE shapes (CUDA events);
# 3. tools/micro/membench: memory-system ceiling of a warp-per-stream gather + scatter-add (no arithmetic).
# Everything lands in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 tools/micro/membench > gpurun_out/membench.md 2> gpurun_out/membench.err
echo "membench: exit $?" | tee -a gpurun_out/first_call_summary.txt
export W2B_TEST_EXPERIMENTAL=1
ok=0  # the sweep only times variants whose tests passed (a variant that hangs would cost the whole sweep)
for v in 2 5 3 4; do
  W2B_VARIANTS=$v timeout 420 python -m pytest tests/test_gpu_variant.py -m gpu -x -q \
    > gpurun_out/variant_tests_k$v.log 2>&1
  rc=$?
  echo "variant $v tests: exit $rc" | tee -a gpurun_out/first_call_summary.txt
  [ $rc -eq 0 ] && ok="$ok,$v"
done
timeout 600 python tools/variant_sweep.py --kernels "$ok" --shapes c2,c3,c4,d200 --groups 0 --steps 3 \
  --out gpurun_out/variants.md > gpurun_out/variant_sweep.log 2>&1
echo "variant sweep: exit $?" | tee -a gpurun_out/first_call_summary.txt
cat gpurun_out/variants.md 2>/dev/null
tail -3 gpurun_out/variant_tests_k*.log
cat gpurun_out/membench.md
