#!/bin/bash
# Round-2 GPU call 9: tensor-core evaluator (tcgen05 + TMA filter, fp32 re-score) — exactness tests, timing at the
# Google-set shape, SASS evidence; whole GPU suite after the clean-up.
set -u
mkdir -p gpurun_out
: > gpurun_out/call9_summary.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "evaluator" > gpurun_out/eval_tests.log 2>&1
echo "evaluator tests: exit $?" | tee -a gpurun_out/call9_summary.txt
grep -E "passed|failed|^evaluator D=|^FAILED|^ERROR|Error" gpurun_out/eval_tests.log | tail -15
timeout 900 python tests/tools/eval_perf.py 400000 800 19544 eval-only > gpurun_out/eval_perf.txt 2>&1
echo "eval perf: exit $?" | tee -a gpurun_out/call9_summary.txt
cat gpurun_out/eval_perf.txt | tail -6
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "not evaluator" > gpurun_out/parity9.log 2>&1
echo "GPU parity suite: exit $?" | tee -a gpurun_out/call9_summary.txt
tail -4 gpurun_out/parity9.log
