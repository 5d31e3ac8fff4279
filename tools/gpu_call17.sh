#!/bin/bash
# Round-2 GPU call 17 (last): the staging job now waits on its slot's mbarrier like every other job (synccheck's
# "missing wait"): GPU suite, bench line, compute-sanitizer memcheck + synccheck on the production kernel.
set -u
mkdir -p gpurun_out
S=gpurun_out/call17_summary.txt
: > $S
timeout 300 python -m pytest tests -m gpu -q -x -k "not full_size_shape_loss and not evaluator" > gpurun_out/gpu_suite17.log 2>&1
echo "pytest -m gpu -x (without the full-size L3 and evaluator cases): exit $?" | tee -a $S
tail -2 gpurun_out/gpu_suite17.log | tee -a $S
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err
echo "bench c2: exit $?" | tee -a $S
for tool in synccheck memcheck; do
  timeout 90 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_step.py > gpurun_out/sanitizer17_$tool.txt 2>&1
  echo "compute-sanitizer $tool: exit $?" | tee -a $S
  grep -E "^D=|ERROR SUMMARY|done" gpurun_out/sanitizer17_$tool.txt | tee -a $S
done
