#!/bin/bash
# Round-2 GPU call 16: suite after the stream-ordering change, bench lines (c2 + reference arm, c3, c4), compute-sanitizer
# on the production kernel, full-size L3 at 1776 shards on a 36 M-token corpus, quality at the CLI's new default.
set -u
mkdir -p gpurun_out
S=gpurun_out/call16_summary.txt
: > $S
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_suite16.log 2>&1
echo "pytest -m gpu -x: exit $?" | tee -a $S
tail -3 gpurun_out/gpu_suite16.log | tee -a $S
timeout 600 python bench.py > gpurun_out/bench16.json 2> gpurun_out/bench16.err
echo "bench c2: exit $?" | tee -a $S
timeout 600 python bench.py --impl reference > gpurun_out/bench16_ref.json 2> gpurun_out/bench16_ref.err
echo "bench reference arm: exit $?" | tee -a $S
for tool in memcheck synccheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "compute-sanitizer $tool: exit $?" | tee -a $S
  tail -4 gpurun_out/sanitizer_$tool.txt | tee -a $S
done
timeout 600 python tests/tools/full_size_l3.py 36000000 1776 > gpurun_out/l3_36m_1776.txt 2>&1
echo "full-size L3 36M tokens, 1776 shards: exit $?" | tee -a $S
cat gpurun_out/l3_36m_1776.txt | tail -3 | tee -a $S
for w in c3 c4; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench16_$w.json 2> gpurun_out/bench16_$w.err
  echo "bench $w: exit $?" | tee -a $S
done
timeout 300 python tests/tools/quality_planted.py 0 > gpurun_out/quality_default16.txt 2>&1
echo "quality at the CLI default: exit $?" | tee -a $S
cat gpurun_out/quality_default16.txt | tee -a $S
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches16.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu16.log 2>&1
echo "ncu launch list: exit $?" | tee -a $S
du -sh gpurun_out | tee -a $S
