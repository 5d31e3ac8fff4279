#!/bin/bash
# Round-2 regression call (1 GPU): whole GPU suite, smoke(), bench (both arms), C5 on one GPU, ncu launch list.
set -u
mkdir -p gpurun_out
: > gpurun_out/call10_summary.txt
timeout 2400 python -m pytest tests -m gpu -q -rA > gpurun_out/gpu_suite.log 2>&1
echo "pytest -m gpu: exit $?" | tee -a gpurun_out/call10_summary.txt
grep -E "passed|failed|^FAILED|^ERROR|full-size L3|evaluator D=|rel-L2" gpurun_out/gpu_suite.log | tail -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke: exit $?" | tee -a gpurun_out/call10_summary.txt
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench reference arm: exit $?" | tee -a gpurun_out/call10_summary.txt
cut -c1-1200 gpurun_out/bench_ref.json
timeout 900 python bench.py > gpurun_out/bench10.json 2> gpurun_out/bench10.err
echo "bench: exit $?" | tee -a gpurun_out/call10_summary.txt
cat gpurun_out/bench10.json
timeout 900 python bench.py --workload c5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_1gpu.json 2> gpurun_out/bench_c5_1gpu.err
echo "bench c5 1 GPU: exit $?" | tee -a gpurun_out/call10_summary.txt
cut -c1-2000 gpurun_out/bench_c5_1gpu.json; tail -2 gpurun_out/bench_c5_1gpu.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_bench_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launch list: exit $?" | tee -a gpurun_out/call10_summary.txt
du -sh gpurun_out
