#!/bin/bash
# Round-2 sixth GPU call: whole GPU suite on the leaner kernel, sweep, bench, DRAM traffic at the bench step size.
set -u
mkdir -p gpurun_out
: > gpurun_out/call6_summary.txt
timeout 600 python tools/warp_sweep.py --shapes c2,c3,c4,d200,d100 --configs 0:0:0:1,0:1:0:1 --out gpurun_out/warp_sweep6.md > gpurun_out/warp_sweep6.log 2>&1
echo "sweep: exit $?" | tee -a gpurun_out/call6_summary.txt
cat gpurun_out/warp_sweep6.md
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA > gpurun_out/parity6.log 2>&1
echo "GPU parity suite: exit $?" | tee -a gpurun_out/call6_summary.txt
grep -E "passed|failed|full-size L3|rel-L2 vs oracle|^FAILED|^ERROR" gpurun_out/parity6.log | tail -30
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err
echo "bench: exit $?" | tee -a gpurun_out/call6_summary.txt
cat gpurun_out/bench6.json
bash tools/measure_traffic.sh c2 c3 c4 | tee -a gpurun_out/call6_summary.txt
du -sh gpurun_out
