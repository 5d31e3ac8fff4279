#!/bin/bash
# ncu --set full captures of the production kernel (exported to text on the box) and the GPU suite.
set -u
mkdir -p gpurun_out
: > gpurun_out/ncu_captures_summary.txt
cap() {  # name D neg bits
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_warp -s 1 -c 1 \
    -o /tmp/$1 python tools/prof_step.py $2 $3 $4 0 1500 > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1: exit $?" | tee -a gpurun_out/ncu_captures_summary.txt
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
  tail -3 gpurun_out/ncu_$1.log
}
cap r02_final_c2 800 24 1
cap r02_final_c3 400 12 2
cap r02_final_d200 200 24 1
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_suite_after_ncu.log 2>&1
echo "pytest -m gpu -x: exit $?" | tee -a gpurun_out/ncu_captures_summary.txt
tail -4 gpurun_out/gpu_suite_after_ncu.log
du -sh gpurun_out
