#!/bin/bash
# Round-2 fifth GPU call: scatter path A/B (bulk reduce vs red.global.add.v4), membench with the LSU scatter,
# whole GPU suite on the refactored default, ncu captures (exported to CSV on the box).
set -u
mkdir -p gpurun_out
: > gpurun_out/call5_summary.txt
timeout 300 tools/micro/membench > gpurun_out/membench3.md 2> gpurun_out/membench3.err
echo "membench: exit $?" | tee -a gpurun_out/call5_summary.txt
grep -E "train-mix|uniform" gpurun_out/membench3.md
for red in 0 1; do
  W2B_WARP_RED=$red timeout 600 python tools/warp_sweep.py --shapes c2,c3,c4,d200 --configs 0:0:0:1,0:1:0:1 --out gpurun_out/warp_sweep_red$red.md > gpurun_out/warp_sweep_red$red.log 2>&1
  echo "sweep red=$red: exit $?" | tee -a gpurun_out/call5_summary.txt
  cat gpurun_out/warp_sweep_red$red.md
done
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/parity5.log 2>&1
echo "GPU parity suite: exit $?" | tee -a gpurun_out/call5_summary.txt
tail -8 gpurun_out/parity5.log
cap() {  # name D neg bits red
  W2B_WARP_RED=$5 timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_warp -s 1 -c 1 \
    -o /tmp/$1 python tools/prof_step.py $2 $3 $4 0 1500 > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1: exit $?" | tee -a gpurun_out/call5_summary.txt
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
}
cap r02_warp_d200 200 24 1 0
cap r02_warp_c2_red 800 24 1 1
du -sh gpurun_out
