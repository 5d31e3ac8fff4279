#!/bin/bash
# Round-2 fourth GPU call: ncu captures of the warp kernel (exported to CSV on the box: the .ncu-rep files are too
# big to travel), the GPU parity suite incl. the double-buffered streaming path, a bench run.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
cap() {  # name D neg bits
  W2B_DEFAULT_KERNEL=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_warp -s 1 -c 1 \
    -o /tmp/$1 python tools/prof_step.py $2 $3 $4 0 1500 > gpurun_out/ncu_$1.log 2>&1
  echo "ncu $1: exit $?" | tee -a gpurun_out/call4_summary.txt
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details > gpurun_out/$1_details.txt 2>/dev/null
  ls -la /tmp/$1.ncu-rep gpurun_out/$1_*
}
: > gpurun_out/call4_summary.txt
cap r02_warp_c2 800 24 1
cap r02_warp_d200 200 24 1
W2B_DEFAULT_KERNEL=6 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --deselect tests/test_gpu_parity.py::test_ring_kernel_odd_shapes > gpurun_out/warp_parity3.log 2>&1
echo "parity suite with the warp kernel as default: exit $?" | tee -a gpurun_out/call4_summary.txt
tail -5 gpurun_out/warp_parity3.log
W2B_DEFAULT_KERNEL=6 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_warp.json 2> gpurun_out/bench_warp.err
echo "bench: exit $?" | tee -a gpurun_out/call4_summary.txt
cat gpurun_out/bench_warp.json; tail -3 gpurun_out/bench_warp.err
du -sh gpurun_out
