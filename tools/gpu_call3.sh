#!/bin/bash
# Round-2 third GPU call: warp kernel with the sentence in shared memory — reduce-depth sweep, ncu capture, parity.
set -u
mkdir -p gpurun_out
timeout 900 python tools/warp_sweep.py --shapes c2,c3,c4,d200 --out gpurun_out/warp_sweep2.md > gpurun_out/warp_sweep2.log 2>&1
echo "warp sweep: exit $?" | tee gpurun_out/call3_summary.txt
cat gpurun_out/warp_sweep2.md
W2B_DEFAULT_KERNEL=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_warp -s 1 -c 1 \
  -o gpurun_out/r02_warp_c2 python tools/prof_step.py 800 24 1 0 1500 > gpurun_out/ncu_c2.log 2>&1
echo "ncu c2: exit $?" | tee -a gpurun_out/call3_summary.txt
W2B_DEFAULT_KERNEL=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:train_warp -s 1 -c 1 \
  -o gpurun_out/r02_warp_d200 python tools/prof_step.py 200 24 1 0 1500 > gpurun_out/ncu_d200.log 2>&1
echo "ncu d200: exit $?" | tee -a gpurun_out/call3_summary.txt
W2B_DEFAULT_KERNEL=6 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --deselect tests/test_gpu_parity.py::test_ring_kernel_odd_shapes > gpurun_out/warp_parity2.log 2>&1
echo "parity suite with the warp kernel as default: exit $?" | tee -a gpurun_out/call3_summary.txt
tail -5 gpurun_out/warp_parity2.log
ls -la gpurun_out/*.ncu-rep
