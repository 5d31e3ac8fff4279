#!/bin/bash
# Round-2 second GPU call: the warp-per-shard kernel (cfg.kernel = 6) for the first time on hardware.
set -u
mkdir -p gpurun_out
timeout 300 tools/micro/membench > gpurun_out/membench2.md 2> gpurun_out/membench2.err
echo "membench: exit $?" | tee gpurun_out/call2_summary.txt
W2B_DEFAULT_KERNEL=6 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --deselect tests/test_gpu_parity.py::test_ring_kernel_odd_shapes > gpurun_out/warp_parity.log 2>&1
echo "parity suite with the warp kernel as default: exit $?" | tee -a gpurun_out/call2_summary.txt
tail -15 gpurun_out/warp_parity.log
timeout 900 python tools/warp_sweep.py --out gpurun_out/warp_sweep.md > gpurun_out/warp_sweep.log 2>&1
echo "warp sweep: exit $?" | tee -a gpurun_out/call2_summary.txt
cat gpurun_out/warp_sweep.md
tail -5 gpurun_out/warp_sweep.log
cat gpurun_out/membench2.md
