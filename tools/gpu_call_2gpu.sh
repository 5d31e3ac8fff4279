#!/bin/bash
# 2-GPU call: multi-GPU parity tests (NCCL replica average, CLI on two GPUs) and a short 2-rank bench line.
set -u
mkdir -p gpurun_out
: > gpurun_out/call_2gpu_summary.txt
nvidia-smi -L | tee -a gpurun_out/call_2gpu_summary.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -rA > gpurun_out/multi_tests.log 2>&1
echo "multi-GPU tests: exit $?" | tee -a gpurun_out/call_2gpu_summary.txt
grep -E "passed|failed|skipped|MGPU_OK|cli 1 vs 2|^FAILED|^ERROR" gpurun_out/multi_tests.log | tail
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench 2 GPUs: exit $?" | tee -a gpurun_out/call_2gpu_summary.txt
cat gpurun_out/bench_2gpu.json | cut -c1-3000
tail -3 gpurun_out/bench_2gpu.err
