#!/bin/bash
# 8-GPU call: BASELINE configs[4] (V = 3.7 M, D = 800, 8 x B200 with the 23.7 GB replica average) and the headline
# shape at 8 GPUs.  Short runs: an 8-GPU minute costs eight.
set -u
mkdir -p gpurun_out
: > gpurun_out/call_8gpu_summary.txt
nvidia-smi -L | wc -l | tee -a gpurun_out/call_8gpu_summary.txt
for se in 4 8; do
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$se \
    bench.py --gpus 8 --workload c5 --steps 8 --warmup 3 --sync-every $se > gpurun_out/bench_c5_8gpu_sync$se.json 2> gpurun_out/bench_c5_8gpu_sync$se.err
  echo "bench c5 8 GPUs sync-every $se: exit $?" | tee -a gpurun_out/call_8gpu_summary.txt
  cut -c1-2500 gpurun_out/bench_c5_8gpu_sync$se.json; tail -2 gpurun_out/bench_c5_8gpu_sync$se.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/bench_c2_8gpu.json 2> gpurun_out/bench_c2_8gpu.err
echo "bench c2 8 GPUs: exit $?" | tee -a gpurun_out/call_8gpu_summary.txt
cut -c1-2500 gpurun_out/bench_c2_8gpu.json
