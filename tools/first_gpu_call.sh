#!/bin/bash
# First GPU call of a round (run through gpurun from the repo root):
#   gpurun --timeout 4500 -- 'bash tools/first_gpu_call.sh'
# 1. the opt-in tests of the experimental ring-kernel variants (cfg.kernel 2..5), each pytest under its own
#    timeout so that a wedged kernel ends the step instead of the box;
# 2. the A/B sweep of the variants against the measured default on the BASELINE shapes (CUDA events);
# 3. a short bench run of the default for reference;
# 4. the ncu launch list of the bench command at the default step size.
# Everything lands in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
export W2B_TEST_EXPERIMENTAL=1
ok=0  # the sweep only times variants whose tests passed (a variant that hangs would cost the whole sweep)
for v in 2 5 3 4; do
  W2B_VARIANTS=$v timeout 900 python -m pytest tests/test_gpu_variant.py -m gpu -x -q \
    > gpurun_out/variant_tests_k$v.log 2>&1
  rc=$?
  echo "variant $v tests: exit $rc" | tee -a gpurun_out/first_call_summary.txt
  [ $rc -eq 0 ] && ok="$ok,$v"
done
timeout 1500 python tools/variant_sweep.py --kernels "$ok" --out gpurun_out/variants.md > gpurun_out/variant_sweep.log 2>&1
echo "variant sweep: exit $?" | tee -a gpurun_out/first_call_summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench: exit $?" | tee -a gpurun_out/first_call_summary.txt
# 4. launch list of the bench command at its current default step size (profiles/r01_bench_launches.csv was
#    taken with 16 384-word steps); numbers printed under ncu are not bench values
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
  --log-file gpurun_out/launches_default_step.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
  > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launch list: exit $?" | tee -a gpurun_out/first_call_summary.txt
cat gpurun_out/variants.md 2>/dev/null
tail -3 gpurun_out/variant_tests_k*.log
