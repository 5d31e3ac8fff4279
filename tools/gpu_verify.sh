#!/bin/bash
# One-GPU verification pass, as run at the end of round 2 (gpurun -- 'bash tools/gpu_verify.sh'): GPU suite, smoke(),
# bench (both arms; c3 / c4 lines), compute-sanitizer on the production kernel, ncu launch list of the bench command.
# Everything lands in gpurun_out/; the files worth keeping are copied to profiles/ by hand (profiles/README.md).
set -u
mkdir -p gpurun_out
S=gpurun_out/verify_summary.txt
: > $S
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_suite.log 2>&1
echo "pytest -m gpu -x: exit $?" | tee -a $S
tail -2 gpurun_out/gpu_suite.log | tee -a $S
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke: exit $?" | tee -a $S
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench c2: exit $?" | tee -a $S
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench reference arm: exit $?" | tee -a $S
for w in c3 c4; do
  timeout 400 python bench.py --workload $w --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "bench $w: exit $?" | tee -a $S
done
for tool in synccheck memcheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_step.py > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "compute-sanitizer $tool: exit $?" | tee -a $S
  grep -E "^D=|ERROR SUMMARY|done" gpurun_out/sanitizer_$tool.txt | tee -a $S
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/bench_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launch list: exit $?" | tee -a $S
du -sh gpurun_out | tee -a $S
