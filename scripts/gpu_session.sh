#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list + full capture of the hot kernel.
# usage: gpurun --timeout 1800 -- 'bash scripts/gpu_session.sh [tests|bench|ncu|all]'
set -u
what=${1:-all}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $out/gpu.txt 2>&1
if [[ $what == all || $what == tests ]]; then
  timeout 1200 python -m pytest tests -q -m gpu -x --durations=15 > $out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $out/pytest_gpu.log
  tail -25 $out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log; tail -3 $out/smoke.log
fi
if [[ $what == all || $what == bench ]]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench exit $?"; cat $out/bench.json; tail -5 $out/bench.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; cat $out/bench_ref.json
fi
if [[ $what == all || $what == ncu ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --cpu-sample 0 > $out/ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:select_umma -s 3 -c 1 -f -o $out/prof_select \
      python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 > $out/ncu_full.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:ap_kernel -s 9 -c 1 -f -o $out/prof_ap \
      python bench.py --steps 1 --warmup 3 --no-e2e --cpu-sample 0 > $out/ncu_full_ap.log 2>&1
  ls -la $out
fi
