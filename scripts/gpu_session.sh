#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list + full capture of the hot kernels.
# usage: gpurun --timeout 1800 -- 'bash scripts/gpu_session.sh [tests|bench|bench3|ncu|ncu3|multi N|all]'
set -u
what=${1:-all}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $out/gpu.txt 2>&1
if [[ $what == all || $what == tests ]]; then
  timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 > $out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $out/pytest_gpu.log
  tail -25 $out/pytest_gpu.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit $?" >> $out/smoke.log; tail -3 $out/smoke.log
fi
if [[ $what == all || $what == bench ]]; then
  timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench exit $?"; cat $out/bench.json; tail -5 $out/bench.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err; cat $out/bench_ref.json
  for w in C1 C2 C5; do
    timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --cpu-sample 32 > $out/bench_$w.json 2> $out/bench_$w.err; echo "bench $w exit $?"; cat $out/bench_$w.json; tail -3 $out/bench_$w.err
  done
  timeout 600 python bench.py --correlated 0.25 --steps 10 --warmup 3 --cpu-sample 0 > $out/bench_C4_corr.json 2> $out/bench_C4_corr.err; echo "bench corr exit $?"; cat $out/bench_C4_corr.json; tail -3 $out/bench_C4_corr.err
fi
if [[ $what == all || $what == bench3 ]]; then
  timeout 900 python bench.py --workload C3 --steps 10 --warmup 3 > $out/bench_C3.json 2> $out/bench_C3.err; echo "bench C3 exit $?"; cat $out/bench_C3.json; tail -5 $out/bench_C3.err
  timeout 600 python bench.py --workload C3 --impl reference --steps 2 --warmup 1 > $out/bench_C3_ref.json 2> $out/bench_C3_ref.err; cat $out/bench_C3_ref.json
fi
if [[ $what == all || $what == ncu ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-parity --cpu-sample 0 > $out/ncu_launches.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"select_umma|select_q" -s 3 -c 1 -f -o $out/prof_select \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-parity --cpu-sample 0 > $out/ncu_full.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^ap_kernel -s 9 -c 1 -f -o $out/prof_ap \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-parity --cpu-sample 0 > $out/ncu_full_ap.log 2>&1
  ls -la $out
fi
if [[ $what == ncu3 ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches_C3.csv \
      python bench.py --workload C3 --steps 1 --warmup 3 --ref-images 1 > $out/ncu_launches_C3.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tf32 -s 10 -c 7 -f -o $out/prof_conv \
      python bench.py --workload C3 --steps 1 --warmup 3 --ref-images 1 > $out/ncu_full_conv.log 2>&1
  ls -la $out
fi
if [[ $what == multi ]]; then
  n=${2:-2}
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -x > $out/pytest_multi.log 2>&1; echo "pytest multi exit $?" >> $out/pytest_multi.log; tail -8 $out/pytest_multi.log
  for w in C4 C5; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --workload $w --steps 10 --warmup 3 --cpu-sample 0 \
        > $out/bench_${w}_n$n.json 2> $out/bench_${w}_n$n.err; echo "bench $w n=$n exit $?"; cat $out/bench_${w}_n$n.json; tail -5 $out/bench_${w}_n$n.err
  done
fi
