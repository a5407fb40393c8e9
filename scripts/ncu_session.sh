#!/bin/bash
# One gpurun call (1 GPU): launch lists + ncu --set full captures of the hot kernels of the metric (C4) and the encoder (C3).
# usage: gpurun --timeout 2400 -- 'bash scripts/ncu_session.sh'
set -u
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-parity --cpu-sample 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02_launches_C4.csv $B > $out/ncu_l1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"select_umma|select_q" -s 3 -c 1 -f -o $out/r02_prof_select $B > $out/ncu_f1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^ap_kernel -s 6 -c 1 -f -o $out/r02_prof_ap $B > $out/ncu_f2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hist_kernel -s 6 -c 1 -f -o $out/r02_prof_hist $B > $out/ncu_f3.log 2>&1
E="python bench.py --workload C3 --steps 1 --warmup 3 --ref-images 1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/r02_launches_C3.csv $E > $out/ncu_l2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv1_stage -s 3 -c 1 -f -o $out/r02_prof_stage1 $E > $out/ncu_f4.log 2>&1
# conv2 (first group), conv3, fc6 of the 4th encode call
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tf32 -s 30 -c 10 -f -o $out/r02_prof_conv $E > $out/ncu_f5.log 2>&1
ls -la $out/*.ncu-rep $out/r02_launches_*.csv
