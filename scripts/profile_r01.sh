#!/bin/bash
# ncu evidence for the round: launch list + full capture of the M1 and K tile kernels (1 GPU).
set -o pipefail
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph"
$B > gpurun_out/plain_m1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r01b_launches_m1.csv $B > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_apply_m1_tma -s 3 -c 1 -f -o gpurun_out/prof_m1_r01b $B > gpurun_out/ncu_m1.log 2>&1
$B --op K > gpurun_out/plain_k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_apply_k_tma -s 3 -c 1 -f -o gpurun_out/prof_k_r01b $B --op K > gpurun_out/ncu_k.log 2>&1
ls -la gpurun_out/*.ncu-rep
