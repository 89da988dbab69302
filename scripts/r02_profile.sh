#!/bin/bash
# ncu evidence (1 GPU): launch list of the bench command + one full capture per operator kernel on C5.
# Each ncu command runs only after the same command exited 0 without ncu.  The .ncu-rep files are summarised on the
# box (raw page -> csv) because gpurun_out/ returns at most 64 MiB; only the M1 report itself comes back.
mkdir -p gpurun_out/prof
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-sustained"
$B > gpurun_out/prof/plain_M1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/prof/launches_M1.csv $B > gpurun_out/prof/ncu_launches.log 2>&1
for op in ${OPS:-M1 M1h K M2 M0 E21 E12}; do
  $B --op $op > gpurun_out/prof/plain_$op.log 2>&1 || continue
  ncu --set full --clock-control none --import-source on -k regex:k_apply -s 3 -c 1 -f -o /tmp/prof_$op $B --op $op > gpurun_out/prof/ncu_$op.log 2>&1
  ncu -i /tmp/prof_$op.ncu-rep --page raw --csv > gpurun_out/prof/raw_$op.csv 2>/dev/null
  ncu -i /tmp/prof_$op.ncu-rep --page details --csv > gpurun_out/prof/details_$op.csv 2>/dev/null
done
cp /tmp/prof_M1.ncu-rep gpurun_out/prof/ 2>/dev/null
ncu -i /tmp/prof_M1.ncu-rep --page source --csv > gpurun_out/prof/source_M1.csv 2>/dev/null
ncu -i /tmp/prof_M1h.ncu-rep --page source --csv > gpurun_out/prof/source_M1h.csv 2>/dev/null
ls -la gpurun_out/prof
