#!/bin/bash
# ncu evidence (1 GPU): launch list of the bench command + one full capture per operator kernel on C5.
# Each ncu command runs only after the same command exited 0 without ncu.  The .ncu-rep files are summarised on the
# box (raw page -> csv) because gpurun_out/ returns at most 64 MiB.
mkdir -p gpurun_out/prof
rm -f gpurun_out/prof/raw_*.csv gpurun_out/prof/details_*.csv
# launch list of the default bench command (graphs of consecutive steps, programmatic dependent launch)
BL="python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-sustained"
$BL > gpurun_out/prof/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/prof/launches_M1.csv $BL > gpurun_out/prof/ncu_launches.log 2>&1
# one full capture per operator: eager launches (no graph) so that -s / -c count launches of the kernel itself
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-sustained"
for op in ${OPS:-M1 M1tile M1h K M2 M0 E21 E12}; do
  extra=""; bop=$op
  [ $op = M1 ] && extra="--opt m1_variant=3"
  [ $op = M1tile ] && bop=M1
  $B --op $bop $extra > gpurun_out/prof/plain_$op.log 2>&1 || continue
  ncu --set full --clock-control none --import-source on -k "regex:k_apply|k_inc_tile" -s 3 -c 1 -f -o /tmp/prof_$op $B --op $bop $extra > gpurun_out/prof/ncu_$op.log 2>&1
  ncu -i /tmp/prof_$op.ncu-rep --page raw --csv > gpurun_out/prof/raw_$op.csv 2>/dev/null
  ncu -i /tmp/prof_$op.ncu-rep --page details --csv > gpurun_out/prof/details_$op.csv 2>/dev/null
done
ncu -i /tmp/prof_M1.ncu-rep --page source --csv > gpurun_out/prof/source_M1.csv 2>/dev/null
ls -la gpurun_out/prof | head -40
