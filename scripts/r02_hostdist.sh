#!/bin/bash
# C++ host layer on N GPUs (N plain processes, file rendezvous): all fifteen operators, the partitioned solve and bursts of
# fused M1 launches, bitwise vs 1 GPU; then the M1 apply of the benchmark shape timed in stream order and as bursts.
#   gpurun --gpus 2 -- 'bash scripts/r02_hostdist.sh 2'
N=${1:-2}
mkdir -p gpurun_out
exe=mimsem_b200/host/build/host_dist_check
for cfg in "sphere 3 6 30" "box 3 6 40" "sphere 4 6 60"; do
  rdv=$(mktemp -d)
  for r in $(seq 0 $((N-1))); do
    MIMSEM_RANK=$r MIMSEM_WORLD=$N timeout 120 $exe $cfg $rdv > gpurun_out/hostdist_${N}_${cfg// /_}_r$r.log 2>&1 &
  done
  wait
  echo "== $cfg"; cat gpurun_out/hostdist_${N}_${cfg// /_}_r0.log; tail -n 2 gpurun_out/hostdist_${N}_${cfg// /_}_r1.log
done
rdv=$(mktemp -d)
for r in $(seq 0 $((N-1))); do
  MIMSEM_RANK=$r MIMSEM_WORLD=$N timeout 150 $exe sphere 4 48 60 $rdv time > gpurun_out/hostdist_${N}_time_r$r.log 2>&1 &
done
wait
echo "== timing, C5"; cat gpurun_out/hostdist_${N}_time_r0.log; tail -n 2 gpurun_out/hostdist_${N}_time_r1.log
grep '^{' gpurun_out/hostdist_${N}_time_r0.log > gpurun_out/r02_hostdist_time_${N}gpu.json
