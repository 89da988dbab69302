#!/bin/bash
# C++ host layer on N GPUs (N plain processes, file rendezvous): all fifteen operators + the partitioned solve, bitwise vs 1 GPU.
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
