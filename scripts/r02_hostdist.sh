#!/bin/bash
# C++ host layer on N GPUs: N plain processes, file rendezvous (no MPI, no Python)
N=${1:-2}
mkdir -p gpurun_out
for cfg in "sphere 3 6 30" "sphere 4 6 60" "box 3 6 40"; do
  rm -rf /tmp/mimsem_rdv; mkdir -p /tmp/mimsem_rdv
  pids=""
  for r in $(seq 0 $((N-1))); do
    MIMSEM_RANK=$r MIMSEM_WORLD=$N timeout 300 mimsem_b200/host/build/host_dist_check $cfg /tmp/mimsem_rdv > gpurun_out/hostdist_n${N}_r$r.log 2>&1 &
    pids="$pids $!"
  done
  rc=0; for p in $pids; do wait $p || rc=1; done
  echo "== $cfg on $N GPUs rc=$rc"; cat gpurun_out/hostdist_n${N}_r0.log; for r in $(seq 1 $((N-1))); do grep -i "error\|rank" gpurun_out/hostdist_n${N}_r$r.log | head -3; done
done
