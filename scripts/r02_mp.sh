#!/bin/bash
# multi-GPU call: N = $1 ranks on one box.  Parity (tests/mp_check.py, both hand-over protocols), then the bench in lockstep
# mode with the in-band protocol (default) and with data + flag for comparison.  Everything under `timeout`.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for ll in 1 0; do
  MIMSEM_HALO_LL=$ll timeout 600 $TR --master-port 29511 tests/mp_check.py > gpurun_out/mp${N}_ll$ll.log 2>&1; echo "mp_check N=$N ll=$ll rc=$?"
  tail -3 gpurun_out/mp${N}_ll$ll.log
done
for ll in 1 0; do
  MIMSEM_HALO_LL=$ll timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 50 --warmup 10 --no-e2e > gpurun_out/bench_n${N}_ll$ll.json 2> gpurun_out/bench_n${N}_ll$ll.err; echo "bench N=$N ll=$ll rc=$?"
  cut -c1-400 gpurun_out/bench_n${N}_ll$ll.json; tail -3 gpurun_out/bench_n${N}_ll$ll.err
done
