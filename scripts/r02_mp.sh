#!/bin/bash
# multi-GPU call: N = $1 ranks on one box.  Parity (tests/mp_check.py, in-band protocol), then the bench in lockstep mode
# with the in-band protocol (default) and with data + flag for comparison.  Everything under a short `timeout`.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -z "$SKIP_CHECK" ]; then
  MIMSEM_HALO_LL=1 timeout 300 $TR --master-port 29511 tests/mp_check.py > gpurun_out/mp${N}_ll1.log 2>&1; echo "mp_check N=$N ll=1 rc=$?"
  tail -2 gpurun_out/mp${N}_ll1.log
fi
for ll in ${LLS:-1 0}; do
  MIMSEM_HALO_LL=$ll timeout 240 $TR --master-port 29512 bench.py --gpus $N --steps 50 --warmup 10 --no-e2e > gpurun_out/bench_n${N}_ll$ll.json 2> gpurun_out/bench_n${N}_ll$ll.err; echo "bench N=$N ll=$ll rc=$?"
  cut -c1-1500 gpurun_out/bench_n${N}_ll$ll.json; grep -v "^W\|^\*\|OMP_NUM" gpurun_out/bench_n${N}_ll$ll.err | tail -3
done
