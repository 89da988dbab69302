#!/bin/bash
# multi-GPU call: N = $1 ranks on one box.  Parity (tests/mp_check.py), then the bench (fused ghost refresh + M1: graphs of
# consecutive independent steps under programmatic dependent launch = headline, the same launches in strict stream
# order = "dependent_applies").  Everything under a short `timeout`.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -z "$SKIP_CHECK" ]; then
  timeout 300 $TR --master-port 29511 tests/mp_check.py > gpurun_out/mp${N}.log 2>&1; echo "mp_check N=$N rc=$?"
  grep -v "^W\|^\*\|OMP_NUM" gpurun_out/mp${N}.log | tail -4
fi
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 54 --warmup 10 --no-e2e > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err; echo "bench N=$N rc=$?"
cut -c1-2500 gpurun_out/r02_bench_n${N}.json; grep -v "^W\|^\*\|OMP_NUM" gpurun_out/r02_bench_n${N}.err | tail -5
