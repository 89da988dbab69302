#!/bin/bash
N=$1
mkdir -p gpurun_out
if [ -n "$2" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mp_check.py > gpurun_out/mp${N}.log 2>&1; echo "mp_check N=$N rc=$?"
grep -v "^W\|^\*\|OMP_NUM" gpurun_out/mp${N}.log | tail -2
fi
bash scripts/r02_chain.sh $N
