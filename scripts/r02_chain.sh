#!/bin/bash
# dependent chain M1h -> solve_M1 -> E21 on N GPUs (stream order): scripts/r02_chain.sh N
N=${1:-1}
mkdir -p gpurun_out
if [ "$N" = 1 ]; then
  timeout 300 python bench.py --dependent-chain --steps 5 --warmup 2 > gpurun_out/r02_chain_n1.json 2> gpurun_out/r02_chain_n1.err
else
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --dependent-chain --steps 5 --warmup 2 > gpurun_out/r02_chain_n$N.json 2> gpurun_out/r02_chain_n$N.err
fi
echo "rc=$?"; grep "^{" gpurun_out/r02_chain_n$N.json | cut -c1-400; grep -v "^W\|^\*\|OMP_NUM" gpurun_out/r02_chain_n$N.err | tail -4
