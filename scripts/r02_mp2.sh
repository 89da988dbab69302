#!/bin/bash
# bench only, sweep of the push-CTA cap: scripts/r02_mp2.sh N cap1 cap2 ...
N=$1; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for cap in "$@"; do
  MIMSEM_PUSH_CTAS=$cap timeout 200 $TR --master-port 29512 bench.py --gpus $N --steps 54 --warmup 10 --no-e2e --no-sustained > gpurun_out/r02_bench_n${N}_cap$cap.json 2> gpurun_out/r02_bench_n${N}_cap$cap.err; echo "bench N=$N cap=$cap rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_n${N}_cap$cap.json") if l.startswith("{")][-1])
print("cap $cap", d["n_gpus"], round(d["value"],1), d["ms_per_step"], d["parity_check"]["bitwise_vs_1gpu"], round(d["dependent_applies"]["value"],1), "pipelined", d.get("pipelined",{}).get("value"))
PY
done
