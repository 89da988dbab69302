#!/bin/bash
# per-operator bench table on C5 (one line per operator) -> gpurun_out/ops_table.jsonl
out=gpurun_out/ops_table.jsonl
: > $out
for op in M1 M1h M2 M0 K E21 E12; do
  python bench.py --op $op --steps 20 --warmup 5 --no-cpu-baseline --no-e2e >> $out 2>> gpurun_out/ops_table.err
done
python - <<'PY'
import json
for l in open("gpurun_out/ops_table.jsonl"):
    d = json.loads(l)
    print(d["config"]["workload"].split("operator ")[1].split()[0], "%.1f GDOF/s" % d["value"], "%.3f ms" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"])
PY
