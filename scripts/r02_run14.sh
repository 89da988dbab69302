#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest14.log
tail -4 gpurun_out/r02_pytest14.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench14.json 2> gpurun_out/r02_bench14.err; echo "bench rc=$?"; cut -c1-1800 gpurun_out/r02_bench14.json; tail -3 gpurun_out/r02_bench14.err
for op in M1h K M2 M0 E21 E12; do python bench.py --op $op --no-cpu-baseline --no-e2e 2>>gpurun_out/r02_ops14.err; done > gpurun_out/r02_ops14.jsonl
python bench.py --op M1 --no-pdl --no-cpu-baseline --no-e2e >> gpurun_out/r02_ops14.jsonl 2>>gpurun_out/r02_ops14.err
python - <<'PY'
import json
for l in open("gpurun_out/r02_ops14.jsonl"):
    d = json.loads(l)
    print(d["config"]["workload"].split("operator ")[1].split()[0], d["config"]["options"], "%.1f GDOF/s" % d["value"], "%.4f ms" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "sustained %.3f" % d["roofline"].get("sustained_frac", 0))
PY
