#!/bin/bash
# BASELINE configs C1-C4 (+ the C5 chain): single applies of the operators each config names and the
# diagnose chain (4 x M1h + E21 in one CUDA graph) -> gpurun_out/r02_workloads.jsonl
out=gpurun_out/r02_workloads.jsonl; : > $out; : > gpurun_out/r02_workloads.err
run() { python bench.py --no-cpu-baseline --no-e2e --no-sustained --steps 40 --warmup 10 "$@" >> $out 2>> gpurun_out/r02_workloads.err; }
for op in M1 M2 M0; do run --workload C1 --op $op; done
for op in M1h K R_up M0h_up; do run --workload C2 --op $op; done
for op in M1 M1h K M2; do run --workload C3 --op $op; done
for op in M1 M1h K M2; do run --workload C4 --op $op; done
for w in C2 C3 C4 C5; do run --workload $w --chain; done
python - <<'PY'
import json
for l in open("gpurun_out/r02_workloads.jsonl"):
    d = json.loads(l)
    w = d["config"]["workload"]
    print(w.split(":")[0], ("chain" if "chain" in w else w.split("operator ")[1].split()[0]).ljust(7), "%9.2f GDOF/s" % d["value"], "%8.4f ms" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"])
PY
tail -3 gpurun_out/r02_workloads.err
