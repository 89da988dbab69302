#!/bin/bash
python -m pytest tests -m gpu -x -q -k "src or galewsky or upwind or multi_gpu" 2>&1 | tail -3
for op in R_up M0h_up; do python bench.py --workload C2 --op $op --no-cpu-baseline --no-e2e --no-sustained --steps 40 --warmup 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'][:40], '%.4f ms'%d['ms_per_step'])"; done
