#!/bin/bash
for ch in 2 4 6 10 12 20 30; do python bench.py --opt host_chunk=$ch --no-cpu-baseline --no-sustained --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('host_chunk $ch e2e %.3f GDOF/s %.3f ms'%(d['e2e']['value'], d['e2e']['ms_per_step']))"; done
