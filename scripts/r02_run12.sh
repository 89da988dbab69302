#!/bin/bash
mkdir -p gpurun_out/prof
CMD="python scripts/tune_ops.py --op M1 --sweep m1_variant=3 --steps 3"
$CMD > gpurun_out/prof/plain_pipe.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_apply_m1_pipe -s 4 -c 1 -f -o /tmp/prof_pipe $CMD > gpurun_out/prof/ncu_pipe.log 2>&1
ncu -i /tmp/prof_pipe.ncu-rep --page raw --csv > gpurun_out/prof/raw_pipe.csv 2>/dev/null
ncu -i /tmp/prof_pipe.ncu-rep --page source --csv > gpurun_out/prof/source_pipe.csv 2>/dev/null
ls -la gpurun_out/prof | tail -5
