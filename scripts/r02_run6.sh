#!/bin/bash
mkdir -p gpurun_out
python scripts/tune_ops.py --op M1 --sweep m1_min_blocks=0,5,6,7 --sweep stream_stores=0,1 > gpurun_out/r02_tune6.jsonl 2> gpurun_out/r02_tune6.err
python scripts/tune_ops.py --op M1h --sweep stream_stores=0,1 >> gpurun_out/r02_tune6.jsonl 2>> gpurun_out/r02_tune6.err
cut -c1-210 gpurun_out/r02_tune6.jsonl; tail -3 gpurun_out/r02_tune6.err
