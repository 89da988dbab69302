#!/bin/bash
# round 2, GPU call 1: parity of the re-tiled M1 kernel + register-budget / prefetch sweep + per-operator table
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
python scripts/tune_ops.py --op M1 --sweep m1_min_blocks=4,5,6 --sweep prefetch_ahead=0,296,444,740 > gpurun_out/r02_tune_m1.jsonl 2> gpurun_out/r02_tune_m1.err
python scripts/tune_ops.py --op M1h --sweep prefetch_ahead=0,296,444,740 > gpurun_out/r02_tune_m1h.jsonl 2>> gpurun_out/r02_tune_m1.err
python scripts/tune_ops.py --op K --op M2 --op M0 --op E21 --op E12 > gpurun_out/r02_tune_rest.jsonl 2>> gpurun_out/r02_tune_m1.err
tail -5 gpurun_out/r02_pytest1.log; cat gpurun_out/r02_tune_m1.jsonl gpurun_out/r02_tune_m1h.jsonl gpurun_out/r02_tune_rest.jsonl | cut -c1-200
