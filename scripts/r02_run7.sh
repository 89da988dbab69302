#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest7.log
python scripts/tune_ops.py --op M1 --sweep m1_min_blocks=0,5,6 --sweep prefetch_ahead=296,444 > gpurun_out/r02_tune7.jsonl 2> gpurun_out/r02_tune7.err
python scripts/tune_ops.py --op M1h --sweep m1_min_blocks=0,5 >> gpurun_out/r02_tune7.jsonl 2>> gpurun_out/r02_tune7.err
tail -6 gpurun_out/r02_pytest7.log; cut -c1-200 gpurun_out/r02_tune7.jsonl; tail -3 gpurun_out/r02_tune7.err
