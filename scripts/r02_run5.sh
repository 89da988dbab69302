#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest5.log
python scripts/tune_ops.py --op M2 --sweep m2_variant=0,1 > gpurun_out/r02_tune5.jsonl 2> gpurun_out/r02_tune5.err
python scripts/tune_ops.py --op E21 --op E12 --sweep inc_variant=0,1 >> gpurun_out/r02_tune5.jsonl 2>> gpurun_out/r02_tune5.err
tail -6 gpurun_out/r02_pytest5.log; cut -c1-200 gpurun_out/r02_tune5.jsonl; tail -3 gpurun_out/r02_tune5.err
