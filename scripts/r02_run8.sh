#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "variants_agree" > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest8.log
timeout 300 python scripts/tune_ops.py --op M1 --op M1h --sweep m1_variant=2,3 --sweep prefetch_ahead=0,444 > gpurun_out/r02_tune8.jsonl 2> gpurun_out/r02_tune8.err
tail -6 gpurun_out/r02_pytest8.log; cut -c1-200 gpurun_out/r02_tune8.jsonl; tail -3 gpurun_out/r02_tune8.err
