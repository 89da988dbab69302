#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest3.log
python scripts/tune_ops.py --op M1 --sweep m1_min_blocks=0,4 --sweep prefetch_ahead=296,444,592 > gpurun_out/r02_tune3_m1.jsonl 2> gpurun_out/r02_tune3.err
python scripts/tune_ops.py --op M1h --sweep m1_min_blocks=0,5 --sweep prefetch_ahead=296,444 > gpurun_out/r02_tune3_m1h.jsonl 2>> gpurun_out/r02_tune3.err
tail -5 gpurun_out/r02_pytest3.log; cat gpurun_out/r02_tune3_m1.jsonl gpurun_out/r02_tune3_m1h.jsonl | cut -c1-230
