#!/bin/bash
# round 2, GPU call: parity after merging the two direction code paths + new operators; tuning sweeps; M1 ncu
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest2.log
python scripts/tune_ops.py --op M1 --sweep m1_min_blocks=4,0,6 --sweep prefetch_ahead=0,444 > gpurun_out/r02_tune2_m1.jsonl 2> gpurun_out/r02_tune2.err
python scripts/tune_ops.py --op M1h --sweep m1_min_blocks=0,5 --sweep prefetch_ahead=0,444 > gpurun_out/r02_tune2_m1h.jsonl 2>> gpurun_out/r02_tune2.err
python scripts/tune_ops.py --op K > gpurun_out/r02_tune2_k.jsonl 2>> gpurun_out/r02_tune2.err
OPS="M1 M1h" bash scripts/r02_profile.sh > /dev/null 2>&1
tail -5 gpurun_out/r02_pytest2.log; cat gpurun_out/r02_tune2_m1.jsonl gpurun_out/r02_tune2_m1h.jsonl gpurun_out/r02_tune2_k.jsonl | cut -c1-230
