#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "variants_agree" > gpurun_out/r02_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest8.log
tail -3 gpurun_out/r02_pytest8.log
timeout 300 python scripts/tune_ops.py --op M1 --op M1h --sweep m1_variant=2,3 > gpurun_out/r02_tune10.jsonl 2> gpurun_out/r02_tune10.err
cut -c1-200 gpurun_out/r02_tune10.jsonl; sort gpurun_out/r02_tune10.err | uniq -c | tail -4
export MIMSEM_GPU_LIB=$PWD/mimsem_b200/libmimsem_gpu_diag.so
timeout 300 python scripts/pipe_times.py M1 2>&1 | tail -7
timeout 300 python scripts/pipe_times.py M1h 2>&1 | tail -7
