#!/bin/bash
export MIMSEM_GPU_LIB=$PWD/mimsem_b200/libmimsem_gpu_diag.so
for d in 1 33 64 96 32; do
timeout 300 python scripts/pipe_times.py M1 diag_debug=$d 2>&1 | tail -7
done
