#!/bin/bash
mkdir -p gpurun_out
MIMSEM_PIPE_VERBOSE=1 timeout 300 python scripts/tune_ops.py --op M1 --op M1h --sweep m1_variant=3 --steps 3 2>&1 | grep -v "^{" | sort | uniq -c | head
