#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/tune_ops.py --op M1 --sweep m1_variant=2,3 --sweep pdl_independent=0,1 --burst 12 --steps 48 > gpurun_out/r02_tune13.jsonl 2> gpurun_out/r02_tune13.err
timeout 300 python scripts/tune_ops.py --op M1h --sweep pdl_independent=0,1 --burst 12 --steps 48 >> gpurun_out/r02_tune13.jsonl 2>> gpurun_out/r02_tune13.err
timeout 300 python scripts/tune_ops.py --workload C5_eighth --op M1 --sweep m1_variant=2,3 --sweep pdl_independent=0,1 --burst 12 --steps 48 >> gpurun_out/r02_tune13.jsonl 2>> gpurun_out/r02_tune13.err
cut -c1-220 gpurun_out/r02_tune13.jsonl; tail -3 gpurun_out/r02_tune13.err
