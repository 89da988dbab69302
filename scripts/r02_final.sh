#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_final.log
tail -4 gpurun_out/r02_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02_bench_final.json; tail -3 gpurun_out/r02_bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
