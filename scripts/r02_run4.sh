#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest4.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke4.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke4.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err
tail -6 gpurun_out/r02_pytest4.log; tail -3 gpurun_out/r02_smoke4.log; cut -c1-1800 gpurun_out/r02_bench4.json; tail -3 gpurun_out/r02_bench4.err
