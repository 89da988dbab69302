#!/bin/bash
# Round-end check on one B200: GPU test suite, smoke(), the default bench line, the reference arm, the MatShell path.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_final.log
tail -4 gpurun_out/r02_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02_bench_final.json; tail -3 gpurun_out/r02_bench_final.err
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c1-300
timeout 120 mimsem_b200/host/build/host_shell_bench 4 48 60 2 > gpurun_out/r02_shell_bench.json 2> gpurun_out/r02_shell_bench.err; echo "shell bench rc=$?"; cat gpurun_out/r02_shell_bench.json
