#!/bin/bash
# first GPU contact: smoke, parity tests, short benches
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; which gfortran >> gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --size 129 --steps 2 --warmup 1 > gpurun_out/bench129.json 2> gpurun_out/bench129.err
timeout 600 python bench.py --size 513 --steps 2 --warmup 1 > gpurun_out/bench513.json 2> gpurun_out/bench513.err
tail -5 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench129.json gpurun_out/bench513.json gpurun_out/*.err
