#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 30 gpurun_out/pytest_gpu.log
rm -f gpurun_out/stages.log
for n in 129 513; do timeout 600 python scripts/time_stages.py $n 3 2>&1 | tee -a gpurun_out/stages.log; done
timeout 600 python bench.py --size 513 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench513.json 2> gpurun_out/bench513.err
