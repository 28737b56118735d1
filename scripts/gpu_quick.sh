#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 30 gpurun_out/pytest_gpu.log
rm -f gpurun_out/stages.log
for n in 129 513; do timeout 600 python scripts/time_stages.py $n 3 2>&1 | tee -a gpurun_out/stages.log; done
NDSM_VIRTUAL_SLABS=4 timeout 600 python scripts/time_stages.py 257 3 2>&1 | tee -a gpurun_out/stages.log
