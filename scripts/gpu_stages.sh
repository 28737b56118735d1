#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
for n in 129 257 513; do
  for g in 0 1; do
    NDSM_B200_GRAPH=$g timeout 600 python scripts/time_stages.py $n 5 2>&1 | tee -a gpurun_out/stages.log
  done
done
