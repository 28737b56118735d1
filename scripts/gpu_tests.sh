#!/bin/bash
# GPU parity run: smoke + full gpu test-suite (no -x: list every failure)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/smoke.log; tail -n 25 gpurun_out/pytest_gpu.log
