#!/bin/bash
# round-2 single-GPU run: smoke, full GPU suite (+ the opt-in fused-mean test, + the unmodified reference ndsm.py
# when its text is handed in through $NDSM_REFPY_B64), host-trace of the stages, default bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
if [ -n "$NDSM_REFPY_B64" ]; then
  mkdir -p /tmp/refpy && echo "$NDSM_REFPY_B64" | base64 -d > /tmp/refpy/ndsm.py && export NDSM_REFERENCE_DIR=/tmp/refpy
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
NDSM_RUN_EXPERIMENTAL=1 timeout 1700 python -m pytest tests -m gpu -q --timeout=1200 -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/smoke.log; tail -n 25 gpurun_out/pytest_gpu.log
rm -f gpurun_out/stages.log
for fm in 0 1; do
  echo "== NDSM_B200_FUSED_MEAN=$fm" >> gpurun_out/stages.log
  NDSM_B200_FUSED_MEAN=$fm NDSM_B200_TRACE=1 timeout 600 python scripts/time_stages.py 513 3 >> gpurun_out/stages.log 2>&1
done
NDSM_B200_TRACE=1 timeout 600 python scripts/time_stages.py 129 3 >> gpurun_out/stages.log 2>&1
tail -n 60 gpurun_out/stages.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench513.json 2> gpurun_out/bench513.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench513.json
