#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1700 python -m pytest tests -m gpu -q --timeout=1200 -x -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/smoke.log; tail -n 12 gpurun_out/pytest_gpu.log
rm -f gpurun_out/stages.log
for pdl in 0 1; do
  echo "== NDSM_B200_PDL=$pdl" >> gpurun_out/stages.log
  NDSM_B200_PDL=$pdl NDSM_B200_TRACE=1 timeout 600 python scripts/time_stages.py 513 3 >> gpurun_out/stages.log 2>&1
  NDSM_B200_PDL=$pdl NDSM_B200_TRACE=1 timeout 600 python scripts/time_stages.py 129 3 >> gpurun_out/stages.log 2>&1
done
grep -E "^==|^n=|chi V-cycle" gpurun_out/stages.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench513.json 2> gpurun_out/bench513.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench513.json | cut -c1-400
