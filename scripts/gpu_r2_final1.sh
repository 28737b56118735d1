#!/bin/bash
# round-2 single-GPU evidence run: smoke, full GPU suite (with the unmodified reference ndsm.py when handed in),
# default bench, reference arm, ncu launch list + full capture of the finest-level kernels
mkdir -p gpurun_out
if [ -n "$NDSM_REFPY_B64" ]; then
  mkdir -p /tmp/refpy && echo "$NDSM_REFPY_B64" | base64 -d > /tmp/refpy/ndsm.py && export NDSM_REFERENCE_DIR=/tmp/refpy
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1700 python -m pytest tests -m gpu -q --timeout=1200 -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/smoke.log; tail -n 8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench513.json 2> gpurun_out/bench513.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench513.json | cut -c1-500
timeout 300 python scripts/prof_target.py 513 1 > gpurun_out/plain513.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_513_1cycle.csv python scripts/prof_target.py 513 1 > gpurun_out/ncu_launch.log 2>&1
timeout 300 python scripts/prof_transfers.py 513 > gpurun_out/plain_tr.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_relax3d|k_restrict_direct|k_interp_add_zt|k_residual3d|k_diff_partial" -s 6 -c 6 -o gpurun_out/prof_finest_r02 python scripts/prof_transfers.py 513 > gpurun_out/ncu_tr.log 2>&1
ncu -i gpurun_out/prof_finest_r02.ncu-rep --page raw --csv > gpurun_out/prof_finest_r02_raw.csv 2>/dev/null
tail -n 2 gpurun_out/ncu_launch.log gpurun_out/ncu_tr.log
