#!/bin/bash
# final profiles: (1) launch list of one V-cycle of every solve at 513^3, (2) full captures of the finest-level kernels
mkdir -p gpurun_out
timeout 300 python scripts/prof_target.py 513 1 > gpurun_out/plain513.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_513_1cycle.csv python scripts/prof_target.py 513 1 > gpurun_out/ncu_launch.log 2>&1
timeout 300 python scripts/prof_transfers.py 513 > gpurun_out/plain_tr.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_relax3d|k_restrict_direct|k_interp_add_zt|k_residual3d|k_diff_partial" -s 6 -c 6 -o gpurun_out/prof_finest_v3 python scripts/prof_transfers.py 513 > gpurun_out/ncu_tr.log 2>&1
tail -n 2 gpurun_out/ncu_launch.log gpurun_out/ncu_tr.log
