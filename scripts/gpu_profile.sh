#!/bin/bash
# final profiles: (1) launch list of one V-cycle of every solve at 513^3, (2) full captures of the finest-level kernels
mkdir -p gpurun_out
timeout 300 python scripts/prof_target.py 513 1 > gpurun_out/plain513.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_513_1cycle.csv python scripts/prof_target.py 513 1 > gpurun_out/ncu_launch.log 2>&1
timeout 300 python scripts/prof_target.py 513 1 > gpurun_out/plain513b.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_relax3d<\(bool\)0>|k_relax3d<false>|k_relax3d<0>' -s 20 -c 2 -o gpurun_out/prof_relax3d_v2 python scripts/prof_target.py 513 1 > gpurun_out/ncu_full.log 2>&1
timeout 300 python scripts/prof_transfers.py 513 > gpurun_out/plain_tr.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_restrict_sep|k_interp_add_tiled|k_residual3d|k_diff_partial" -c 6 -o gpurun_out/prof_transfers_v2 python scripts/prof_transfers.py 513 > gpurun_out/ncu_tr.log 2>&1
tail -n 2 gpurun_out/ncu_launch.log gpurun_out/ncu_full.log gpurun_out/ncu_tr.log
