#!/bin/bash
# bench at 129/257/513 + ncu launch list and one full capture of the dominant kernel (after a plain run exited 0)
mkdir -p gpurun_out
N=${1:-513}
timeout 600 python bench.py --size 129 --steps 3 --warmup 3 > gpurun_out/bench129.json 2> gpurun_out/bench129.err
timeout 600 python bench.py --size 257 --steps 3 --warmup 3 > gpurun_out/bench257.json 2> gpurun_out/bench257.err
timeout 900 python bench.py --size 513 --steps 3 --warmup 3 > gpurun_out/bench513.json 2> gpurun_out/bench513.err
timeout 900 python bench.py --impl reference --size 513 --steps 2 --warmup 1 > gpurun_out/bench513_ref.json 2> gpurun_out/bench513_ref.err
# profile target: one device-resident 257^3 solve (short enough for ncu)
timeout 300 python scripts/prof_target.py 257 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_257.csv python scripts/prof_target.py 257 > gpurun_out/ncu_launch.log 2>&1
timeout 300 python scripts/prof_target.py 513 1 > gpurun_out/plain513.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_relax3d -s 40 -c 2 -o gpurun_out/prof_relax3d python scripts/prof_target.py 513 1 > gpurun_out/ncu_full.log 2>&1
tail -n 2 gpurun_out/*.err gpurun_out/ncu_launch.log gpurun_out/ncu_full.log
ls -la gpurun_out
