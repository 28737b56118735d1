#!/bin/bash
# iteration loop on the GPU box: parity tests, benches, optional ncu
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
for n in 129 257 513; do
  timeout 900 python bench.py --size $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench$n.json 2> gpurun_out/bench$n.err || tail -n 5 gpurun_out/bench$n.err
done
if [ "$1" = "ncu" ]; then
  timeout 300 python scripts/prof_target.py 129 > gpurun_out/plain129.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/launches_129.csv python scripts/prof_target.py 129 > gpurun_out/ncu_launch.log 2>&1
  timeout 300 python scripts/prof_target.py 513 1 > gpurun_out/plain513.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_relax3d<\(bool\)0>|k_relax3d<false>|k_relax3d<0>' -s 20 -c 2 -o gpurun_out/prof_relax3d python scripts/prof_target.py 513 1 > gpurun_out/ncu_full.log 2>&1
  tail -n 3 gpurun_out/ncu_launch.log gpurun_out/ncu_full.log
fi
