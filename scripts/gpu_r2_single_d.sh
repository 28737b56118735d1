#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_transfers.py 513 > gpurun_out/transfers.log 2>&1; cat gpurun_out/transfers.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench513.json 2> gpurun_out/bench513.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench513.json | cut -c1-400
