#!/bin/bash
# N-GPU run: NCCL slab correctness worker + bench
NG=${1:-2}
SIZES=${2:-"513"}
mkdir -p gpurun_out
if [ "$3" != "nocheck" ]; then
NDSM_SLAB_MIN_PLANES=8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_worker.py > gpurun_out/multi_worker.log 2>&1; echo "worker rc=$?" >> gpurun_out/multi_worker.log
grep -E "MULTI_GPU_OK|worker rc|Error" gpurun_out/multi_worker.log | tail -n 12
if [ "$NG" -ge 3 ]; then  # the same check with component groups off: all ranks on z-slabs
NDSM_HYBRID=0 NDSM_SLAB_MIN_PLANES=8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29535 tests/multi_gpu_worker.py > gpurun_out/multi_worker_slabs.log 2>&1; echo "slab-only worker rc=$?" >> gpurun_out/multi_worker_slabs.log
grep -E "MULTI_GPU_OK|worker rc|Error" gpurun_out/multi_worker_slabs.log | tail -n 12
fi
fi
for n in $SIZES; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $NG --size $n --steps 3 --warmup 3 > gpurun_out/bench${n}_g$NG.json 2> gpurun_out/bench${n}_g$NG.err
  tail -n 2 gpurun_out/bench${n}_g$NG.err
done
