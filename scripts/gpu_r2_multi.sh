#!/bin/bash
# round-2 N-GPU run: slab correctness worker (peer-memory transport, then NCCL fallback when $3 = both), bench
NG=${1:-2}
SIZES=${2:-"513"}
MODE=${3:-p2p}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_n$NG.txt 2>&1
NDSM_SLAB_MIN_PLANES=8 NDSM_P2P_TIMEOUT_MS=20000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_worker.py > gpurun_out/multi_worker_n$NG.log 2>&1; echo "worker rc=$?" >> gpurun_out/multi_worker_n$NG.log
grep -E "MULTI_GPU_OK|worker rc|Error|ERROR|WARNING" gpurun_out/multi_worker_n$NG.log | tail -n 24
if [ "$MODE" = "both" ]; then
NDSM_P2P=0 NDSM_SLAB_MIN_PLANES=8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29535 tests/multi_gpu_worker.py > gpurun_out/multi_worker_nccl_n$NG.log 2>&1; echo "nccl worker rc=$?" >> gpurun_out/multi_worker_nccl_n$NG.log
grep -E "MULTI_GPU_OK|worker rc|Error|ERROR" gpurun_out/multi_worker_nccl_n$NG.log | tail -n 12
fi
for n in $SIZES; do
  NDSM_P2P_TIMEOUT_MS=20000 NDSM_B200_TRACE=${TRACE:-0} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $NG --size $n --steps 3 --warmup 3 > gpurun_out/bench${n}_g$NG.json 2> gpurun_out/bench${n}_g$NG.err
  echo "bench rc=$?"
  tail -n 30 gpurun_out/bench${n}_g$NG.err | cut -c1-300
  python scripts/show_bench.py gpurun_out/bench${n}_g$NG.json 2>/dev/null | head -40
done
