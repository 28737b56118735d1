#!/bin/bash
# round-2 scheduling sweeps of the three component solves (profiles/r02_sweep*.txt); run on an 8-GPU box:
#   gpurun --gpus 8 -- 'bash scripts/gpu_r2_sweeps.sh'
# Every variant runs in one launch of the ranks (scripts/sweep_groups.py); '|' stands for ',' inside a value.
mkdir -p gpurun_out
export NDSM_P2P_TIMEOUT_MS=20000
run() {  # nproc port steps configs tag
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$1" --master-addr 127.0.0.1 --master-port "$2" \
    scripts/sweep_groups.py --steps "$3" --warmup 2 --configs "$4" > "gpurun_out/$5.log" 2> "gpurun_out/$5.err"
  echo "$5 rc=$?"; grep SWEEP "gpurun_out/$5.log" | cut -c1-260
}
# batched groups, halo depth
run 8 29549 3 ";NDSM_COMPONENT_GROUPS=012;NDSM_COMPONENT_GROUPS=01|2;NDSM_COMPONENT_GROUPS=02|1;NDSM_COMPONENT_GROUPS=0|12;NDSM_COMPONENT_GROUPS=012 NDSM_HALO_PLANES=8;NDSM_COMPONENT_GROUPS=012 NDSM_HALO_PLANES=10;NDSM_HALO_PLANES=8;NDSM_COMPONENT_GROUPS=01|2 NDSM_HALO_PLANES=8;NDSM_COMPONENT_GROUPS=0|1|2" sweep_g8
run 4 29550 3 ";NDSM_COMPONENT_GROUPS=012;NDSM_COMPONENT_GROUPS=01|2;NDSM_COMPONENT_GROUPS=0|1|2" sweep_g4
# PDL, split wait, staggered starts, partition depth, graphs
run 8 29551 3 ";NDSM_B200_PDL=0;NDSM_P2P_SPLIT_WAIT=1;NDSM_STAGGER_US=700;NDSM_STAGGER_US=1400;NDSM_STAGGER_US=2500;NDSM_SLAB_MIN_POINTS=1000000;NDSM_P2P_SPLIT_WAIT=1 NDSM_STAGGER_US=1400;NDSM_B200_GRAPH=0;NDSM_P2P_SPLIT_WAIT=1 NDSM_B200_PDL=0;" sweep2_g8
run 4 29552 3 ";NDSM_B200_PDL=0;NDSM_P2P_SPLIT_WAIT=1;NDSM_STAGGER_US=1400;NDSM_STAGGER_US=2500" sweep2_g4
# run-to-run stability of the default (26 consecutive solves, every graph capture reported)
NDSM_B200_TRACE=2 run 2 29555 26 "" sweep6_g2
grep -c "graph capture:" gpurun_out/sweep6_g2.err
