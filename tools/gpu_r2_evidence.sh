#!/bin/bash
# ncu --set full captures of the final kernels: the fused sweep on the mixed corpus and on config 2, the detokenizer's emit kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${1:-r4ev}
mkdir -p gpurun_out
BLT_DENSE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_sweep --launch-skip 2 -c 1 -f -o gpurun_out/${T}_fused_mixed python tools/kbench.py --bytes 268435456 --iters 1 --variants 3 --configs 9 > gpurun_out/${T}_ncu_fused_mixed.log 2>&1
BLT_DENSE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_sweep --launch-skip 2 -c 1 -f -o gpurun_out/${T}_fused_cfg2 python tools/kbench.py --bytes 268435456 --iters 1 --variants 3 --configs 2 > gpurun_out/${T}_ncu_fused_cfg2.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
