#!/bin/bash
# round-2 GPU session A: parity of the fused sweep, kernel timings, one ncu capture
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2g_gpu.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or no_writes" > gpurun_out/r2g_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2g_pytest.log
BLT_DENSE=0 timeout 900 python tools/kbench.py --variants 3,4 --configs 2,3 > gpurun_out/r2g_kbench.log 2>&1
echo "kbench exit $?" >> gpurun_out/r2g_kbench.log
BLT_DENSE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused --launch-skip 2 -c 1 -f -o gpurun_out/r2g_fused_r8_cfg2 \
    python tools/kbench.py --bytes 268435456 --iters 1 --variants 3 --configs 2 > gpurun_out/r2g_ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/r2g_ncu.log
tail -5 gpurun_out/r2g_pytest.log; cat gpurun_out/r2g_kbench.log
