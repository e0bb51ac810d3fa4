#!/bin/bash
# round-2 kernel iteration: parity of the fused sweep, kernel timings, phase probe, one ncu capture
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${1:-r2x}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or no_writes" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
BLT_DENSE=0 timeout 900 python tools/kbench.py --variants 3,4 --configs 2,3 > gpurun_out/${T}_kbench.log 2>&1
echo "kbench exit $?" >> gpurun_out/${T}_kbench.log
if [ -f blt_b200/lib_prof/libblt_cuda_prof.so ]; then
  BLT_PROF_LIB=$PWD/blt_b200/lib_prof/libblt_cuda_prof.so timeout 300 python tools/fused_phase_probe.py > gpurun_out/${T}_probe.log 2>&1; BLT_SWEEP_VARIANT=4 BLT_PROF_LIB=$PWD/blt_b200/lib_prof/libblt_cuda_prof.so timeout 300 python tools/fused_phase_probe.py >> gpurun_out/${T}_probe.log 2>&1
fi
if [ "$2" = "ncu" ]; then
BLT_DENSE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused --launch-skip 2 -c 1 -f -o gpurun_out/${T}_fused_cfg2 \
    python tools/kbench.py --bytes 268435456 --iters 1 --variants 3 --configs 2 > gpurun_out/${T}_ncu.log 2>&1
fi
tail -3 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_kbench.log; cat gpurun_out/${T}_probe.log 2>/dev/null
