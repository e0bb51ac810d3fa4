#!/bin/bash
# detokenizer iteration: parity of both forms, timings of both forms, one ncu capture of the fused form on config 2's tokens
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${1:-r4d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "detok or round_trip or fuzz or no_writes" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 600 python tools/kbench.py --configs 7 > gpurun_out/${T}_kbench.log 2>&1
BLT_DETOK_VARIANT=1 timeout 600 python tools/kbench.py --configs 7 >> gpurun_out/${T}_kbench.log 2>&1
if [ "$2" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:detok_emit --launch-skip 5 -c 1 -f -o gpurun_out/${T}_detok_emit \
    python tools/kbench.py --bytes 268435456 --iters 1 --configs 7 > gpurun_out/${T}_ncu.log 2>&1
fi
tail -4 gpurun_out/${T}_pytest.log; cut -c1-300 gpurun_out/${T}_kbench.log
