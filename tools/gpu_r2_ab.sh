#!/bin/bash
# A/B of two builds of the library on one box: parity of the fused sweep with the default build, then kernel timings of both.
#   gpu_r2_ab.sh TAG ALT_LIB     ALT_LIB: a second build of libblt_cuda.so, e.g. made with the nvcc line of
#   tools/build_prof_lib.sh and -DBLT_FZ_NO_BULK_FLUSH / -DBLT_FZ_UNPRED_STS into blt_b200/lib_prof/ (kbench loads it through BLT_ALT_LIB)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${1:-r4ab}; ALT=${2:?path of the alternative library}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "fused or no_writes or fuzz or random_vs_oracle or config3 or sparse_table" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
for rep in 1 2; do
BLT_DENSE=0 timeout 600 python tools/kbench.py --variants 3 --configs 2,3,9 >> gpurun_out/${T}_kbench_default.log 2>&1
BLT_ALT_LIB=$PWD/$ALT BLT_DENSE=0 timeout 600 python tools/kbench.py --variants 3 --configs 2,3,9 >> gpurun_out/${T}_kbench_alt.log 2>&1
done
tail -3 gpurun_out/${T}_pytest.log; echo default; cut -c1-260 gpurun_out/${T}_kbench_default.log; echo alt; cut -c1-260 gpurun_out/${T}_kbench_alt.log
