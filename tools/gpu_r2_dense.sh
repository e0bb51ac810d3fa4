#!/bin/bash
# dense-pass iteration: parity of the partial fall-back, the rare-pair probe, kernel timings of the dense path
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${1:-r4p}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -k "dense or random_vs_oracle or no_writes or fuzz or config3 or config4" > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 600 python tools/rare_pair_probe.py > gpurun_out/${T}_rare_pair.json 2> gpurun_out/${T}_rare_pair.err
timeout 600 python tools/kbench.py --configs 3,4 --variants 3 > gpurun_out/${T}_kbench.log 2>&1
tail -4 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_rare_pair.json; cut -c1-300 gpurun_out/${T}_kbench.log
