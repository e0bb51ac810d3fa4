#!/bin/bash
# round-2 validation: the whole GPU suite, both bench arms, kernel micro-benchmarks
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${1:-r2full}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench exit $?" >> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
echo "bench ref exit $?" >> gpurun_out/${T}_bench_ref.err
tail -5 gpurun_out/${T}_pytest.log; tail -3 gpurun_out/${T}_bench.err; cut -c1-3000 gpurun_out/${T}_bench.json; cut -c1-600 gpurun_out/${T}_bench_ref.json
