#!/bin/bash
# First GPU contact: smoke, parity tests, kernel micro-bench, bench line.  Everything is logged under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/gpu.txt
free -g | head -2 >> gpurun_out/gpu.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
echo "== kbench" ; timeout 900 python tools/kbench.py > gpurun_out/kbench.log 2> gpurun_out/kbench.err ; echo "kbench rc=$?"
cat gpurun_out/kbench.log ; tail -3 gpurun_out/kbench.err
echo "== bench" ; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err ; echo "bench rc=$?"
cat gpurun_out/bench.log ; tail -5 gpurun_out/bench.err
