#!/bin/bash
# round-2 evidence: kernel micro-benchmarks of every path, ncu captures (launch list of bench.py, full set of the fused sweep)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${1:-r2final}
mkdir -p gpurun_out
python tools/kbench.py --configs 1,2,3,4,9 --variants 3 > gpurun_out/${T}_kbench_default.jsonl 2>&1
BLT_DENSE=0 python tools/kbench.py --configs 2,3,4,9 --variants 0,3,4 > gpurun_out/${T}_kbench_exact.jsonl 2>&1
python tools/kbench.py --configs 6,7,8 > gpurun_out/${T}_kbench_other.jsonl 2>&1
python bench.py --steps 5 --warmup 3 --no-config5 > gpurun_out/${T}_bench_short.json 2> gpurun_out/${T}_bench_short.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 5 --warmup 3 --no-config5 --no-cpu > gpurun_out/${T}_ncu_bench.log 2>&1
BLT_DENSE=0 ncu --set full --clock-control none --import-source on -k regex:fused --launch-skip 2 -c 1 -f -o gpurun_out/${T}_fused_mixed python tools/kbench.py --bytes 268435456 --iters 1 --variants 3 --configs 9 > gpurun_out/${T}_ncu_fused.log 2>&1
ncu --set full --clock-control none -k regex:pair_hist --launch-skip 1 -c 1 -f -o gpurun_out/${T}_pairhist python tools/kbench.py --bytes 268435456 --iters 1 --configs 8 > gpurun_out/${T}_ncu_pairhist.log 2>&1
for f in kbench_default kbench_exact kbench_other; do echo "== $f"; cut -c1-330 gpurun_out/${T}_$f.jsonl; done
