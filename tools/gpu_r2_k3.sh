#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "general_maps or fuzz or derived or ref_bpe" 2>&1 | tail -3
python tools/kbench.py --configs 6 2>&1 | cut -c1-400
ncu --set full --clock-control none -k regex:count_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r2_k3_count python tools/kbench.py --configs 6 > gpurun_out/r2_k3_ncu.log 2>&1
python tools/ncu_keys.py gpurun_out/r2_k3_count.ncu-rep | grep -E "==|lts__t_sector_hit|duration|dram__bytes"
