#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "concurrent or fused or no_writes or predictor" 2>&1 | tail -5
BLT_DENSE=0 timeout 300 python tools/kbench.py --variants 3 --configs 2,9 2>&1 | cut -c1-300
