#!/bin/bash
# host-side paths: file to file (stage log), per-chunk API with and without the pieced path, tmpfs page production
cd "${GRAFT_REPO_ROOT:-/root/repo}"
T=${1:-r2host}
mkdir -p gpurun_out
g++ -O2 -pthread -o /tmp/tmpfs_probe tools/tmpfs_probe.cpp && /tmp/tmpfs_probe > gpurun_out/${T}_tmpfs_probe.log 2>&1
python tools/file_bench.py --bytes 2147483648 --gpus 1 --merges 32768 > gpurun_out/${T}_file_bench.log 2>&1
BLT_POPULATE=1 python tools/file_bench.py --bytes 2147483648 --gpus 1 --merges 32768 > gpurun_out/${T}_file_bench_populate.log 2>&1
python tools/chunk_api_bench.py --threads 1,2,8,16 > gpurun_out/${T}_chunk_api.log 2>&1
BLT_NO_PIECED=1 python tools/chunk_api_bench.py --threads 1,2,8,16 > gpurun_out/${T}_chunk_api_nopieced.log 2>&1
for f in tmpfs_probe file_bench file_bench_populate chunk_api chunk_api_nopieced; do echo "== $f"; cut -c1-400 gpurun_out/${T}_$f.log | tail -8; done
