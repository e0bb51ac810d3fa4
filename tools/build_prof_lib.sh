#!/bin/bash
# builds blt_b200/lib_prof/libblt_cuda_prof.so: the library with the fused sweep's cycle counters (-DBLT_FUSED_PROF), for tools/fused_phase_probe.py
cd "$(dirname "$0")/.." && mkdir -p blt_b200/lib_prof
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -rdc=true -shared -Xcompiler -fPIC,-fvisibility=hidden,-pthread -DBLT_FUSED_PROF \
  -o blt_b200/lib_prof/libblt_cuda_prof.so blt_b200/csrc/kernels.cu blt_b200/csrc/cabi.cu blt_b200/csrc/pipeline.cu blt_b200/csrc/host_config.cpp -lcudadevrt
