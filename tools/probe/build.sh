#!/bin/bash
# builds tools/probe/probe (sm_100a); the binary travels to the GPU box with the snapshot (git-ignored)
cd "$(dirname "$0")"
/usr/local/cuda/bin/nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DSP_NI_FQ -o probe probe.cu
