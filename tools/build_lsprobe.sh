#!/bin/bash
# private copy of csrc/lanczos_small.cu with phase clocks (PTB_LS_PROFILE) for tools/local_step_probe.py --phases
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_probe
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -DPTB_LS_PROFILE \
     -shared pytenet_b200/csrc/lanczos_small.cu -o tools/_probe/liblsprobe.so
echo built tools/_probe/liblsprobe.so
