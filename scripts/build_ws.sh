#!/bin/bash
# rebuild only the warp-specialised translation unit (extra nvcc flags as arguments) and relink the library
set -e
cd /root/repo/sde4mbrl_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC "$@" -c inst_hexa_ws.cu -o _build/inst_hexa_ws.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libsde4b200.so _build/*.o -lcudart
