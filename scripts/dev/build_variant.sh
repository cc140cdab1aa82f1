#!/bin/bash
# scripts/dev/build_variant.sh <name> [-DEMPC_...=..]...  ->  build_var/libempc_<name>.so (same flags as the Makefile + the defines)
name=$1; shift
mkdir -p build_var
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c -o build_var/$name.o eagle-mpc_b200/csrc/solver.cu && \
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_var/libempc_$name.so build_var/$name.o && rm -f build_var/$name.o && echo built $name
