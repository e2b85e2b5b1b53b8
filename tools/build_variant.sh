#!/bin/bash
# Developer harness: builds the library from the current sources with extra nvcc flags into build/variants/<name>.so
# usage: tools/build_variant.sh <name> [nvcc flags, e.g. -DDTO_SCAN_THREADS=448]
set -e
name=$1; shift
src=dual_threshold_optimization_b200/csrc
tmp=$(mktemp -d)
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-ffp-contract=off $@"
nvcc $F -c $src/dto_kernels.cu -o $tmp/k.o &
nvcc $F -c $src/dto_engine.cu -o $tmp/e.o &
wait
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $tmp/k.o $tmp/e.o dual_threshold_optimization_b200/lib/obj/dto_host.o -cudart static -lpthread -ldl
rm -rf $tmp
