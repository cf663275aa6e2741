#!/bin/bash
# usage: build_variant.sh <name> [nvcc -D flags...]: builds the working tree's csrc with extra flags into ab/liblist_<name>.so
set -e
name=$1; shift
src=learning-implicitly-from-spatial-transformers-network_b200/csrc
mkdir -p ab /tmp/ab_obj_$name
objs=""
for f in $src/*.cu; do
  o=/tmp/ab_obj_$name/$(basename $f .cu).o
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr "$@" -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ab/liblist_$name.so $objs
ls -la ab/liblist_$name.so
