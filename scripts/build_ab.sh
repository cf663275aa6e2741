#!/bin/bash
# usage: build_ab.sh <git-ref> <name>: builds that revision's csrc into ab/liblist_<name>.so (same-box A/B with scripts/gpu_ab.sh)
set -e
ref=$1; name=$2
rm -rf /tmp/wt_ab && git worktree add -q /tmp/wt_ab $ref
src=/tmp/wt_ab/learning-implicitly-from-spatial-transformers-network_b200/csrc
mkdir -p ab /tmp/ab_obj
objs=""
for f in $src/*.cu; do
  o=/tmp/ab_obj/$(basename $f .cu).o
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ab/liblist_$name.so $objs
git worktree remove --force /tmp/wt_ab
ls -la ab/liblist_$name.so
