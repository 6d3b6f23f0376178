#!/bin/bash
# Build a tuning variant of the CUDA library next to the in-tree one: tools/build_variant.sh <name> [git-rev] [nvcc flags...]
# -> variants/<name>.so (git-ignored; travels to the GPU box). With a git revision the csrc/ and include/ of that
# revision are used (A/B against an older kernel), otherwise the working tree.
set -e
name=$1; shift
rev=""; if [ -n "$1" ] && [[ "$1" != -* ]]; then rev=$1; shift; fi
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/variants
src=$root
if [ -n "$rev" ]; then
  src=$(mktemp -d)
  (cd $root && git archive $rev f16_jsb_b200/csrc include) | tar -x -C $src
fi
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o $root/variants/$name.so $src/f16_jsb_b200/csrc/*.cu
echo built variants/$name.so
