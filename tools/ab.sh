#!/bin/bash
# Interleaved A/B of library builds on one GPU box: tools/ab.sh "<ab_step.py args>" lib1.so lib2.so ... ("main" = in-tree)
# Three rounds, every library once per round, so that clock / thermal drift hits all of them alike.
args=$1; shift
for round in 1 2 3; do
  for lib in "$@"; do
    if [ "$lib" = main ]; then unset F16_B200_LIB; else export F16_B200_LIB=$PWD/$lib; fi
    python tools/ab_step.py $args
  done
done
