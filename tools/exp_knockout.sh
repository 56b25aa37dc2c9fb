#!/bin/bash
# Knock-out experiment (library built with LY_NVCC_EXTRA=-DLY_TC_EXP): which role bounds a conv_tc layer?
cd "$GRAFT_REPO_ROOT" || exit 1
SPECS="conv:k=3,cin=64,cout=64,hw=80 conv:k=3,cin=64,cout=64,hw=80,res=1 conv:k=3,cin=32,cout=32,hw=160 conv:k=3,cin=32,cout=32,hw=160,res=1 conv:k=1,cin=64,cout=64,hw=160 conv:k=3,cin=128,cout=128,hw=40 conv:k=3,cin=128,cout=128,hw=40,res=1 conv:k=1,cin=512,cout=512,hw=20 conv:k=1,cin=128,cout=256,hw=80"
for e in 0 1 2 3 4 5 6 7; do
  echo "=== LY_TC_EXP=$e"
  LY_TC_EXP=$e LY_TC_DEBUG=$([ $e = 0 ] && echo 1 || echo 0) python tools/bench_ops.py $SPECS 2>&1 | grep -v "^$"
done
