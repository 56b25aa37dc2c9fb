#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
export LEANYOLO_B200_LIB=$GRAFT_REPO_ROOT/leanyolo_b200/_lib_prof/libleanyolo_b200.so LY_BENCH_ITERS=1
for spec in conv:k=3,cin=64,cout=64,hw=80 conv:k=1,cin=64,cout=64,hw=160 conv:k=3,cin=32,cout=32,hw=160 conv:k=1,cin=128,cout=128,hw=80; do
for e in 0 1 2; do
  echo "=== $spec EXP=$e"
  LY_TC_EXP=$e python tools/bench_ops.py $spec 2>&1 | tail -6
done; done
