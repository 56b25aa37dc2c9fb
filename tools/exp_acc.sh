#!/bin/bash
# TMEM accumulator stages x issuer wait flavour (libraries built with -DLY_TC_EXP; _lib_spin also with -DLY_MBAR_SPIN_ISSUER)
cd "$GRAFT_REPO_ROOT" || exit 1
SPECS="conv:k=3,cin=64,cout=64,hw=80 conv:k=3,cin=64,cout=64,hw=80,res=1 conv:k=3,cin=32,cout=32,hw=160 conv:k=3,cin=32,cout=32,hw=160,res=1 conv:k=1,cin=64,cout=64,hw=160 conv:k=3,cin=128,cout=128,hw=40 conv:k=3,cin=128,cout=128,hw=40,res=1 conv:k=1,cin=512,cout=512,hw=20 conv:k=1,cin=128,cout=128,hw=80 conv:k=3,s=2,cin=32,cout=64,hw=320 conv:k=3,s=2,cin=64,cout=128,hw=160 conv:k=1,cin=128,cout=80,hw=80,nchw=1,act=0"
for lib in _lib _lib_spin; do
for acc in 2 4 8; do
for e in 0 7; do
  echo "=== lib=$lib ACC=$acc EXP=$e"
  LEANYOLO_B200_LIB=$GRAFT_REPO_ROOT/leanyolo_b200/$lib/libleanyolo_b200.so LY_TC_ACC=$acc LY_TC_EXP=$e python tools/bench_ops.py $SPECS 2>&1 | grep "^conv"
done; done; done
echo "=== parity (ACC=8, default lib)"
python tools/gpu_diag.py --filter conv 2>&1 | tail -4
python tools/gpu_diag.py --filter model 2>&1 | tail -4
