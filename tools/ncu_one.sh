#!/bin/bash
# ncu --set full (with source) of ONE bench_ops spec: tools/ncu_one.sh <out-name> <spec> [ENV=VAL ...]
cd "$GRAFT_REPO_ROOT" || exit 1
name=$1; spec=$2; shift 2
for kv in "$@"; do export "$kv"; done
export LY_BENCH_ITERS=1
python tools/bench_ops.py $spec > gpurun_out/${name}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -o gpurun_out/$name python tools/bench_ops.py $spec > gpurun_out/${name}_ncu.log 2>&1
tail -2 gpurun_out/${name}_plain.log; tail -2 gpurun_out/${name}_ncu.log
