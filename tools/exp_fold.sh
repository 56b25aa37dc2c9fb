#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
SPECS="conv:k=3,cin=64,cout=64,hw=80 conv:k=3,cin=64,cout=64,hw=80,res=1 conv:k=3,cin=32,cout=32,hw=160 conv:k=3,cin=32,cout=32,hw=160,res=1 conv:k=3,cin=64,cout=64,hw=40 conv:k=3,cin=64,cout=64,hw=20"
for f in 0 2; do
  echo "=== FOLD=$f"
  LY_TC_FOLD=$f python tools/bench_ops.py $SPECS 2>&1 | grep "^conv"
done
echo "=== parity FOLD=2"
LY_TC_FOLD=2 python tools/gpu_diag.py --filter conv 2>&1 | tail -4
LY_TC_FOLD=2 python tools/gpu_diag.py --filter model 2>&1 | tail -4
for f in 0 2; do
  echo "=== bench FOLD=$f"
  LY_TC_FOLD=$f python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/fold${f}_perop.json 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], {k:v['ms'] for k,v in d['roofline']['by_kind'].items()})"
done
