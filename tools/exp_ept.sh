#!/bin/bash
# tile-parallel epilogue on/off: per-op times, parity, whole step
cd "$GRAFT_REPO_ROOT" || exit 1
SPECS="conv:k=3,cin=64,cout=64,hw=80 conv:k=3,cin=64,cout=64,hw=80,res=1 conv:k=3,cin=32,cout=32,hw=160 conv:k=3,cin=32,cout=32,hw=160,res=1 conv:k=1,cin=64,cout=64,hw=160 conv:k=1,cin=96,cout=64,hw=160 conv:k=1,cin=128,cout=128,hw=80 conv:k=1,cin=128,cout=128,hw=80,up=1 conv:k=1,cin=256,cout=128,hw=80 conv:k=3,s=2,cin=32,cout=64,hw=320 conv:k=3,s=2,cin=64,cout=128,hw=160 conv:k=1,cin=128,cout=80,hw=80,nchw=1,act=0 conv:k=1,cin=64,cout=64,hw=80,nchw=1,act=0 conv:k=1,cin=512,cout=256,hw=20 conv:k=3,cin=64,cout=64,hw=40"
for ept in 0 1; do
  echo "=== EPT=$ept"
  LY_TC_EPT=$ept python tools/bench_ops.py $SPECS 2>&1 | grep "^conv"
done
echo "=== parity"
python tools/gpu_diag.py --filter conv 2>&1 | tail -3
python tools/gpu_diag.py --filter model 2>&1 | tail -3
for ept in 0 1; do
  echo "=== bench EPT=$ept"
  LY_TC_EPT=$ept python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/ept${ept}_perop.json 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], {k:v['ms'] for k,v in d['roofline']['by_kind'].items()})"
done
