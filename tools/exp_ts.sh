#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
SPECS="conv:k=1,cin=64,cout=64,hw=160 conv:k=1,cin=128,cout=128,hw=80 conv:k=1,cin=256,cout=128,hw=80 conv:k=3,cin=128,cout=128,hw=40 conv:k=3,cin=128,cout=128,hw=40,res=1 conv:k=3,cin=128,cout=128,hw=80 conv:k=3,s=2,cin=32,cout=64,hw=320 conv:k=3,s=2,cin=64,cout=128,hw=160 conv:k=1,cin=512,cout=128,hw=20 conv:k=1,cin=256,cout=128,hw=40 conv:k=3,cin=256,cout=128,hw=40"
echo "=== parity"
python tools/gpu_diag.py --filter tc_ 2>&1 | tail -3
python tools/gpu_diag.py --filter model 2>&1 | tail -3
for t in 0 1 2; do
  echo "=== TMASTORE=$t"
  LY_TC_TMASTORE=$t python tools/bench_ops.py $SPECS 2>&1 | grep "^conv"
done
for t in 0 2; do
  echo "=== bench TMASTORE=$t"
  LY_TC_TMASTORE=$t python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/ts${t}_perop.json 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], {k:v['ms'] for k,v in d['roofline']['by_kind'].items()})"
done
