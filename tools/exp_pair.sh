#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
SPECS="conv:k=3,cin=128,cout=128,hw=40 conv:k=3,cin=128,cout=128,hw=40,res=1 conv:k=3,cin=128,cout=128,hw=80 conv:k=3,cin=256,cout=128,hw=40 conv:k=3,cin=512,cout=128,hw=20 conv:k=3,s=2,cin=128,cout=128,hw=80"
for e in 0 1; do
  echo "=== EPT_PAIR=$e"
  LY_TC_EPT_PAIR=$e python tools/bench_ops.py $SPECS 2>&1 | grep "^conv"
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/pair_perop.json 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], {k:v['ms'] for k,v in d['roofline']['by_kind'].items()})"
python tools/gpu_diag.py --filter nms 2>&1 | grep -i "flips\|PASS\|FAIL" | head -20
