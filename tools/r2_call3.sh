#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
python tools/gpu_diag.py --filter topk 2>&1 | tail -6
python tools/gpu_diag.py --filter decode_e2e 2>&1 | tail -3
python tools/gpu_diag.py --filter smoke 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2_b3.err > gpurun_out/r2_b3.json; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_b3.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['by_kind']['decode'], d.get('fused_detect'))
PY
