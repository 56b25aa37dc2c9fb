#!/bin/bash
# full GPU test suite + default bench (+ per-op table)
cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --profile-out gpurun_out/check_perop.json 2>/dev/null > gpurun_out/check_bench.json
python -c "import json; d=json.load(open('gpurun_out/check_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['whole_step_tensor_frac'], {k:v['ms'] for k,v in d['roofline']['by_kind'].items()})"
