#!/bin/bash
# All six variants + BASELINE configs 3 and 5 on one GPU with the current build (3 timed steps each)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2var; mkdir -p $O
run() { # name, args...
  local name=$1; shift
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > $O/$name.json 2> $O/$name.err || echo "FAILED $name"
}
run n --model yolov10n
run s --model yolov10s
run m --model yolov10m
run b --model yolov10b
run l --model yolov10l
run x --model yolov10x
run cfg3_m_nms --model yolov10m --decode nms --conf 0.001 --iou 0.7
run cfg5_l_1280 --model yolov10l --imgsz 1280 --batch 64
python - <<'PY'
import json,glob,os
out={}
for f in sorted(glob.glob('gpurun_out/r2var/*.json')):
    try: d=json.load(open(f))
    except Exception as e: print(f, 'unreadable', e); continue
    r=d.get('roofline',{})
    out[d['config']['workload']]={'value':d['value'],'ms_per_step':d['ms_per_step'],'conv_tc_frac':r.get('frac'),'whole_step_tensor_frac':r.get('whole_step_tensor_frac'),'clocks':d.get('clocks'),'by_kind_ms':{k:v['ms'] for k,v in r.get('by_kind',{}).items()}}
    print(d['config']['workload'], d['value'], d['ms_per_step'], r.get('frac'), r.get('whole_step_tensor_frac'))
json.dump(out, open('gpurun_out/r2var/summary.json','w'), indent=1)
PY
