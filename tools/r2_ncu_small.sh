#!/bin/bash
# ncu --set full over the non-GEMM kernels of one headline step (bandwidth / latency kernels)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2ncu; mkdir -p $O
timeout 900 ncu --set full --clock-control none -k "regex:dw_|dw7|pool|attn|stem|topk|best_|dwpw" --launch-skip 70 -c 36 -o $O/small \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu.log 2>&1
ncu -i $O/small.ncu-rep --page details --csv > $O/small_details.csv 2>/dev/null
rm -f $O/small.ncu-rep
tail -3 $O/ncu.log
