#!/bin/bash
# round-2 GPU call 1: baseline bench, sanitizer passes over small parity cases, NMS ncu capture (config 3)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2c1; mkdir -p $O
python bench.py --steps 10 --warmup 3 --profile-out $O/per_op.json > $O/bench_base.json 2> $O/bench_base.err
python bench.py --model yolov10m --decode nms --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_cfg3.json 2> $O/bench_cfg3.err
CASES=tc_1x1_64_64,tc_3x3_64_64,tc_3x3_32_32,tc_3x3_s2,tc_3x3_res,tc_1x1_up_128,tc_3x3_128_128_nonres,tc_1x1_nchw_80,dwpw_128_128_20,dwpw_256_128_40,dwtma_3_c128_20_res,dwtma_3s2_c64_40,dwtma_7_c64_40,pool_bf16,attn_32_64_bf16,stem_bf16,topk_golden,nms_exact_3000,nms_exact_classwise,model_yolov10n_bf16
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --target-processes all --print-limit 20 python tools/gpu_diag.py --worker $CASES > $O/san_$tool.log 2>&1
  echo "rc=$?" >> $O/san_$tool.log
done
timeout 600 ncu --set full --clock-control none -k regex:nms -c 3 -o $O/ncu_nms python bench.py --model yolov10m --decode nms --steps 1 --warmup 1 --no-cpu-baseline > $O/ncu_nms.log 2>&1
ncu -i $O/ncu_nms.ncu-rep --page details --csv > $O/ncu_nms_details.csv 2>/dev/null
rm -f $O/ncu_nms.ncu-rep
tail -3 $O/san_*.log; cat $O/bench_base.json | head -c 1500
