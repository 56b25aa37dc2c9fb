set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/plain_bench_r1f.log 2> gpurun_out/plain_bench_r1f.err || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^(attn_mma|conv_tc|dfl|dw_strip|dw_tma|dw7|dwpw|dwpw_mma|pool|stem_mma|topk|up)_kernel' -s 300 -c 100 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_bench_r1f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 213 -c 16 -o gpurun_out/prof_conv_tc_r1f -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_full_r1f.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dwpw_mma -s 24 -c 1 -o gpurun_out/prof_dwpw_mma_r1f -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_full2_r1f.log 2>&1
ls -la gpurun_out/
