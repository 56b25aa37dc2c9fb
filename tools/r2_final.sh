#!/bin/bash
# round-2 final evidence: default bench (with cpu baseline + per-op table), reference arm, ncu launch list of one bench step,
# ncu --set full of conv_tc launches of one step (+ the other hot kernels).  The .ncu-rep files are summarised ON THE BOX and
# deleted (gpurun_out/ only comes back below 64 MiB).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r2final; mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 5 --profile-out $O/per_op.json > $O/bench_default.json 2> $O/bench_default.err
kill $SMI
python bench.py --impl reference --steps 8 --warmup 2 > $O/bench_reference.json 2> $O/bench_reference.err
K='regex:^(attn_mma|conv_tc|conv_b2b|best|dw_strip|dw_tma|dw7|dwpw|dwpw_mma|pool|stem_mma|topk)_kernel'
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -s 282 -c 94 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $O/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 183 -c ${NCONV:-24} -o $O/prof_conv_tc -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $O/ncu_full_conv.log 2>&1
python tools/ncu_table.py $O/prof_conv_tc.ncu-rep > $O/ncu_conv_tc_table.txt 2>&1
ncu -i $O/prof_conv_tc.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
keep=[i for i,k in enumerate(h) if any(x in k for x in ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_pipe_tc_wavefronts_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.avg','smsp__inst_executed.sum','launch__registers_per_thread','sm__throughput.avg.pct','gpu__dram_throughput.avg.pct'])]
w=csv.writer(sys.stdout)
for r in rows: w.writerow([r[i] for i in keep])
" > $O/ncu_conv_tc_datapipe.csv
rm -f $O/prof_conv_tc.ncu-rep
ncu --set full --clock-control none -k 'regex:^(conv_b2b|dwpw_mma|stem_mma|dw7|attn_mma|best|topk)_kernel' -s 72 -c 18 -o $O/prof_other -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > $O/ncu_full_other.log 2>&1
python tools/ncu_table.py $O/prof_other.ncu-rep > $O/ncu_other_table.txt 2>&1
rm -f $O/prof_other.ncu-rep
du -sh gpurun_out; ls -la $O; head -c 400 $O/bench_default.json
