#!/bin/bash
# ncu --set full of ONE kernel of a bench_ops spec, summarised on the box: tools/ncu_kernel.sh <out-name> <kernel-regex> <spec> [ENV=VAL ...]
cd "$GRAFT_REPO_ROOT" || exit 1
name=$1; kre=$2; spec=$3; shift 3
for kv in "$@"; do export "$kv"; done
python tools/bench_ops.py $spec > gpurun_out/${name}_plain.log 2>&1 || { tail -5 gpurun_out/${name}_plain.log; exit 1; }
LY_BENCH_ITERS=1 ncu --set full --clock-control none --import-source on -k regex:$kre -s 1 -c 1 -o gpurun_out/$name -f python tools/bench_ops.py $spec > gpurun_out/${name}_ncu.log 2>&1
python tools/ncu_table.py gpurun_out/$name.ncu-rep > gpurun_out/${name}_table.txt 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,r=rows[0],rows[1],rows[2]
pat=['inst_executed','pipe_','issue','warp_issue_stalled','warps_active','average_warp','registers','occupancy','l1tex__data_bank','throughput','dram__bytes','lts__t_bytes','shared','cycles_elapsed.max','gpu__time']
for k,un,v in zip(h,u,r):
    if any(p in k for p in pat): print(k,'|',un,'|',v)
" > gpurun_out/${name}_metrics.txt
ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_source.csv 2>/dev/null
ls -la gpurun_out/$name.ncu-rep; rm -f gpurun_out/$name.ncu-rep
tail -2 gpurun_out/${name}_plain.log; cat gpurun_out/${name}_table.txt
