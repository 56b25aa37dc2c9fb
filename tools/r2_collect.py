"""Copy the outputs of tools/r2_final.sh (gpurun_out/r2final/) into profiles/ with their summaries."""
import csv
import json
import os
import shutil
import statistics
import subprocess
import sys

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O = os.path.join(R, "gpurun_out", "r2final")
P = os.path.join(R, "profiles")
for src, dst in (("bench_default.json", "r2_final_bench_default.json"), ("bench_reference.json", "r2_final_bench_reference_arm.json"),
                 ("per_op.json", "r2_final_bench_per_op_cuda_events.json"), ("launches.csv", "r2_final_ncu_launch_list_bench_step.csv")):
    shutil.copy(os.path.join(O, src), os.path.join(P, dst))
out = subprocess.run([sys.executable, os.path.join(R, "tools", "ncu_launch_summary.py"), os.path.join(O, "launches.csv"),
                      os.path.join(P, "r2_traffic.json")], capture_output=True, text=True).stdout
open(os.path.join(P, "r2_final_ncu_launch_summary.txt"), "w").write(out)
hdr = ("# ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 183 -c 24  (python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs)\n"
       "# the first 24 conv_tc launches of one step of yolov10s 640x640 batch 256 (backbone cv1 .. c6), round-2 final build: tile-parallel epilogue, 8 TMEM stages,\n"
       "# TMA-store epilogue on resident-weight layers.  The .ncu-rep (358 MB for a whole step) is summarised on the GPU box and not kept.\n")
open(os.path.join(P, "r2_final_ncu_conv_tc_full_24_launches.txt"), "w").write(hdr + open(os.path.join(O, "ncu_conv_tc_table.txt")).read())
rows = list(csv.reader(open(os.path.join(O, "ncu_conv_tc_datapipe.csv"))))
idx = {k: i for i, k in enumerate(rows[0])}
lines = ["# l1tex data stage of the same 24 launches (ncu raw page).  tc wavefronts = the MMAs' shared-memory operand reads (32 for A + N/4 for B per\n"
         "# M = 128, K = 16 instruction; the counter peaks at one per cycle and SM); lsu shared = LDS / STS / shuffles; lsu total adds global loads and stores.\n"
         "# tc% / lsu% = wavefronts / (elapsed cycles x 148 SMs): the two counters overlap (their sum exceeds 100 % on the shortcut layers).\n",
         "%3s %9s %9s %12s %12s %12s %7s %7s %8s\n" % ("#", "time_us", "cycles", "tc_wavefr", "lsu_shared", "lsu_total", "tc%", "lsu%", "tensor%")]
for n, r in enumerate(rows[2:]):
    cyc = float(r[idx["sm__cycles_elapsed.avg"]])
    tc = float(r[idx["l1tex__data_pipe_tc_wavefronts_mem_shared.sum"]])
    ls = float(r[idx["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]])
    lsu_pct = float(r[idx["l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"]])
    lines.append("%3d %9.1f %9.0f %12.0f %12.0f %12.0f %7.1f %7.1f %8.1f\n" % (
        n, float(r[idx["gpu__time_duration.sum"]]), cyc, tc, ls, lsu_pct / 100 * cyc * 148, tc / (cyc * 148) * 100, lsu_pct,
        float(r[idx["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]])))
open(os.path.join(P, "r2_final_ncu_conv_tc_datapipe.txt"), "w").writelines(lines)
hdr2 = ("# ncu --set full --clock-control none -k 'regex:^(conv_b2b|dwpw_mma|stem_mma|dw7|attn_mma|best|topk)_kernel' -s 72 -c 18 "
        "(python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs), round-2 final build\n")
open(os.path.join(P, "r2_final_ncu_other_kernels.txt"), "w").write(hdr2 + open(os.path.join(O, "ncu_other_table.txt")).read())
rows = list(csv.reader(open(os.path.join(O, "clocks.csv"))))[1:]
sm = [int(r[1].split()[0]) for r in rows]
pw = [float(r[3].split()[0]) for r in rows]
load = [s for s, w in zip(sm, pw) if w > 300]
reasons = sorted(set(r[4].strip() for r in rows))
quiet = all(r[5].strip() == "Not Active" and r[6].strip() == "Not Active" and r[7].strip() == "Not Active" for r in rows)
open(os.path.join(P, "r2_final_clocks.txt"), "w").write(
    f"nvidia-smi -lms 200 during `python bench.py --steps 20 --warmup 5` (round-2 final build): {len(sm)} samples, {len(load)} above 300 W: "
    f"median SM clock under load {statistics.median(load) if load else 'n/a'} MHz (max {max(sm)}), power max {max(pw):.0f} W, "
    f"clocks_event_reasons.active values seen: {reasons} (0x1 = gpu_idle, 0x4 = sw_power_cap); hw_slowdown / hw_thermal / sw_thermal never active: {quiet}\n")
d = json.load(open(os.path.join(O, "bench_default.json")))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["whole_step_tensor_frac"], d["cpu_baseline"]["value"])
print(out)
print(open(os.path.join(P, "r2_final_clocks.txt")).read())
