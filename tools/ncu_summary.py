"""Summarise an .ncu-rep: headline metrics + top SASS lines by stall samples.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [ntop]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[-1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct", "l1tex__throughput.avg.pct",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct", "sm__inst_executed_pipe_alu.sum.pct",
        "sm__inst_executed_pipe_lsu.sum.pct", "launch__grid_size", "launch__block_size", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled"]
for i, h in enumerate(hdr):
    if any(h.startswith(w) for w in want) and "pcsamp" not in h:
        print(f"{h} = {vals[i]} {units[i]}")
print("--- stall reasons (pc samples)")
st = [(int(float(vals[i] or 0)), h) for i, h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
tot = sum(v for v, _ in st) or 1
for v, h in sorted(st, reverse=True)[:8]:
    print(f"  {h.split('stalled_')[1]:24s} {v:8d} {100*v/tot:5.1f}%")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))[2:]
tots = sum(int(r[4]) for r in srows) or 1
print(f"--- top SASS by samples (total {tots}, instructions executed {sum(int(r[5]) for r in srows)})")
for r in sorted(srows, key=lambda r: -int(r[4]))[:ntop]:
    print(f"  {int(r[4]):6d} {100*int(r[4])/tots:5.1f}%  exec {int(r[5]):9d}  {r[1].strip()[:100]}")
