"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list.
    python tools/ncu_launch_summary.py gpurun_out/launches.csv [traffic.json]"""
import collections
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
L = collections.OrderedDict()
for r in rows:
    d = L.setdefault(int(r[0]), {"k": r[4].split("(")[0].split("::")[-1]})
    d[r[12]] = float(r[14].replace(",", ""))
agg = collections.OrderedDict()
for d in L.values():
    a = agg.setdefault(d["k"], [0, 0.0, 0.0])
    a[0] += 1
    a[1] += d["gpu__time_duration.sum"] / 1e6
    a[2] += (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / 1e9
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':40s} {'launches':>8s} {'time ms':>9s} {'share':>7s} {'DRAM GB':>9s} {'DRAM GB/launch':>15s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:40]:40s} {a[0]:8d} {a[1]:9.3f} {a[1] / tot:7.3f} {a[2]:9.3f} {a[2] / a[0]:15.4f}")
print(f"{'total':40s} {sum(a[0] for a in agg.values()):8d} {tot:9.3f}")
if len(sys.argv) > 2:
    out = {k: {"launches_per_step": a[0], "dram_bytes_per_launch": int(a[2] / a[0] * 1e9), "dram_bytes_per_step": int(a[2] * 1e9),
               "time_ms_per_step_under_ncu": round(a[1], 3)} for k, a in agg.items()}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
