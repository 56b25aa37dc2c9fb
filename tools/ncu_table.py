"""One line per profiled launch of an .ncu-rep (ncu --set full): duration, DRAM bytes, tensor-pipe %, issue %.
    python tools/ncu_table.py gpurun_out/prof.ncu-rep"""
import csv
import io
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%")]
print(" | ".join(f"{n} [{units[col[k]]}]" if k in col and units[col[k]] else n for k, n in want))
for r in rows[2:]:
    out = []
    for k, n in want:
        v = r[col[k]] if k in col else "-"
        if k == "Kernel Name":
            v = v.split("(")[0].split("::")[-1]
        else:
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
        out.append(v)
    print(" | ".join(out))
