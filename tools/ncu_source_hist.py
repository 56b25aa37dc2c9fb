"""Opcode histogram and hottest SASS lines of an `ncu --page source --csv` dump (tools/ncu_kernel.sh writes one).
    python tools/ncu_source_hist.py gpurun_out/<name>_source.csv [warps]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
c = {k: i for i, k in enumerate(h)}
ops, smp = collections.Counter(), collections.Counter()
lines = []
for n, r in enumerate(rows[hi + 1:]):
    if len(r) < len(h):
        continue
    toks = r[c["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    cnt, s = int(r[c["Instructions Executed"]]), int(r[c["# Samples"]])
    ops[op] += cnt
    smp[op] += s
    lines.append((s, cnt, n, r[c["Source"]].strip()))
tot, tots = sum(ops.values()), sum(smp.values())
print(f"instructions executed (warp level): {tot}   stall samples: {tots}")
for k, v in ops.most_common(22):
    print(f"  {k:10s} {v:12d} {100 * v / tot:5.1f} %   samples {100 * smp[k] / max(tots, 1):5.1f} %")
print("hottest lines (samples, executed, index, SASS):")
for s, cnt, n, src in sorted(lines, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"  {s:6d} {cnt:10d} {n:5d}  {src[:100]}")
