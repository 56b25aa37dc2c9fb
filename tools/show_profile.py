"""Print a per-op profile JSON written by bench.py --profile-out (optionally diff against another)."""
import json
import sys

a = json.load(open(sys.argv[1]))["rows"]
b = json.load(open(sys.argv[2]))["rows"] if len(sys.argv) > 2 else None
tot = {}
print("idx kind  k s  cin cout   hw      ms   TFLOP/s    GB/s" + ("   prev_ms" if b else ""))
for i, r in enumerate(a):
    tf = r["flops"] / (r["ms"] / 1e3) / 1e12 if r["ms"] > 0 else 0
    gb = r["bytes"] / (r["ms"] / 1e3) / 1e9 if r["ms"] > 0 else 0
    key = ("conv%dx%d" % (r["k"], r["k"]) if r["kind"] == "conv" else r["kind"])
    tot[key] = tot.get(key, 0) + r["ms"]
    extra = f" {b[i]['ms']:9.3f}" if b and i < len(b) else ""
    if "-q" not in sys.argv:
        print(f"{i:3d} {r['kind']:5s} {r['k']} {r['stride']} {r['cin']:4d} {r['cout']:4d} {r['hw']:4d} {r['ms']:7.3f} {tf:9.1f} {gb:7.0f}{extra}")
print({k: round(v, 3) for k, v in tot.items()}, "total", round(sum(tot.values()), 3))
