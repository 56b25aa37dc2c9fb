"""Time the back-to-back 3x3 -> 1x1 tail (LY_OP_CHAIN -> conv_b2b.cu) at the head shapes: python tools/bench_b2b.py [hw ...]"""
import ctypes as C, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from leanyolo_b200 import _native as N
import gpu_checks_chain as CH

def run(hw, B=256, c=64, cout=64, cmid=None, stride0=1):
    op, keep = CH.build_tail_op(B, hw, hw, c, cout, cmid=cmid, stride0=stride0)
    lib = N.lib(); stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); N.check(lib.ly_launch(C.byref(op), stream), "b2b"); e1.record(); torch.cuda.synchronize()
        if i: ts.append(e0.elapsed_time(e1))
    print(f"b2b s{stride0} {c}->{cmid or c}->{cout} @{hw}x{hw} B{B}: {statistics.median(ts):.3f} ms", flush=True)

for a in sys.argv[1:] or ["80", "40", "20"]:
    if a.startswith("s2:"):          # s2:320 = backbone cv1 -> c2.cv1 of yolov10s on a 320x320 input
        run(int(a[3:]), c=32, cmid=64, cout=64, stride0=2)
    else:
        run(int(a))
