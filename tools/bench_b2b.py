"""Time the back-to-back 3x3 -> 1x1 tail (LY_OP_CHAIN -> conv_b2b.cu) at the head shapes: python tools/bench_b2b.py [hw ...]"""
import ctypes as C, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from leanyolo_b200 import _native as N
import gpu_checks_chain as CH

def run(hw, B=256, c=64, cout=64):
    op, keep = CH.build_tail_op(B, hw, hw, c, cout)
    lib = N.lib(); stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); N.check(lib.ly_launch(C.byref(op), stream), "b2b"); e1.record(); torch.cuda.synchronize()
        if i: ts.append(e0.elapsed_time(e1))
    print(f"b2b tail {c}->{c}->{cout} @{hw}x{hw} B{B}: {statistics.median(ts):.3f} ms", flush=True)

for hw in [int(a) for a in sys.argv[1:]] or [80, 40, 20]:
    run(hw)
