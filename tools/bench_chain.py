"""Time LY_OP_CHAIN alone (CUDA events) on the shapes yolov10s uses: the 160x160 C2f block and the regression tails."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from leanyolo_b200 import _native as N  # noqa: E402
import gpu_checks_chain as CH  # noqa: E402
from gpu_checks import DEV, view  # noqa: E402


def run(kind, B, H, W, iters=5, **kw):
    g = torch.Generator().manual_seed(0)
    regions, n_in, stages = CH.SPECS[kind](g, **kw)
    cin, cout = sum(regions[:n_in]), stages[-1]["cout"]
    keep = []
    ch = CH.make_chain(regions, n_in, stages, keep)
    x = torch.randn(B, H, W, cin, device=DEV).to(torch.bfloat16)
    d = torch.empty(B, H, W, cout, device=DEV, dtype=torch.bfloat16)
    op = N.LyOp()
    op.kind, op.dtype, op.B, op.k, op.stride, op.act, op.ext_slot = N.OP_CHAIN, N.LY_BF16, B, 1, 1, 0, -1
    op.src, op.dst = view(x, 0, cin), view(d, 0, cout)
    op.chain = C.pointer(ch)
    st = torch.cuda.current_stream().cuda_stream
    lib = N.lib()
    N.check(lib.ly_launch(C.byref(op), C.c_void_p(st)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        N.check(lib.ly_launch(C.byref(op), C.c_void_p(st)))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    fl = sum(2 * B * H * W * s["cout"] * s["w"].shape[3] * s["k"] ** 2 for s in stages)
    print(f"RESULT {kind} {kw} B{B} {H}x{W}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    B = int(os.environ.get("B", "256"))
    run("c2f", B, 160, 160, c=32)
    run("tail", B, 80, 80, c=64, cout=64)
    run("single", B, 80, 80, k=3, cin=64, cout=64)
    run("single", B, 160, 160, k=3, cin=32, cout=32)
    run("single", B, 160, 160, k=1, cin=64, cout=64)
