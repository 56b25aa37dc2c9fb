"""Fixed cost of one launch: tiny problems (<= 1 tile per SM) launched back to back.
    python tools/launch_floor.py
Prints the per-launch time of (a) an empty-ish kernel (upsample of a 1-pixel map), (b) the
tcgen05 conv on a one-tile-per-SM problem, (c) the TMA depthwise kernel, each from N
back-to-back launches on one stream (CUDA events around the whole train)."""
import ctypes as C, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from leanyolo_b200 import _native as N
from gpu_checks import view
DEV = "cuda"

def conv_op(B, hw, cin, cout, k=1):
    x = torch.randn(B, hw, hw, cin, device=DEV).to(torch.bfloat16)
    w = (torch.randn(cout, k, k, cin, device=DEV) / math.sqrt(cin * k * k)).to(torch.bfloat16)
    b = torch.randn(cout, device=DEV)
    y = torch.empty(B, hw, hw, cout, device=DEV, dtype=torch.bfloat16)
    op = N.LyOp(); op.dtype, op.B, op.ext_slot = N.LY_BF16, B, -1
    op.kind, op.k, op.stride, op.act = N.OP_CONV, k, 1, 1
    op.src, op.dst = view(x, 0, cin), view(y, 0, cout)
    op.w, op.bias = w.data_ptr(), b.data_ptr()
    return op, (x, w, b, y)

def dw_op(B, hw, c):
    x = torch.randn(B, hw, hw, c, device=DEV).to(torch.bfloat16)
    w = torch.randn(9, c, device=DEV).to(torch.bfloat16); b = torch.randn(c, device=DEV)
    y = torch.empty_like(x)
    op = N.LyOp(); op.dtype, op.B, op.ext_slot = N.LY_BF16, B, -1
    op.kind, op.k, op.stride, op.act = N.OP_DW, 3, 1, 1
    op.src, op.dst = view(x), view(y); op.w, op.bias = w.data_ptr(), b.data_ptr()
    return op, (x, w, b, y)

def plan_time(ops, reps=200):
    lib = N.lib()
    arr = (N.LyOp * len(ops))(*ops)
    h = C.c_void_p()
    N.check(lib.ly_plan_create(arr, len(ops), C.byref(h)), "create")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ext = (C.c_void_p * 1)()
    for _ in range(3):
        N.check(lib.ly_plan_run(h, ext, 1, 0, st), "run")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        N.check(lib.ly_plan_run(h, ext, 1, 0, st), "run")
    e1.record(); torch.cuda.synchronize()
    lib.ly_plan_destroy(h)
    return e0.elapsed_time(e1) * 1e3 / (reps * len(ops))

for name, mk in (("conv 1x1 64->64, 50 tiles", lambda: conv_op(1, 80, 64, 64)),
                 ("conv 1x1 64->64, 148 tiles", lambda: conv_op(1, 137, 64, 64)),
                 ("conv 1x1 64->64, 1480 tiles", lambda: conv_op(10, 137, 64, 64)),
                 ("conv 3x3 128->128, 1 tile/SM", lambda: conv_op(3, 80, 128, 128, 3)),
                 ("conv 1x1 512->512 20x20 B=8", lambda: conv_op(8, 20, 512, 512)),
                 ("dw 3x3 c128 80x80 B=1", lambda: dw_op(1, 80, 128))):
    made = [mk() for _ in range(8)]
    us = plan_time([m[0] for m in made])
    print(f"{name:40s} {us:7.2f} us / launch", flush=True)
