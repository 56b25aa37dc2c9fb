"""Micro-benchmark single ops of the plan at the headline shapes (CUDA events, L2-cold).

    python tools/bench_ops.py conv:k=3,cin=64,cout=64,hw=80 conv:k=1,cin=64,cout=64,hw=160 dw:k=3,c=128,hw=80 ...

Each spec runs with B images (default 256) on fresh buffers several times and prints the
median ms, TFLOP/s and GB/s (algorithmic bytes: input + output (+ residual) once).
"""
from __future__ import annotations

import ctypes as C
import math
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from leanyolo_b200 import _native as N  # noqa: E402
from gpu_checks import view  # noqa: E402

DEV = "cuda"


def parse(spec):
    kind, _, rest = spec.partition(":")
    kw = {}
    for item in filter(None, rest.split(",")):
        k, v = item.split("=")
        kw[k] = int(v)
    return kind, kw


def run(spec, iters=int(os.environ.get("LY_BENCH_ITERS", "5"))):
    kind, kw = parse(spec)
    B = kw.get("B", 256)
    hw = kw.get("hw", 80)
    op = N.LyOp()
    op.dtype, op.B, op.ext_slot = N.LY_BF16, B, -1
    keep = []
    flops = byts = 0
    if kind == "conv":
        k, s, cin, cout = kw.get("k", 1), kw.get("s", 1), kw["cin"], kw["cout"]
        ctot_in, ctot_out = kw.get("ctot_in", cin), kw.get("ctot_out", cout)
        x = torch.randn(B, hw, hw, ctot_in, device=DEV).to(torch.bfloat16)
        w = (torch.randn(cout, k, k, cin, device=DEV) / math.sqrt(cin * k * k)).to(torch.bfloat16)
        b = torch.randn(cout, device=DEV)
        y = torch.empty(B, hw // s, hw // s, ctot_out, device=DEV, dtype=torch.bfloat16)
        op.kind, op.k, op.stride, op.act = N.OP_CONV, k, s, kw.get("act", 1)
        op.impl = N.IMPL_SIMT if kw.get("simt") else N.IMPL_AUTO
        op.src, op.dst = view(x, 0, cin), view(y, 0, cout)
        op.w, op.bias = w.data_ptr(), b.data_ptr()
        if kw.get("nchw"):     # public NCHW fp32 output (head finals) instead of the NHWC buffer
            o = torch.empty(B, cout, hw // s, hw // s, device=DEV)
            op.dst = N.LyView(None, 0, 0, 0, 0, 0)
            op.nchw, op.nchw_ctot, op.nchw_c0, op.nchw_c = o.data_ptr(), cout, 0, cout
            keep.append(o)
        if kw.get("up"):      # half-resolution pre-activation addend (folded upsample + concat)
            u = torch.randn(B, hw // s // 2, hw // s // 2, cout, device=DEV).to(torch.bfloat16)
            op.up = view(u, 0, cout)
            keep.append(u)
        if kw.get("res"):
            r = torch.randn_like(y)
            op.res = view(r, 0, cout)
            keep.append(r)
        keep += [x, w, b, y]
        M = B * (hw // s) ** 2
        flops = 2 * M * cout * cin * k * k
        byts = B * hw * hw * cin * 2 + M * cout * (4 if kw.get("nchw") else 2) * (2 if kw.get("res") else 1)
    elif kind == "dw":
        k, s, c = kw.get("k", 3), kw.get("s", 1), kw["c"]
        x = torch.randn(B, hw, hw, c, device=DEV).to(torch.bfloat16)
        w = torch.randn(k * k, c, device=DEV).to(torch.bfloat16)
        b = torch.randn(c, device=DEV)
        y = torch.empty(B, (hw + s - 1) // s, (hw + s - 1) // s, c, device=DEV, dtype=torch.bfloat16)
        op.kind, op.k, op.stride, op.act = N.OP_DW, k, s, kw.get("act", 1)
        op.src, op.dst = view(x), view(y)
        op.w, op.bias = w.data_ptr(), b.data_ptr()
        keep += [x, w, b, y]
        byts = (x.numel() + y.numel()) * 2
    elif kind == "dwpw":
        c, cout = kw["c"], kw["cout"]
        x = torch.randn(B, hw, hw, c, device=DEV).to(torch.bfloat16)
        dww = (torch.randn(9, c, device=DEV) / 3).to(torch.bfloat16)
        dwb = torch.randn(c, device=DEV)
        w = (torch.randn(cout, 1, 1, c, device=DEV) / math.sqrt(c)).to(torch.bfloat16)
        b = torch.randn(cout, device=DEV)
        y = torch.empty(B, hw, hw, cout, device=DEV, dtype=torch.bfloat16)
        op.kind, op.k, op.stride, op.act, op.pre_k, op.pre_act = N.OP_DWPW, 1, 1, kw.get("act", 1), 3, 1
        op.src, op.dst = view(x), view(y)
        op.w, op.bias, op.pre_w, op.pre_bias = w.data_ptr(), b.data_ptr(), dww.data_ptr(), dwb.data_ptr()
        keep += [x, dww, dwb, w, b, y]
        flops = 2 * B * hw * hw * cout * c
        byts = (x.numel() + y.numel()) * 2
    elif kind == "pool":
        c = kw["c"]
        buf = torch.randn(B, hw, hw, 4 * c, device=DEV).to(torch.bfloat16)
        op.kind = N.OP_POOL
        op.src, op.dst = view(buf, 0, c), view(buf, c, 3 * c)
        keep.append(buf)
        byts = buf.numel() * 2
    elif kind == "attn":
        nh, kd, hd = kw.get("nh", 4), kw.get("kd", 32), kw.get("hd", 64)
        q = torch.randn(B, hw, hw, 2 * nh * kd + nh * hd, device=DEV).to(torch.bfloat16)
        o = torch.empty(B, hw, hw, nh * hd, device=DEV, dtype=torch.bfloat16)
        op.kind, op.nh, op.kdp, op.hd, op.scale = N.OP_ATTN, nh, kd, hd, kd ** -0.5
        op.src, op.dst = view(q), view(o)
        keep += [q, o]
        n = hw * hw
        flops = 2 * B * nh * n * n * (kd + hd)
        byts = (q.numel() + o.numel()) * 2
    elif kind == "stem":
        cout = kw.get("cout", 32)
        x = torch.rand(B, 3, hw, hw, device=DEV) * 255
        w = torch.randn(cout, 27, device=DEV)
        b = torch.randn(cout, device=DEV)
        y = torch.empty(B, hw // 2, hw // 2, cout, device=DEV, dtype=torch.bfloat16)
        op.kind, op.k, op.stride, op.act = N.OP_STEM, 3, 2, 1
        for j in range(3):
            op.sub[j], op.div[j] = 0.0, 255.0
        op.dst = view(y)
        op.w, op.bias, op.nchw = w.data_ptr(), b.data_ptr(), x.data_ptr()
        keep += [x, w, b, y]
        flops = 2 * B * (hw // 2) ** 2 * cout * 27
        byts = x.numel() * 4 + y.numel() * 2
    else:
        raise SystemExit(f"unknown op kind {kind}")
    lib = N.lib()
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    times = []
    for i in range(iters + 1):
        flush.zero_()   # evict L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N.check(lib.ly_launch(C.byref(op), stream), spec)
        e1.record()
        torch.cuda.synchronize()
        if i:
            times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)
    print(f"{spec:60s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s  {byts / ms / 1e6:8.0f} GB/s", flush=True)


if __name__ == "__main__":
    for s in sys.argv[1:]:
        run(s)
