"""GPU parity checks shared by the pytest suite (-m gpu) and tools/gpu_diag.py.

Every check calls the CUDA path through the C ABI (ctypes) and compares with either
plain torch fp32 math on the same (already quantised) operands or the CPU oracle.
Each returns a dict of error figures and raises AssertionError on failure.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn.functional as F

from leanyolo_b200 import _native as N

DEV = "cuda"


def _dt(dtype):
    return (N.LY_BF16, torch.bfloat16) if dtype == "bf16" else (N.LY_F32, torch.float32)


def view(t: torch.Tensor, c0: int = 0, c: int | None = None) -> N.LyView:
    """t: NHWC tensor [B,H,W,Ctot]"""
    return N.LyView(t.data_ptr(), t.shape[1], t.shape[2], t.shape[3], c0, t.shape[3] - c0 if c is None else c)


def launch(op: N.LyOp) -> None:
    stream = torch.cuda.current_stream().cuda_stream
    N.check(N.lib().ly_launch(C.byref(op), C.c_void_p(stream)), "ly_launch")
    torch.cuda.synchronize()


def relmax(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.float() - b.float()).abs().max() / b.float().abs().max().clamp(min=1e-12))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.float() - b.float()).norm() / b.float().norm().clamp(min=1e-12))


# ------------------------------------------------------------------------------- dense conv
def check_conv(dtype="bf16", impl="auto", B=2, H=16, W=16, cin=64, cout=64, k=3, stride=1, act=True, res=False,
               src_off=0, src_extra=0, dst_off=0, dst_extra=0, nchw=False, inplace_res=False, seed=0, nchw_c=None,
               tol=None, up=False):
    code, tdt = _dt(dtype)
    g = torch.Generator().manual_seed(seed)
    Ho, Wo = H // stride, W // stride
    xs = torch.randn(B, H, W, src_off + cin + src_extra, generator=g).to(tdt)
    w = (torch.randn(cout, k, k, cin, generator=g) / math.sqrt(cin * k * k)).to(tdt)
    bias = torch.randn(cout, generator=g)
    dst_tot = dst_off + cout + dst_extra
    d0 = torch.randn(B, Ho, Wo, dst_tot, generator=g).to(tdt)
    rs = torch.randn(B, Ho, Wo, cout, generator=g).to(tdt) if res else None
    # reference on the quantised operands, fp32 math
    xin = xs[..., src_off:src_off + cin].float().permute(0, 3, 1, 2)
    y = F.conv2d(xin, w.float().permute(0, 3, 1, 2), bias, stride, k // 2)
    ups = None
    if up:   # half-resolution pre-activation addend, upsampled x2 on the fly (channel slice of a wider buffer)
        ups = torch.randn(B, Ho // 2, Wo // 2, 16 + cout, generator=g).to(tdt)
        y = y + F.interpolate(ups[..., 16:].float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    if act:
        y = F.silu(y)
    y = y.permute(0, 2, 3, 1)
    if inplace_res:
        y = y + d0[..., dst_off:dst_off + cout].float()
    elif res:
        y = y + rs.float()

    x_d, w_d, b_d, d_d = xs.to(DEV), w.to(DEV).contiguous(), bias.to(DEV), d0.to(DEV)
    op = N.LyOp()
    op.kind, op.dtype, op.B, op.k, op.stride, op.act = N.OP_CONV, code, B, k, stride, int(act)
    op.impl = N.IMPL_SIMT if impl == "simt" else N.IMPL_AUTO
    op.src = view(x_d, src_off, cin)
    op.w, op.bias = w_d.data_ptr(), b_d.data_ptr()
    op.ext_slot = -1
    out_nchw = None
    if nchw:
        cr = nchw_c or cout
        out_nchw = torch.zeros(B, cr + 3, Ho, Wo, device=DEV)
        op.nchw, op.nchw_ctot, op.nchw_c0, op.nchw_c = out_nchw.data_ptr(), cr + 3, 2, cr
    else:
        op.dst = view(d_d, dst_off, cout)
    if inplace_res:
        op.res = view(d_d, dst_off, cout)
    elif res:
        r_d = rs.to(DEV)
        op.res = view(r_d, 0, cout)
    if up:
        u_d = ups.to(DEV)
        op.up = view(u_d, 16, cout)
    launch(op)
    tol = tol if tol is not None else (2e-2 if dtype == "bf16" else 1e-4)
    if nchw:
        got = out_nchw[:, 2:2 + cr].permute(0, 2, 3, 1).cpu()
        e = relmax(got, y[..., :cr])
        assert float(out_nchw[:, :2].abs().max()) == 0 and float(out_nchw[:, 2 + cr:].abs().max()) == 0, "nchw overrun"
    else:
        got = d_d.cpu()
        e = relmax(got[..., dst_off:dst_off + cout], y)
        # untouched channels of the concat buffer must be preserved bit-for-bit
        if dst_off:
            assert torch.equal(got[..., :dst_off], d0[..., :dst_off]), "clobbered channels before slice"
        if dst_extra:
            assert torch.equal(got[..., dst_off + cout:], d0[..., dst_off + cout:]), "clobbered channels after slice"
    assert e < tol, f"conv relmax {e:.3e} >= {tol}"
    return {"relmax": e}


# ------------------------------------------------------------------------------- depthwise
def check_dw(dtype="bf16", B=2, H=12, W=12, c=64, k=3, stride=1, act=True, res=False, seed=0):
    code, tdt = _dt(dtype)
    g = torch.Generator().manual_seed(seed)
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    xs = torch.randn(B, H, W, c + 16, generator=g).to(tdt)
    w = (torch.randn(k * k, c, generator=g) / k).to(tdt)
    bias = torch.randn(c, generator=g)
    rs = torch.randn(B, Ho, Wo, c, generator=g).to(tdt) if res else None
    y = F.conv2d(xs[..., 8:8 + c].float().permute(0, 3, 1, 2), w.float().t().reshape(c, 1, k, k), bias, stride, k // 2, 1, c)
    if act:
        y = F.silu(y)
    y = y.permute(0, 2, 3, 1)
    if res:
        y = y + rs.float()
    x_d, w_d, b_d = xs.to(DEV), w.to(DEV), bias.to(DEV)
    d_d = torch.zeros(B, Ho, Wo, c, device=DEV, dtype=tdt)
    op = N.LyOp()
    op.kind, op.dtype, op.B, op.k, op.stride, op.act, op.ext_slot = N.OP_DW, code, B, k, stride, int(act), -1
    op.src, op.dst = view(x_d, 8, c), view(d_d)
    op.w, op.bias = w_d.data_ptr(), b_d.data_ptr()
    if res:
        r_d = rs.to(DEV)
        op.res = view(r_d)
    launch(op)
    e = relmax(d_d.cpu(), y)
    assert e < (1e-2 if dtype == "bf16" else 1e-5), f"dw relmax {e:.3e}"
    return {"relmax": e}


def check_dwpw(B=2, H=20, W=20, c=128, cout=128, dw_act=True, act=True, src_off=0, src_extra=0, dst_off=0, dst_extra=0,
               nchw=False, nchw_c=None, seed=0):
    """fused depthwise 3x3 -> 1x1 (bf16 only) against torch fp32 on the quantised operands, with the
    depthwise result rounded to bf16 like the kernel hands it to the GEMM."""
    g = torch.Generator().manual_seed(seed)
    tdt = torch.bfloat16
    xs = torch.randn(B, H, W, src_off + c + src_extra, generator=g).to(tdt)
    dww = (torch.randn(9, c, generator=g) / 3).to(tdt)
    dwb = torch.randn(c, generator=g) * 0.2
    w = (torch.randn(cout, 1, 1, c, generator=g) / math.sqrt(c)).to(tdt)
    bias = torch.randn(cout, generator=g)
    y = F.conv2d(xs[..., src_off:src_off + c].float().permute(0, 3, 1, 2), dww.float().t().reshape(c, 1, 3, 3), dwb, 1, 1, 1, c)
    if dw_act:
        y = F.silu(y)
    y = y.to(tdt).float()
    y = F.conv2d(y, w.float().permute(0, 3, 1, 2), bias)
    if act:
        y = F.silu(y)
    y = y.permute(0, 2, 3, 1)
    dst_tot = dst_off + cout + dst_extra
    d0 = torch.randn(B, H, W, dst_tot, generator=g).to(tdt)
    x_d, dww_d, dwb_d, w_d, b_d, d_d = xs.to(DEV), dww.to(DEV), dwb.to(DEV), w.to(DEV).contiguous(), bias.to(DEV), d0.to(DEV)
    op = N.LyOp()
    op.kind, op.dtype, op.B, op.k, op.stride, op.act, op.ext_slot = N.OP_DWPW, N.LY_BF16, B, 1, 1, int(act), -1
    op.pre_k, op.pre_act = 3, int(dw_act)
    op.pre_w, op.pre_bias = dww_d.data_ptr(), dwb_d.data_ptr()
    op.src = view(x_d, src_off, c)
    op.w, op.bias = w_d.data_ptr(), b_d.data_ptr()
    out_nchw = None
    if nchw:
        cr = nchw_c or cout
        out_nchw = torch.zeros(B, cr + 3, H, W, device=DEV)
        op.nchw, op.nchw_ctot, op.nchw_c0, op.nchw_c = out_nchw.data_ptr(), cr + 3, 2, cr
    else:
        op.dst = view(d_d, dst_off, cout)
    launch(op)
    if nchw:
        e = relmax(out_nchw[:, 2:2 + cr].permute(0, 2, 3, 1).cpu(), y[..., :cr])
        assert float(out_nchw[:, :2].abs().max()) == 0 and float(out_nchw[:, 2 + cr:].abs().max()) == 0, "nchw overrun"
    else:
        got = d_d.cpu()
        e = relmax(got[..., dst_off:dst_off + cout], y)
        if dst_off:
            assert torch.equal(got[..., :dst_off], d0[..., :dst_off]), "clobbered channels before slice"
        if dst_extra:
            assert torch.equal(got[..., dst_off + cout:], d0[..., dst_off + cout:]), "clobbered channels after slice"
    assert e < 2e-2, f"dwpw relmax {e:.3e}"
    return {"relmax": e}


def check_pool(dtype="bf16", B=2, H=20, W=20, c=32, seed=0):
    code, tdt = _dt(dtype)
    g = torch.Generator().manual_seed(seed)
    buf = torch.randn(B, H, W, 4 * c, generator=g).to(tdt)
    x = buf[..., :c].float().permute(0, 3, 1, 2)
    y1 = F.max_pool2d(x, 5, 1, 2)
    y2 = F.max_pool2d(y1, 5, 1, 2)
    y3 = F.max_pool2d(y2, 5, 1, 2)
    ref = torch.cat([x, y1, y2, y3], 1).permute(0, 2, 3, 1)
    b_d = buf.to(DEV)
    op = N.LyOp()
    op.kind, op.dtype, op.B, op.ext_slot = N.OP_POOL, code, B, -1
    op.src, op.dst = view(b_d, 0, c), view(b_d, c, 3 * c)
    launch(op)
    assert torch.equal(b_d.cpu().float(), ref), "sppf pool mismatch (must be exact)"
    return {"relmax": 0.0}


def check_up(dtype="bf16", B=2, H=5, W=7, c=32, seed=0):
    code, tdt = _dt(dtype)
    g = torch.Generator().manual_seed(seed)
    xs = torch.randn(B, H, W, c + 8, generator=g).to(tdt)
    d0 = torch.randn(B, 2 * H, 2 * W, c + 16, generator=g).to(tdt)
    x_d, d_d = xs.to(DEV), d0.to(DEV)
    op = N.LyOp()
    op.kind, op.dtype, op.B, op.ext_slot = N.OP_UP, code, B, -1
    op.src, op.dst = view(x_d, 8, c), view(d_d, 16, c)
    launch(op)
    ref = d0.clone()
    ref[..., 16:] = F.interpolate(xs[..., 8:].float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1).to(tdt)
    assert torch.equal(d_d.cpu(), ref), "upsample mismatch (must be exact)"
    return {"relmax": 0.0}


def check_attn(dtype="bf16", B=2, H=20, W=20, nh=4, kd=32, hd=64, seed=0):
    code, tdt = _dt(dtype)
    g = torch.Generator().manual_seed(seed)
    kdp = (kd + 7) // 8 * 8
    ctot = 2 * nh * kdp + nh * hd
    qkv = torch.randn(B, H, W, ctot, generator=g).to(tdt)
    if kdp != kd:  # padded key rows are exact zeros in the real pipeline
        t = qkv[..., :2 * nh * kdp].reshape(B, H, W, 2 * nh, kdp)
        t[..., kd:] = 0
        qkv[..., :2 * nh * kdp] = t.reshape(B, H, W, 2 * nh * kdp)
    n = H * W
    t = qkv.float().reshape(B, n, ctot)
    q = t[..., :nh * kdp].view(B, n, nh, kdp).permute(0, 2, 1, 3)
    kk = t[..., nh * kdp:2 * nh * kdp].view(B, n, nh, kdp).permute(0, 2, 1, 3)
    v = t[..., 2 * nh * kdp:].view(B, n, nh, hd).permute(0, 2, 1, 3)
    scale = kd ** -0.5
    ref = (((q @ kk.transpose(-1, -2)) * scale).softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, H, W, nh * hd)
    q_d = qkv.to(DEV)
    o_d = torch.zeros(B, H, W, nh * hd, device=DEV, dtype=tdt)
    op = N.LyOp()
    op.kind, op.dtype, op.B, op.ext_slot = N.OP_ATTN, code, B, -1
    op.nh, op.kdp, op.hd, op.scale = nh, kdp, hd, scale
    op.src, op.dst = view(q_d), view(o_d)
    launch(op)
    e = relmax(o_d.cpu(), ref)
    assert e < (1e-2 if dtype == "bf16" else 1e-4), f"attention relmax {e:.3e}"
    return {"relmax": e}


def check_stem(dtype="bf16", B=2, H=64, W=96, cout=32, sub=(0.0, 0.0, 0.0), div=(255.0, 255.0, 255.0), seed=0):
    code, tdt = _dt(dtype)
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 3, H, W, generator=g) * 255
    w = torch.randn(cout, 3, 3, 3, generator=g) * 0.3
    bias = torch.randn(cout, generator=g) * 0.1
    xn = (x - torch.tensor(sub).view(1, 3, 1, 1)) / torch.tensor(div).view(1, 3, 1, 1)
    ref = F.silu(F.conv2d(xn, w, bias, 2, 1)).permute(0, 2, 3, 1)
    cp = (cout + 15) // 16 * 16
    wp = torch.zeros(cp, 27)
    wp[:cout] = w.permute(0, 2, 3, 1).reshape(cout, 27)
    bp = torch.zeros(cp)
    bp[:cout] = bias
    x_d, w_d, b_d = x.to(DEV), wp.to(DEV), bp.to(DEV)
    d_d = torch.zeros(B, H // 2, W // 2, cp, device=DEV, dtype=tdt)
    op = N.LyOp()
    op.kind, op.dtype, op.B, op.k, op.stride, op.act, op.ext_slot = N.OP_STEM, code, B, 3, 2, 1, -1
    for j in range(3):
        op.sub[j], op.div[j] = sub[j], div[j]
    op.dst = view(d_d)
    op.w, op.bias, op.nchw = w_d.data_ptr(), b_d.data_ptr(), x_d.data_ptr()
    launch(op)
    e = relmax(d_d.cpu()[..., :cout], ref)
    assert e < (1e-2 if dtype == "bf16" else 1e-5), f"stem relmax {e:.3e}"
    return {"relmax": e}


def check_export_import(dtype="bf16", B=2, H=6, W=10, c=48, seed=0):
    code, tdt = _dt(dtype)
    g = torch.Generator().manual_seed(seed)
    xs = torch.randn(B, H, W, c + 16, generator=g).to(tdt)
    x_d = xs.to(DEV)
    out = torch.zeros(B, c, H, W, device=DEV)
    op = N.LyOp()
    op.kind, op.dtype, op.B, op.ext_slot = N.OP_EXPORT, code, B, -1
    op.src = view(x_d, 16, c)
    op.nchw, op.nchw_ctot, op.nchw_c0, op.nchw_c = out.data_ptr(), c, 0, c
    launch(op)
    assert torch.equal(out.cpu(), xs[..., 16:].float().permute(0, 3, 1, 2))
    back = torch.zeros(B, H, W, c, device=DEV, dtype=tdt)
    op2 = N.LyOp()
    op2.kind, op2.dtype, op2.B, op2.ext_slot = N.OP_IMPORT, code, B, -1
    op2.dst = view(back)
    op2.nchw, op2.nchw_ctot, op2.nchw_c0, op2.nchw_c = out.data_ptr(), c, 0, c
    launch(op2)
    assert torch.equal(back.cpu(), xs[..., 16:])
    return {"relmax": 0.0}
