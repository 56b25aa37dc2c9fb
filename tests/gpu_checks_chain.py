"""GPU parity checks of LY_OP_CHAIN (fused conv chains with shared-memory intermediates) through the C ABI,
against torch fp32 math on the quantised operands with bf16-rounded intermediates (tests/plan_interp.run_chain)."""
from __future__ import annotations

import ctypes as C
import math

import torch

from leanyolo_b200 import _native as N
from gpu_checks import DEV, launch, relmax, view
from plan_interp import run_chain


def _bf16(t):
    return t.to(torch.bfloat16).float()


def make_chain(regions, n_in, stages, keep, stride0=1):
    """ctypes ly_chain from the python spec; device tensors are appended to ``keep``."""
    ch = N.LyChain()
    ch.n_regions, ch.n_in, ch.n_stages = len(regions), n_in, len(stages)
    ch.reserved = 2 if stride0 == 2 else 0
    for i, c in enumerate(regions):
        ch.region_c[i] = c
    for i, st in enumerate(stages):
        s = ch.st[i]
        s.k, s.act, s.cout, s.n_src = st["k"], int(st["act"]), st["cout"], len(st["src"])
        for j, (r, c0, c) in enumerate(st["src"]):
            s.src[j] = N.LyChainBlk(r, c0, c)
        s.dst = N.LyChainBlk(*(st["dst"] if st.get("dst") is not None else (-1, 0, 0)))
        s.res = N.LyChainBlk(*(st["res"] if st.get("res") is not None else (-1, 0, 0)))
        w_d = st["w"].to(torch.bfloat16).to(DEV).contiguous()
        b_d = st["b"].float().to(DEV).contiguous()
        keep += [w_d, b_d]
        s.w, s.bias = w_d.data_ptr(), b_d.data_ptr()
    return ch


def _stage(g, k, cin, cout, act, src, dst=None, res=None):
    w = _bf16(torch.randn(cout, k, k, cin, generator=g) / math.sqrt(cin * k * k))
    b = torch.randn(cout, generator=g) * 0.5
    return dict(k=k, act=act, cout=cout, src=src, dst=dst, res=res, w=w, b=b)


def spec_c2f(g, c, shortcut=True):
    """C2f with one Bottleneck (layers.py:129-173): regions X(2c) Y(2c) T(c)."""
    return [2 * c, 2 * c, c], 1, [
        _stage(g, 1, 2 * c, 2 * c, True, [(0, 0, 2 * c)], dst=(1, 0, 2 * c)),
        _stage(g, 3, c, c, True, [(1, c, c)], dst=(2, 0, c)),
        _stage(g, 3, c, c, True, [(2, 0, c)], dst=(2, 0, c), res=(1, c, c) if shortcut else None),
        _stage(g, 1, 3 * c, 2 * c, True, [(1, 0, c), (1, c, c), (2, 0, c)]),
    ]


def spec_tail(g, c, cout, act_last=False, cmid=None):
    """3x3 c->cmid (+SiLU) -> 1x1 cmid->cout (+bias): the tail of a v10Detect regression stack (head.py:86-92); with
    check_chain(stride0=2) and cmid != c the backbone's cv1 (3x3 / s2) -> c2.cv1 (1x1) pair."""
    cm = cmid or c
    return [c, cm], 1, [
        _stage(g, 3, c, cm, True, [(0, 0, c)], dst=(1, 0, cm)),
        _stage(g, 1, cm, cout, act_last, [(1, 0, cm)]),
    ]


def spec_single(g, k, cin, cout):
    return [cin], 1, [_stage(g, k, cin, cout, True, [(0, 0, cin)])]


def spec_two_in(g):
    """two 64-channel input regions -> 1x1 128->64 -> 3x3 64->32"""
    return [64, 64, 64], 2, [
        _stage(g, 1, 128, 64, True, [(0, 0, 64), (1, 0, 64)], dst=(2, 0, 64)),
        _stage(g, 3, 64, 32, False, [(2, 0, 64)]),
    ]


SPECS = {"c2f": spec_c2f, "tail": spec_tail, "single": spec_single, "two_in": spec_two_in}


def build_tail_op(B, H, W, c, cout, seed=0, cmid=None, stride0=1):
    """A ready-to-launch 3x3 -> 1x1 op on random data (timing tool: tools/bench_b2b.py): stride 1 with a public NCHW output (a
    regression tail) or stride 2 with an NHWC bf16 output (backbone cv1 -> c2.cv1)."""
    g = torch.Generator().manual_seed(seed)
    regions, n_in, stages = SPECS["tail"](g, c=c, cout=cout, cmid=cmid, act_last=stride0 == 2)
    keep = []
    ch = make_chain(regions, n_in, stages, keep, stride0)
    x = torch.randn(B, H, W, c, device=DEV).to(torch.bfloat16)
    op = N.LyOp()
    op.kind, op.dtype, op.B, op.k, op.stride, op.act, op.ext_slot = N.OP_CHAIN, N.LY_BF16, B, 1, 1, 0, -1
    op.src = view(x, 0, c)
    op.chain = C.pointer(ch)
    if stride0 == 2:
        out = torch.zeros(B, H // 2, W // 2, cout, device=DEV, dtype=torch.bfloat16)
        op.dst = view(out, 0, cout)
    else:
        out = torch.zeros(B, cout, H, W, device=DEV)
        op.nchw, op.nchw_ctot, op.nchw_c0, op.nchw_c = out.data_ptr(), cout, 0, cout
    keep += [ch, x, out]
    return op, keep


def check_chain(kind="c2f", B=2, H=40, W=40, src_off=0, src_extra=0, dst_off=0, dst_extra=0, nchw=False, nchw_c=None,
                seed=0, tol=2e-2, stride0=1, **kw):
    g = torch.Generator().manual_seed(seed)
    regions, n_in, stages = SPECS[kind](g, **kw)
    cin = sum(regions[:n_in])
    cout = stages[-1]["cout"]
    xs = _bf16(torch.randn(B, H, W, src_off + cin + src_extra, generator=g))
    ref = run_chain(regions, n_in, stages, xs[..., src_off:src_off + cin], quant=_bf16, stride0=stride0)
    d0 = _bf16(torch.randn(B, H // stride0, W // stride0, dst_off + cout + dst_extra, generator=g))
    keep = []
    ch = make_chain(regions, n_in, stages, keep, stride0)
    x_d, d_d = xs.to(torch.bfloat16).to(DEV), d0.to(torch.bfloat16).to(DEV)
    op = N.LyOp()
    op.kind, op.dtype, op.B, op.k, op.stride, op.act, op.ext_slot = N.OP_CHAIN, N.LY_BF16, B, 1, 1, 0, -1
    op.src = view(x_d, src_off, cin)
    op.chain = C.pointer(ch)
    out_nchw = None
    if nchw:
        cr = nchw_c or cout
        out_nchw = torch.zeros(B, cr + 3, H // stride0, W // stride0, device=DEV)
        op.nchw, op.nchw_ctot, op.nchw_c0, op.nchw_c = out_nchw.data_ptr(), cr + 3, 2, cr
    else:
        op.dst = view(d_d, dst_off, cout)
    launch(op)
    if nchw:
        e = relmax(out_nchw[:, 2:2 + cr].permute(0, 2, 3, 1).cpu(), ref[..., :cr])
        assert float(out_nchw[:, :2].abs().max()) == 0 and float(out_nchw[:, 2 + cr:].abs().max()) == 0, "nchw overrun"
    else:
        got = d_d.float().cpu()
        e = relmax(got[..., dst_off:dst_off + cout], ref)
        if dst_off:
            assert torch.equal(got[..., :dst_off], d0[..., :dst_off]), "clobbered channels before slice"
        if dst_extra:
            assert torch.equal(got[..., dst_off + cout:], d0[..., dst_off + cout:]), "clobbered channels after slice"
    assert e < tol, f"chain {kind} relmax {e:.3e} >= {tol}"
    return {"relmax": e}
