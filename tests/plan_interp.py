"""TEST INFRASTRUCTURE: a torch-CPU interpreter of the lowered op list.

Executes ``PlanBuilder.ops`` exactly as the C ABI defines them (NHWC buffers, packed
weights, channel-offset views, residual-after-activation) so the *lowering* — BN
folding, RepVGGDW merge, qkv re-ordering, concat/split offsets, in-place PSA updates —
can be checked against the oracle without a GPU.  Never imported by the product.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from leanyolo_b200.plan import PlanBuilder


def run_plan(pb: PlanBuilder, x: torch.Tensor, quant=None):
    """x: [B,3,H,W] fp32.  quant: optional fn applied to every stored activation
    (e.g. bf16 round-trip) to emulate the storage dtype."""
    q = quant or (lambda t: t)
    w_blob, b_blob = pb.finalize_params()
    w_blob = w_blob.float()
    B = x.shape[0]
    bufs = {b.id: torch.zeros(B, b.H, b.W, b.C) for b in pb.bufs}
    outs = {k: torch.zeros(B, c, h, w) for k, (c, h, w) in pb.outputs.items()}

    def rd(v):
        return bufs[v.buf.id][..., v.c0:v.c0 + v.c]

    def wr(v, t):
        bufs[v.buf.id][..., v.c0:v.c0 + v.c] = q(t)

    for op in pb.ops:
        if op.kind == "stem":
            cp = op.dst.c
            w = b_blob[op.w_off:op.w_off + cp * 27].view(cp, 3, 3, 3).permute(0, 3, 1, 2)   # [co][ky][kx][ci] -> OIHW
            bias = b_blob[op.b_off:op.b_off + cp]
            sub = torch.tensor(op.extra["sub"]).view(1, 3, 1, 1)
            div = torch.tensor(op.extra["div"]).view(1, 3, 1, 1)
            y = F.silu(F.conv2d((x - sub) / div, w, bias, 2, 1))
            wr(op.dst, y.permute(0, 2, 3, 1))
        elif op.kind == "conv":
            cin, k = op.src.c, op.k
            cp = op.extra["cpad"]
            w = w_blob[op.w_off:op.w_off + cp * k * k * cin].view(cp, k, k, cin).permute(0, 3, 1, 2)
            bias = b_blob[op.b_off:op.b_off + cp]
            y = F.conv2d(rd(op.src).permute(0, 3, 1, 2), w, bias, op.stride, k // 2)
            if op.extra.get("up") is not None:      # half-resolution pre-activation addend
                y = y + F.interpolate(rd(op.extra["up"]).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
            if op.act:
                y = F.silu(y)
            y = y.permute(0, 2, 3, 1)
            if op.res is not None:
                y = y + rd(op.res)
            if op.dst is not None:
                wr(op.dst, y)
            if op.nchw is not None:
                name, level, c0, c, _ = op.nchw
                outs[(name, level)][:, c0:c0 + c] = y[..., :c].permute(0, 3, 1, 2)
        elif op.kind == "chain":
            stages = []
            for st in op.extra["stages"]:
                n = st["cout"] * st["k"] * st["k"] * st["cin"]
                stages.append(dict(k=st["k"], act=st["act"], cout=st["cout"], src=st["src"], dst=st["dst"], res=st["res"],
                                   w=w_blob[st["w_off"]:st["w_off"] + n].view(st["cout"], st["k"], st["k"], st["cin"]),
                                   b=b_blob[st["b_off"]:st["b_off"] + st["cout"]]))
            y = run_chain(op.extra["regions"], op.extra["n_in"], stages, rd(op.src), quant=q, stride0=op.extra.get("stride0", 1))
            if op.dst is not None:
                wr(op.dst, y)
            if op.nchw is not None:
                name, level, c0, c, _ = op.nchw
                outs[(name, level)][:, c0:c0 + c] = y[..., :c].permute(0, 3, 1, 2)
        elif op.kind == "dwpw":
            c, cp = op.src.c, op.extra["cpad"]
            dww = w_blob[op.extra["pre_w_off"]:op.extra["pre_w_off"] + 9 * c].view(3, 3, c).permute(2, 0, 1).unsqueeze(1)
            dwb = b_blob[op.extra["pre_b_off"]:op.extra["pre_b_off"] + c]
            y = F.conv2d(rd(op.src).permute(0, 3, 1, 2), dww, dwb, 1, 1, 1, c)
            if op.extra["pre_act"]:
                y = F.silu(y)
            y = q(y)   # the fused kernel hands the depthwise result to the GEMM in the storage dtype
            w = w_blob[op.w_off:op.w_off + cp * c].view(cp, 1, 1, c).permute(0, 3, 1, 2)
            y = F.conv2d(y, w, b_blob[op.b_off:op.b_off + cp])
            if op.act:
                y = F.silu(y)
            y = y.permute(0, 2, 3, 1)
            if op.dst is not None:
                wr(op.dst, y)
            if op.nchw is not None:
                name, level, c0, cc, _ = op.nchw
                outs[(name, level)][:, c0:c0 + cc] = y[..., :cc].permute(0, 3, 1, 2)
        elif op.kind == "dw":
            c, k = op.src.c, op.k
            w = w_blob[op.w_off:op.w_off + k * k * c].view(k, k, c).permute(2, 0, 1).unsqueeze(1)
            bias = b_blob[op.b_off:op.b_off + c]
            y = F.conv2d(rd(op.src).permute(0, 3, 1, 2), w, bias, op.stride, k // 2, 1, c)
            if op.act:
                y = F.silu(y)
            y = y.permute(0, 2, 3, 1)
            if op.res is not None:
                y = y + rd(op.res)
            wr(op.dst, y)
        elif op.kind == "pool":
            c = op.src.c
            t = rd(op.src).permute(0, 3, 1, 2)
            y1 = F.max_pool2d(t, 5, 1, 2)
            y2 = F.max_pool2d(y1, 5, 1, 2)
            y3 = F.max_pool2d(y2, 5, 1, 2)
            wr(op.dst, torch.cat([y1, y2, y3], 1).permute(0, 2, 3, 1))
        elif op.kind == "up":
            t = rd(op.src).permute(0, 3, 1, 2)
            wr(op.dst, F.interpolate(t, scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1))
        elif op.kind == "attn":
            nh, kdp, hd, scale = op.attn
            t = rd(op.src)
            Bn, H, W, _ = t.shape
            n = H * W
            t = t.reshape(Bn, n, -1)
            qq = t[..., :nh * kdp].view(Bn, n, nh, kdp).permute(0, 2, 1, 3)
            kk = t[..., nh * kdp:2 * nh * kdp].view(Bn, n, nh, kdp).permute(0, 2, 1, 3)
            vv = t[..., 2 * nh * kdp:2 * nh * kdp + nh * hd].view(Bn, n, nh, hd).permute(0, 2, 1, 3)
            att = ((qq @ kk.transpose(-1, -2)) * scale).softmax(-1)
            o = (att @ vv).permute(0, 2, 1, 3).reshape(Bn, H, W, nh * hd)
            wr(op.dst, o)
        elif op.kind == "export":
            name, level, c0, c, _ = op.nchw
            outs[(name, level)][:] = rd(op.src)[..., :c].permute(0, 3, 1, 2)
        else:
            raise NotImplementedError(op.kind)
    return outs


def run_chain(regions, n_in, stages, x, quant=None, stride0=1):
    """Reference semantics of LY_OP_CHAIN (include/leanyolo_b200.h) on whole images.

    regions: channels per region; x: [B,H,W,sum(regions[:n_in])] NHWC fp32 (already in the storage
    precision); stages: dicts {k, act, cout, src: [(region, c0, c)], dst: (region, c0, c) | None,
    res: (region, c0, c) | None, w: [cout,k,k,Ctot] fp32, b: [cout] fp32}.  Intermediates are stored
    through ``quant`` (the kernel keeps them as bf16 in shared memory).  Returns the last stage's
    [B,H,W,cout] fp32 result (not quantised)."""
    q = quant or (lambda t: t)
    B, H, W, _ = x.shape
    Ho, Wo = H // stride0, W // stride0      # stride0 = 2: the first stage is a stride-2 3x3, everything after it is half size
    reg = [torch.zeros(B, H, W, c) if i < n_in else torch.zeros(B, Ho, Wo, c) for i, c in enumerate(regions)]
    off = 0
    for i in range(n_in):
        reg[i] = x[..., off:off + regions[i]].clone()
        off += regions[i]
    y = None
    for si, st in enumerate(stages):
        xin = torch.cat([reg[r][..., c0:c0 + c] for r, c0, c in st["src"]], -1).permute(0, 3, 1, 2)
        w = st["w"].float().permute(0, 3, 1, 2)
        y = F.conv2d(xin, w, st["b"].float(), stride0 if si == 0 else 1, st["k"] // 2)
        if st["act"]:
            y = F.silu(y)
        y = y.permute(0, 2, 3, 1)
        if st.get("res") is not None:
            r, c0, c = st["res"]
            y = y + reg[r][..., c0:c0 + c]
        if st.get("dst") is not None:
            r, c0, c = st["dst"]
            reg[r][..., c0:c0 + c] = q(y)
    return y
