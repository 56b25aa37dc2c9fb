"""Parameter tree of YOLOv10 (reference key names) + lowering to the kernel plan.

The modules below only *hold* parameters, under exactly the attribute names the
reference uses, so that any reference ``state_dict`` loads strictly and ours
saves back bit-identically (reference contract: SURVEY §8(b);
leanyolo/tests/test_state_dict_roundtrip.py).  They do not compute anything in
PyTorch: each block has an ``emit(pb, src, dst)`` method that lowers it to flat
kernel ops on NHWC buffers (``leanyolo_b200.plan.PlanBuilder``).  Concat / split
/ residual never become ops — they are channel-offset views into shared buffers
and epilogue flags.

Reference semantics followed (file:line relative to the reference root):
``Conv`` layers.py:51-88 · ``Bottleneck`` :91-126 · ``C2f`` :129-173 · ``SPPF``
:176-217 · ``CIB``/``RepVGGDW`` :243-300 · ``C2fCIB`` :303-335 · ``Attention``
:338-380 · ``PSA`` :383-425 · ``SCDown`` :428-458 · backbone.py:68-106 ·
neck.py:82-129 · head.py:75-122.
"""
from __future__ import annotations

import copy
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from .plan import ShapeOnly
from .variants import REG_MAX, Variant

BN_EPS = 1e-3  # layers.py:84 (not torch's 1e-5 default)


class ConvBN(nn.Module):
    """Holder for the reference ``Conv`` (conv.weight + bn.*), act = SiLU or none."""

    def __init__(self, cin: int, cout: int, k: int = 1, s: int = 1, g: int = 1, act: bool = True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, s, k // 2, groups=g, bias=False)
        self.bn = nn.BatchNorm2d(cout, eps=BN_EPS, momentum=0.03)
        self.k, self.s, self.g, self.act = k, s, g, act

    def folded(self, pb=None):
        """BN folded into the conv (fp64): w' = w*gamma/sqrt(var+eps), b' = beta - mean*gamma/sqrt(var+eps).
        With a dry PlanBuilder (packed parameters already exist: later shapes of the same model, or the on-disk
        pack cache) only the SHAPES are produced (``plan.ShapeOnly``): no fp64 arithmetic, no copies."""
        if pb is not None and pb.dry:
            return ShapeOnly(self.conv.weight.shape), ShapeOnly(self.conv.weight.shape[:1])
        w = self.conv.weight.detach().double().cpu()
        bn = self.bn
        scale = bn.weight.detach().double().cpu() / torch.sqrt(bn.running_var.detach().double().cpu() + BN_EPS)
        return w * scale.view(-1, 1, 1, 1), bn.bias.detach().double().cpu() - bn.running_mean.detach().double().cpu() * scale

    def emit(self, pb, src, dst=None, res=None):
        w, b = self.folded(pb)
        if self.g == 1:
            return pb.conv(src, w, b, k=self.k, stride=self.s, act=self.act, dst=dst, res=res)
        assert self.g == self.conv.in_channels == self.conv.out_channels, "only depthwise groups"
        return pb.dwconv(src, w, b, k=self.k, stride=self.s, act=self.act, dst=dst, res=res)


def emit_dw_pw(pb, dw: "ConvBN", pw: "ConvBN", src, dst=None):
    """dw3x3 (+SiLU) followed by a 1x1: one fused launch when the tensor-core kernel takes the
    pair (bf16, stride 1, C % 64 == 0, Cout <= 256), otherwise the two separate ops."""
    if dw.g > 1 and pw.g == 1 and pw.k == 1 and pw.s == 1 and pb.dwpw_fusable(src, dw.conv.in_channels, pw.conv.out_channels, dw.k, dw.s):
        dw_w, dw_b = dw.folded(pb)
        pw_w, pw_b = pw.folded(pb)
        return pb.dwpw(src, dw_w, dw_b, pw_w, pw_b, dw_act=dw.act, act=pw.act, dst=dst)
    return pw.emit(pb, dw.emit(pb, src), dst)


class Bottleneck(nn.Module):
    def __init__(self, c: int, shortcut: bool):
        super().__init__()
        self.cv1 = ConvBN(c, c, 3)
        self.cv2 = ConvBN(c, c, 3)
        self.add = shortcut

    def emit(self, pb, src, dst):
        return self.cv2.emit(pb, self.cv1.emit(pb, src), dst, res=src if self.add else None)


class RepVGGDW(nn.Module):
    """SiLU(BN7(dw7x7(x)) + BN3(dw3x3(x))) re-parameterised to ONE dw7x7 at plan time."""

    def __init__(self, c: int):
        super().__init__()
        self.conv = ConvBN(c, c, 7, g=c, act=False)
        self.conv1 = ConvBN(c, c, 3, g=c, act=False)

    def emit(self, pb, src, dst=None, res=None):
        w7, b7 = self.conv.folded(pb)
        w3, b3 = self.conv1.folded(pb)
        w = w7.clone()
        w[:, :, 2:5, 2:5] += w3
        return pb.dwconv(src, w, b7 + b3, k=7, stride=1, act=True, dst=dst, res=res)


class CIB(nn.Module):
    def __init__(self, c: int, lk: bool):
        super().__init__()
        self.cv1 = nn.Sequential(
            ConvBN(c, c, 3, g=c),
            ConvBN(c, 2 * c, 1),
            RepVGGDW(2 * c) if lk else ConvBN(2 * c, 2 * c, 3, g=2 * c),
            ConvBN(2 * c, c, 1),
            ConvBN(c, c, 3, g=c),
        )

    def emit(self, pb, src, dst):
        y = emit_dw_pw(pb, self.cv1[0], self.cv1[1], src)
        if isinstance(self.cv1[2], RepVGGDW):
            y = self.cv1[3].emit(pb, self.cv1[2].emit(pb, y))
        else:
            y = emit_dw_pw(pb, self.cv1[2], self.cv1[3], y)
        return self.cv1[4].emit(pb, y, dst, res=src)  # C2fCIB always shortcut=True, c_in == c_out


class C2f(nn.Module):
    """C2f / C2fCIB scaffold: one concat buffer, every chunk written at its channel offset."""

    def __init__(self, cin: int, cout: int, n: int, shortcut: bool, cib: bool = False, lk: bool = False):
        super().__init__()
        c = int(cout * 0.5)
        self.c, self.n = c, n
        self.cv1 = ConvBN(cin, 2 * c, 1)
        self.cv2 = ConvBN((2 + n) * c, cout, 1)
        self.m = nn.ModuleList([CIB(c, lk) if cib else Bottleneck(c, shortcut) for _ in range(n)])

    def emit(self, pb, src, dst=None, upcat=None, pre=None):
        """``upcat=(low, skip)``: the block's input is cat[upsample2x(low), skip] (top-down neck);
        cv1 is then applied without materialising either (PlanBuilder.upcat_conv).
        ``pre``: a stride-2 3x3 ``ConvBN`` whose output is this block's input: ``pre`` -> cv1 run as ONE back-to-back GEMM
        launch on ``src`` (csrc/conv_b2b.cu), the tensor between them never exists (backbone.py: cv1 -> c2)."""
        c = self.c
        if upcat is None and pre is None and self._chain_ok(pb, src):
            return self._emit_chain(pb, src, dst)
        H, W = (src.H // 2, src.W // 2) if pre is not None else (src.H, src.W)
        cat = pb.buffer(H, W, (2 + self.n) * c)
        if pre is not None:
            w0, b0 = pre.folded(pb)
            w1, b1 = self.cv1.folded(pb)
            cm = pre.conv.out_channels
            pb.chain(src, [src.c, cm], 1, [dict(k=3, act=True, w=w0, b=b0, src=[(0, 0, src.c)], dst=(1, 0, cm)),
                                           dict(k=1, act=self.cv1.act, w=w1, b=b1, src=[(1, 0, cm)])],
                     dst=cat.view(0, 2 * c), stride0=2)
        elif upcat is not None:
            w, b = self.cv1.folded(pb)
            pb.upcat_conv(upcat[0], upcat[1], w, b, act=self.cv1.act, dst=cat.view(0, 2 * c))
        else:
            self.cv1.emit(pb, src, cat.view(0, 2 * c))
        y = cat.view(c, c)
        for i, m in enumerate(self.m):
            y = m.emit(pb, y, cat.view((2 + i) * c, c))
        return self.cv2.emit(pb, cat.view(), dst)


    def _chain_ok(self, pb, src) -> bool:
        """One Bottleneck, narrow enough for all four weight sets to stay resident in shared memory (56 c^2 bytes):
        the block's thin intermediates (cv1 output, the two 3x3 results, the concat) then never reach HBM."""
        c = self.c
        return (pb.chain_fusable() and self.n == 1 and isinstance(self.m[0], Bottleneck) and c in (16, 32)
                and src.c == self.cv1.conv.in_channels and src.c in (16, 32, 64) and self.cv2.conv.out_channels % 16 == 0)

    def _emit_chain(self, pb, src, dst):
        """layers.py:129-173 as ONE launch.  Regions: 0 = x, 1 = cv1(x) = [y1 | y2], 2 = the Bottleneck's 3x3
        results (the second overwrites the first in place)."""
        c, m = self.c, self.m[0]
        w1, b1 = self.cv1.folded(pb)
        wa, ba = m.cv1.folded(pb)
        wb, bb = m.cv2.folded(pb)
        w2, b2 = self.cv2.folded(pb)
        stages = [
            dict(k=1, act=True, w=w1, b=b1, src=[(0, 0, src.c)], dst=(1, 0, 2 * c)),
            dict(k=3, act=True, w=wa, b=ba, src=[(1, c, c)], dst=(2, 0, c)),
            dict(k=3, act=True, w=wb, b=bb, src=[(2, 0, c)], dst=(2, 0, c), res=(1, c, c) if m.add else None),
            dict(k=1, act=True, w=w2, b=b2, src=[(1, 0, c), (1, c, c), (2, 0, c)]),
        ]
        return pb.chain(src, [src.c, 2 * c, c], 1, stages, dst=dst)


class SPPF(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.c = cin // 2
        self.cv1 = ConvBN(cin, self.c, 1)
        self.cv2 = ConvBN(self.c * 4, cout, 1)

    def emit(self, pb, src, dst=None):
        cat = pb.buffer(src.H, src.W, 4 * self.c)
        self.cv1.emit(pb, src, cat.view(0, self.c))
        pb.sppf_pool(cat, self.c)
        return self.cv2.emit(pb, cat.view(), dst)


class Attention(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.nh = max(1, dim // 64)
        self.hd = dim // self.nh
        self.kd = int(self.hd * 0.5)
        self.qkv = ConvBN(dim, dim + 2 * self.nh * self.kd, 1, act=False)
        self.proj = ConvBN(dim, dim, 1, act=False)
        self.pe = ConvBN(dim, dim, 3, g=dim, act=False)

    def emit(self, pb, b):
        """b <- b + proj(attn(b) + pe(v)), written in place into b's channel slice."""
        nh, kd, hd = self.nh, self.kd, self.hd
        kdp = (kd + 7) // 8 * 8  # q/k rows padded so every head starts 16-byte aligned
        # Re-order the qkv output channels from the reference's per-head interleave
        # [h][q|k|v] (layers.py:373-375) to [all q | all k | all v]: v becomes one
        # contiguous slice (the pe depthwise conv and the residual read it directly).
        per = 2 * kd + hd
        perm: List[int] = []
        for part, width, padw in ((0, kd, kdp), (kd, kd, kdp), (2 * kd, hd, hd)):
            for h in range(nh):
                perm += [h * per + part + j for j in range(width)] + [-1] * (padw - width)
        w, bias = self.qkv.folded(pb)
        qkv = pb.conv(b, w, bias, k=1, stride=1, act=False, out_perm=perm)
        att = pb.attention(qkv, nh=nh, kdp=kdp, hd=hd, scale=float(kd) ** -0.5)
        xa = self.pe.emit(pb, qkv.sub(2 * nh * kdp, nh * hd), res=att)
        return self.proj.emit(pb, xa, dst=b, res=b)


class PSA(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.c = c // 2
        self.cv1 = ConvBN(c, 2 * self.c, 1)
        self.cv2 = ConvBN(2 * self.c, c, 1)
        self.attn = Attention(self.c)
        self.ffn = nn.Sequential(ConvBN(self.c, 2 * self.c, 1), ConvBN(2 * self.c, self.c, 1, act=False))

    def emit(self, pb, src, dst=None):
        ab = self.cv1.emit(pb, src)
        b = ab.sub(self.c, self.c)
        self.attn.emit(pb, b)
        self.ffn[1].emit(pb, self.ffn[0].emit(pb, b), dst=b, res=b)
        return self.cv2.emit(pb, ab, dst)


class SCDown(nn.Module):
    def __init__(self, cin: int, cout: int):
        super().__init__()
        self.cv1 = ConvBN(cin, cout, 1)
        self.cv2 = ConvBN(cout, cout, 3, s=2, g=cout, act=False)

    def emit(self, pb, src, dst=None):
        return self.cv2.emit(pb, self.cv1.emit(pb, src), dst)


class Backbone(nn.Module):
    def __init__(self, v: Variant, in_channels: int = 3):
        super().__init__()
        w, r = v.width, v.reps
        self.cv0 = ConvBN(in_channels, w[0], 3, 2)
        self.cv1 = ConvBN(w[0], w[1], 3, 2)
        self.c2 = C2f(w[1], w[2], r[2], True)
        self.cv3 = ConvBN(w[2], w[3], 3, 2)
        self.c4 = C2f(w[3], w[4], r[4], True)
        self.sc5 = SCDown(w[4], w[5])
        self.c6 = C2f(w[5], w[6], r[6], True, cib="c6" in v.cib)
        self.sc7 = SCDown(w[6], w[7])
        self.c8 = C2f(w[7], w[8], r[8], True, cib="c8" in v.cib, lk="c8" in v.lk)
        self.sppf9 = SPPF(w[8], w[9])
        self.psa10 = PSA(w[10])
        self.out_c = (w[4], w[6], w[10])

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        """``model.backbone(x)`` (backbone.py:88-106): normalised NCHW image -> (C3, C4, C5) NCHW fp32."""
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] % 32 or x.shape[3] % 32:
            raise ValueError("expected input [B,3,H,W] with H and W multiples of 32")
        out = self._root()._run_part("backbone", {}, (x.shape[2], x.shape[3]), x)
        return out[("c3", 0)], out[("c4", 0)], out[("c5", 0)]

    def emit(self, pb, x, c3_dst=None, c4_dst=None, c5_dst=None):
        cv1, c2 = self.cv1, self.c2
        if (pb.stem_pair_fusable() and cv1.k == 3 and cv1.s == 2 and cv1.g == 1 and cv1.act and x.c in (32, 64) and x.H % 2 == 0 and x.W % 2 == 0
                and cv1.conv.out_channels in (32, 64) and c2.cv1.conv.out_channels in (32, 64) and c2.cv1.act and not pb.chain_fusable()):
            y = c2.emit(pb, x, pre=cv1)       # cv1 (3x3 / s2) -> c2.cv1 (1x1) back to back: yolov10s 32 -> 64 -> 64
        else:
            y = c2.emit(pb, cv1.emit(pb, x))
        y = self.cv3.emit(pb, y)
        c3 = self.c4.emit(pb, y, c3_dst)
        c4 = self.c6.emit(pb, self.sc5.emit(pb, c3), c4_dst)
        y = self.c8.emit(pb, self.sc7.emit(pb, c4))
        c5 = self.psa10.emit(pb, self.sppf9.emit(pb, y), c5_dst)
        return c3, c4, c5


class Neck(nn.Module):
    def __init__(self, v: Variant, c3: int, c4: int, c5: int):
        super().__init__()
        h, r = v.neck, v.reps
        self.p5_p4_c2f = C2f(c5 + c4, h[13], r[13], "p5_p4" in v.cib, cib="p5_p4" in v.cib, lk=False)
        self.p4_p3_c2f = C2f(h[13] + c3, h[16], r[16], False)
        self.p3_down = ConvBN(h[16], h[16], 3, 2)
        self.p3_p4_c2f = C2f(h[16] + h[13], h[19], r[19], "p3_p4" in v.cib, cib="p3_p4" in v.cib, lk=False)
        self.p4_down = SCDown(h[19], h[19])
        self.p4_p5_c2f = C2f(h[19] + c5, h[22], r[22], True, cib=True, lk="p4_p5" in v.lk)
        self.widths = (c3, c4, c5, h[13], h[16], h[19], h[22])
        self.out_c = (h[16], h[19], h[22])

    @torch.no_grad()
    def forward(self, c3: torch.Tensor, c4: torch.Tensor, c5: torch.Tensor):
        """``model.neck(c3, c4, c5)`` (neck.py:102-129): NCHW fp32 in, (P3, P4, P5) NCHW fp32 out."""
        out = self._root()._run_part("neck", {"c3": c3, "c4": c4, "c5": c5}, (8 * c3.shape[2], 8 * c3.shape[3]))
        return out[("p3", 0)], out[("p4", 0)], out[("p5", 0)]

    def concat_buffers(self, pb, h8, w8):
        """The four concat inputs of the neck, allocated before the backbone runs so
        that c3/c4/c5 (and later p4', down3, down4) are produced directly in place."""
        c3, c4, c5, h13, h16, h19, _ = self.widths
        return dict(
            cat_p4=pb.buffer(h8 // 2, w8 // 2, c5 + c4),    # [up(c5) | c4]        neck.py:117
            cat_p3=pb.buffer(h8, w8, h13 + c3),             # [up(p4') | c3]       neck.py:120
            cat_n4=pb.buffer(h8 // 2, w8 // 2, h16 + h13),  # [down(p3) | p4']     neck.py:124
            cat_n5=pb.buffer(h8 // 4, w8 // 4, h19 + c5),   # [down(p4) | c5]      neck.py:127
        )

    def emit(self, pb, cats):
        c3, c4, c5, h13, h16, h19, _ = self.widths
        c5v = cats["cat_n5"].view(h19, c5)
        if pb.upcat_fusable():
            # top-down path: upsample + concat + 1x1 folded into the 1x1 (conv commutes with nearest upsampling)
            p4a = self.p5_p4_c2f.emit(pb, cats["cat_p4"].view(), cats["cat_n4"].view(h16, h13),
                                      upcat=(c5v, cats["cat_p4"].view(c5, c4)))
            p3 = self.p4_p3_c2f.emit(pb, cats["cat_p3"].view(), upcat=(p4a, cats["cat_p3"].view(h13, c3)))
        else:
            pb.upsample2x(c5v, cats["cat_p4"].view(0, c5))
            p4a = self.p5_p4_c2f.emit(pb, cats["cat_p4"].view(), cats["cat_n4"].view(h16, h13))
            pb.upsample2x(p4a, cats["cat_p3"].view(0, h13))
            p3 = self.p4_p3_c2f.emit(pb, cats["cat_p3"].view())
        self.p3_down.emit(pb, p3, cats["cat_n4"].view(0, h16))
        p4 = self.p3_p4_c2f.emit(pb, cats["cat_n4"].view())
        self.p4_down.emit(pb, p4, cats["cat_n5"].view(0, h19))
        p5 = self.p4_p5_c2f.emit(pb, cats["cat_n5"].view())
        return p3, p4, p5


class DFL(nn.Module):
    """Only the ``bins`` buffer matters (it is a state_dict key); head.py:32-49."""

    def __init__(self, c1: int):
        super().__init__()
        self.c1 = int(c1)
        self.register_buffer("bins", torch.arange(self.c1, dtype=torch.float).view(1, 1, self.c1, 1))


class Detect(nn.Module):
    """v10Detect parameter holder: per-level reg (cv2) / cls (cv3) stacks, twice."""

    def __init__(self, nc: int, ch: Sequence[int], reg_max: int = REG_MAX):
        super().__init__()
        self.nc, self.nl, self.reg_max = nc, len(ch), reg_max
        self.no = nc + 4 * reg_max
        self.stride = torch.zeros(self.nl)  # dead field kept for parity (head.py:81)
        c2 = max(16, ch[0] // 4, 4 * reg_max)
        c3 = max(ch[0], min(nc, 100))
        self.cv2 = nn.ModuleList(
            nn.Sequential(ConvBN(x, c2, 3), ConvBN(c2, c2, 3), nn.Conv2d(c2, 4 * reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(
            nn.Sequential(
                nn.Sequential(ConvBN(x, x, 3, g=x), ConvBN(x, c3, 1)),
                nn.Sequential(ConvBN(c3, c3, 3, g=c3), ConvBN(c3, c3, 1)),
                nn.Conv2d(c3, nc, 1),
            ) for x in ch)
        self.one2one_cv2 = copy.deepcopy(self.cv2)
        self.one2one_cv3 = copy.deepcopy(self.cv3)
        self.dfl = DFL(reg_max) if reg_max > 1 else nn.Identity()

    def _run(self, x: Sequence[torch.Tensor]):
        assert len(x) == self.nl
        return self._root()._run_part("head", {f"p{3 + i}": t for i, t in enumerate(x)}, (8 * x[0].shape[2], 8 * x[0].shape[3]))

    @torch.no_grad()
    def forward_feat(self, x: Sequence[torch.Tensor], cv2, cv3) -> List[torch.Tensor]:
        """head.py:118-122.  ``cv2`` / ``cv3`` select the branch: the module lists of this head (the
        one-to-many stacks or their one-to-one copies), as the reference's callers pass them."""
        if cv2 is self.cv2 and cv3 is self.cv3:
            name = "one2many"
        elif cv2 is self.one2one_cv2 and cv3 is self.one2one_cv3:
            name = "one2one"
        else:
            raise ValueError("forward_feat: cv2/cv3 must be this head's (cv2, cv3) or (one2one_cv2, one2one_cv3)")
        out = self._run(x)
        return [out[(name, i)] for i in range(self.nl)]

    @torch.no_grad()
    def forward(self, x: Sequence[torch.Tensor]):
        """Eval: the one2many list (head.py:124-135).  Training is out of scope (raises in the root model)."""
        out = self._run(x)
        return [out[("one2many", i)] for i in range(self.nl)]

    def emit(self, pb, feats, only: Optional[str] = None):
        """Both branches (head.py:118-135), or just ``only`` ("one2one": what ``decode_forward`` consumes).  Each writes [reg(4*reg_max) | cls(nc)] logits of every
        level straight into its public NCHW fp32 tensor from the two final 1x1 epilogues.  The
        first reg conv of the two branches reads the same feature map, so the pair is ONE
        implicit GEMM with the output channels concatenated (N = 2*c2: one pass over the
        input, twice the MMA width); each branch then continues from its channel slice."""
        branches = (("one2many", self.cv2, self.cv3), ("one2one", self.one2one_cv2, self.one2one_cv3))
        if only is not None:
            branches = tuple(b for b in branches if b[0] == only)
        for i, f in enumerate(feats):
            folded = [reg[i][0].folded(pb) for _, reg, _ in branches]
            c2 = folded[0][0].shape[0]
            if c2 % 16 == 0 and len(branches) > 1:
                r01 = pb.conv(f, pb.cat0([w for w, _ in folded]), pb.cat0([b for _, b in folded]), k=3, stride=1, act=True)
                firsts = [r01.sub(j * c2, c2) for j in range(len(branches))]
            else:
                firsts = [reg[i][0].emit(pb, f) for _, reg, _ in branches]
            for (out_name, reg, cls), r in zip(branches, firsts):
                fin = reg[i][2]
                wf, bf = pb.param(fin.weight), pb.param(fin.bias)
                if ((pb.tail_fusable() and r.c in (32, 64)) or (pb.chain_fusable() and r.c == 16)) and r.c == c2 and \
                        (4 * self.reg_max) % 16 == 0 and 4 * self.reg_max <= 64:
                    # 3x3 -> 1x1 tail of the regression stack as one launch (the 3x3 result stays in shared memory)
                    wm, bm = reg[i][1].folded(pb)
                    pb.chain(r, [c2, c2], 1, [dict(k=3, act=True, w=wm, b=bm, src=[(0, 0, c2)], dst=(1, 0, c2)),
                                              dict(k=1, act=False, w=wf, b=bf, src=[(1, 0, c2)])],
                             nchw=(out_name, i, 0, 4 * self.reg_max, self.no))
                else:
                    r = reg[i][1].emit(pb, r)
                    pb.conv(r, wf, bf, k=1, stride=1, act=False, nchw=(out_name, i, 0, 4 * self.reg_max, self.no))
                c = emit_dw_pw(pb, cls[i][0][0], cls[i][0][1], f)
                c = emit_dw_pw(pb, cls[i][1][0], cls[i][1][1], c)
                fin = cls[i][2]
                pb.conv(c, pb.param(fin.weight), pb.param(fin.bias), k=1, stride=1,
                        act=False, nchw=(out_name, i, 4 * self.reg_max, self.nc, self.no))
