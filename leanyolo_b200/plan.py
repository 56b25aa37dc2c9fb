"""Lowering target: a flat list of kernel ops over NHWC buffers in one workspace.

``PlanBuilder`` is filled by the ``emit`` methods in ``modules.py``.  It owns

* the buffer table (NHWC, channel count padded to a multiple of 16, byte offsets
  into a single workspace allocation),
* the packed parameter blob (BN-folded, re-parameterised, re-ordered weights in
  the storage dtype; biases always fp32), and
* the op list (``Op``) that ``engine.py`` serialises into ``ly_op`` structs for
  the C-ABI (include/leanyolo_b200.h).

Data layout in HBM: activations ``[B, H, W, Ctot]`` (channels innermost) in bf16
(or fp32 in check mode); a concat is one buffer and every producer writes its
channel slice; a split is a channel-offset view.  Dense weights are
``[Cout_pad][kh][kw][Cin_pad]`` (K-major rows: the B operand of the implicit
GEMM); depthwise weights are ``[kh*kw][C_pad]``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

CH_ALIGN = 16
BUF_ALIGN = 1024


def _rup(x: int, a: int) -> int:
    return (x + a - 1) // a * a


class ShapeOnly:
    """Stand-in for a weight tensor in a DRY lowering (``PlanBuilder(dry=True)``): it carries the shape through
    the few manipulations the emitters apply (slicing, clone, add) and ignores the arithmetic.  (The meta device
    would do, but its first in-place op or cat imports seconds of symbolic-shape machinery.)"""
    __slots__ = ("shape",)

    def __init__(self, shape):
        self.shape = tuple(int(v) for v in shape)

    def numel(self) -> int:
        n = 1
        for v in self.shape:
            n *= v
        return n

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        shape = []
        for d, n in enumerate(self.shape):
            if d < len(idx):
                i = idx[d]
                if isinstance(i, slice):
                    shape.append(len(range(*i.indices(n))))
                # an int index drops the dimension
            else:
                shape.append(n)
        return ShapeOnly(shape)

    def __setitem__(self, idx, value):
        pass

    def clone(self):
        return ShapeOnly(self.shape)

    def __add__(self, other):
        return self

    __radd__ = __iadd__ = __add__


@dataclass
class Buf:
    id: int
    H: int
    W: int
    C: int          # padded channel count == pixel pitch in elements
    offset: int     # bytes into the workspace

    def view(self, c0: int = 0, c: Optional[int] = None) -> "View":
        c = self.C - c0 if c is None else _rup(c, CH_ALIGN)
        assert c0 % 8 == 0 and c0 + c <= self.C, (c0, c, self.C)
        return View(self, c0, c)


@dataclass
class View:
    buf: Buf
    c0: int
    c: int

    @property
    def H(self):
        return self.buf.H

    @property
    def W(self):
        return self.buf.W

    def sub(self, c0: int, c: int) -> "View":
        assert c0 + c <= self.c and c0 % 8 == 0
        return View(self.buf, self.c0 + c0, c)


@dataclass
class Op:
    kind: str                       # stem | conv | dw | dwpw | pool | up | attn | export
    src: Optional[View] = None
    dst: Optional[View] = None
    res: Optional[View] = None
    w_off: int = -1                 # element offset into the weight blob (storage dtype)
    b_off: int = -1                 # element offset into the fp32 bias blob
    k: int = 1
    stride: int = 1
    act: bool = False
    cin: int = 0                    # real (unpadded) channel counts, for FLOP accounting
    cout: int = 0
    nchw: Optional[Tuple[str, int, int, int, int]] = None   # (tensor name, level, c0, c, ctot)
    attn: Optional[Tuple[int, int, int, float]] = None      # (nh, kdp, hd, scale)
    extra: dict = field(default_factory=dict)


class PlanBuilder:
    def __init__(self, B: int, H: int, W: int, dtype: str = "bf16", tensor_core: bool = True, dry: bool = False):
        """``dry``: lower the op list and the parameter OFFSETS only; weights are shape-only meta tensors (see
        ``ConvBN.folded``), nothing is folded, copied or stored.  Used when the packed blobs already exist."""
        assert dtype in ("bf16", "f32")
        self.dry = dry
        # False: bring-up / bisecting mode (every conv on the CUDA-core kernel): no tensor-core-only fusions
        self.tensor_core = tensor_core and dtype == "bf16"
        assert H % 32 == 0 and W % 32 == 0, "input H, W must be multiples of 32 (neck concat, SURVEY App. C.10)"
        self.B, self.H, self.W, self.dtype = B, H, W, dtype
        self.esize = 2 if dtype == "bf16" else 4
        self.bufs: List[Buf] = []
        self.ops: List[Op] = []
        self.ws_bytes = 0
        self._w: List[torch.Tensor] = []   # fp64 chunks, cast at finalize
        self._w_len = 0
        self._b: List[torch.Tensor] = []
        self._b_len = 0
        self.outputs: Dict[Tuple[str, int], Tuple[int, int, int]] = {}   # (name, level) -> (C, H, W)
        self.inputs: Dict[Tuple[str, int], Tuple[int, int, int]] = {}    # external NCHW fp32 inputs besides the image

    # ---------------------------------------------------------------- buffers
    def buffer(self, H: int, W: int, C: int) -> Buf:
        C = _rup(C, CH_ALIGN)
        b = Buf(len(self.bufs), H, W, C, self.ws_bytes)
        self.ws_bytes += _rup(self.B * H * W * C * self.esize, BUF_ALIGN)
        self.bufs.append(b)
        return b

    def assign_offsets(self, reuse: bool = True) -> int:
        """Place the buffers in the workspace.  The op list is linear and kernels of one forward run in stream
        order, so a buffer only needs memory from the first op that touches it to the last one: buffers with
        disjoint lifetimes share bytes (first-fit over a free list of address ranges, processed in order of first
        use).  Without reuse yolov10s needs 74.8 MB per image at 640x640 and yolov10x 317 MB (19 GB / 81 GB at batch
        256); with it roughly a third of that.  Returns the workspace size in bytes."""
        first, last = {}, {}
        for i, op in enumerate(self.ops):
            for v in (op.src, op.dst, op.res, op.extra.get("up")):
                if v is not None:
                    first.setdefault(v.buf.id, i)
                    last[v.buf.id] = i
        size = {b.id: _rup(self.B * b.H * b.W * b.C * self.esize, BUF_ALIGN) for b in self.bufs}
        if not reuse:
            off = 0
            for b in self.bufs:
                b.offset = off
                off += size[b.id]
            self.ws_bytes = off
            return off
        order = sorted((b for b in self.bufs if b.id in first), key=lambda b: (first[b.id], b.id))
        free: List[Tuple[int, int]] = []          # (offset, bytes), sorted by offset, coalesced
        live: List[Tuple[int, int, int]] = []     # (last use, offset, bytes)
        top = 0

        def release(off: int, n: int) -> None:
            free.append((off, n))
            free.sort()
            merged: List[Tuple[int, int]] = []
            for o, m in free:
                if merged and merged[-1][0] + merged[-1][1] == o:
                    merged[-1] = (merged[-1][0], merged[-1][1] + m)
                else:
                    merged.append((o, m))
            free[:] = merged

        for b in order:
            t = first[b.id]
            for item in [x for x in live if x[0] < t]:
                live.remove(item)
                release(item[1], item[2])
            need = size[b.id]
            slot = next((i for i, (o, m) in enumerate(free) if m >= need), None)
            if slot is None:
                if free and free[-1][0] + free[-1][1] == top:      # grow the free tail instead of leaving a hole
                    o, m = free.pop()
                    b.offset = o
                    top = o + need
                else:
                    b.offset = top
                    top += need
            else:
                o, m = free.pop(slot)
                b.offset = o
                if m > need:
                    release(o + need, m - need)
            live.append((last[b.id], b.offset, need))
        for b in self.bufs:                       # never referenced by an op: park at 0 (no bytes needed)
            if b.id not in first:
                b.offset = 0
        self.ws_bytes = max(top, BUF_ALIGN)
        return self.ws_bytes

    # ---------------------------------------------------------------- params
    def param(self, t: torch.Tensor) -> torch.Tensor:
        """A raw (un-foldable) parameter as fp64 on the CPU; shape only when dry."""
        if self.dry:
            return ShapeOnly(t.shape)
        return t.detach().double().cpu()

    def cat0(self, ts: Sequence[torch.Tensor]) -> torch.Tensor:
        """torch.cat along dim 0 (shape only when dry: the first torch.cat on the meta device costs seconds of lazy imports)."""
        if self.dry:
            return ShapeOnly((sum(t.shape[0] for t in ts),) + tuple(ts[0].shape[1:]))
        return torch.cat(list(ts))

    def _add_w(self, w) -> int:
        """w: tensor, or (dry) the element count."""
        off = self._w_len
        n = w if isinstance(w, int) else w.numel()
        pad = (-n) % 64          # keep every weight block 128-byte aligned (TMA base address)
        if not self.dry:
            self._w.append(torch.cat([w.reshape(-1).double(), torch.zeros(pad, dtype=torch.float64)]))
        self._w_len += n + pad
        return off

    def _add_b(self, b) -> int:
        off = self._b_len
        n = b if isinstance(b, int) else b.numel()
        pad = (-n) % 4
        if not self.dry:
            self._b.append(torch.cat([b.reshape(-1).double(), torch.zeros(pad, dtype=torch.float64)]))
        self._b_len += n + pad
        return off

    def signature(self) -> str:
        """Identifies the parameter packing (op kinds, shapes, blob offsets): a cached blob is only valid for the
        lowering that produced it (fusion switches and package versions change the layout)."""
        import hashlib
        h = hashlib.sha256()
        h.update(f"{self.dtype}|{self.tensor_core}|{self._w_len}|{self._b_len}".encode())
        for op in self.ops:
            h.update(f"{op.kind},{op.w_off},{op.b_off},{op.k},{op.stride},{op.cin},{op.cout},{int(op.act)}".encode())
            for key in ("pre_w_off", "pre_b_off"):
                if key in op.extra:
                    h.update(f",{op.extra[key]}".encode())
            for st in op.extra.get("stages", ()):
                h.update(f",{st['w_off']},{st['b_off']},{st['k']},{st['cout']},{st['cin']}".encode())
        return h.hexdigest()

    def finalize_params(self) -> Tuple[torch.Tensor, torch.Tensor]:
        wdt = torch.bfloat16 if self.dtype == "bf16" else torch.float32
        w = torch.cat(self._w).to(torch.float32).to(wdt) if self._w else torch.zeros(0, dtype=wdt)
        b = torch.cat(self._b).to(torch.float32) if self._b else torch.zeros(0)
        return w, b

    # ---------------------------------------------------------------- ops
    def stem(self, w: torch.Tensor, b: torch.Tensor, sub: Sequence[float], div: Sequence[float]) -> View:
        """3x3 stride-2 conv on the user's NCHW fp32 image, normalisation applied in the
        loader (zero padding happens AFTER normalisation in the reference, yolov10s.py:107-113,
        so a non-zero ``subtract`` cannot be folded into a bias)."""
        cout, cin = w.shape[0], w.shape[1]
        assert cin == 3 and w.shape[2:] == (3, 3)
        dst = self.buffer(self.H // 2, self.W // 2, cout).view()
        if self.dry:
            wp, bp = dst.c * 27, dst.c
        else:
            wp = torch.zeros(dst.c, 27, dtype=torch.float64)
            wp[:cout] = w.permute(0, 2, 3, 1).reshape(cout, 27)      # [co][ky][kx][ci]
            bp = torch.zeros(dst.c, dtype=torch.float64)
            bp[:cout] = b
        op = Op("stem", dst=dst, k=3, stride=2, act=True, cin=3, cout=cout,
                extra=dict(sub=[float(v) for v in sub], div=[float(v) for v in div]))
        # stem weights are consumed by CUDA cores in fp32 regardless of the storage dtype
        op.w_off, op.b_off = self._add_b(wp), self._add_b(bp)
        self.ops.append(op)
        return dst

    def conv(self, src: View, w: torch.Tensor, b: torch.Tensor, *, k: int, stride: int, act: bool,
             dst: Optional[View] = None, res: Optional[View] = None, out_perm: Optional[List[int]] = None,
             nchw: Optional[Tuple[str, int, int, int, int]] = None, up: Optional[View] = None,
             bias: bool = True) -> Optional[View]:
        cout, cin = w.shape[0], w.shape[1]
        assert w.shape[2] == w.shape[3] == k and k in (1, 3) and stride in (1, 2)
        assert cin <= src.c < cin + CH_ALIGN, (cin, src.c)
        if out_perm is not None:
            if self.dry:
                w, b = ShapeOnly((len(out_perm),) + tuple(w.shape[1:])), ShapeOnly((len(out_perm),))
            else:
                wn = torch.zeros(len(out_perm), *w.shape[1:], dtype=torch.float64)
                bn = torch.zeros(len(out_perm), dtype=torch.float64)
                idx = torch.tensor(out_perm)
                m = idx >= 0
                wn[m], bn[m] = w[idx[m]], b[idx[m]]
                w, b = wn, bn
            cout = len(out_perm)
        cpad = _rup(cout, CH_ALIGN)
        Ho, Wo = src.H // stride, src.W // stride
        if nchw is None:
            if dst is None:
                dst = self.buffer(Ho, Wo, cout).view()
            assert dst.c == cpad and (dst.H, dst.W) == (Ho, Wo), (dst.c, cpad, dst.H, Ho)
        else:
            name, level, c0, c, ctot = nchw
            self.outputs[(name, level)] = (ctot, Ho, Wo)
        if self.dry:
            wp, bp = cpad * k * k * src.c, cpad
        else:
            wp = torch.zeros(cpad, k, k, src.c, dtype=torch.float64)
            wp[:cout, :, :, :cin] = w.permute(0, 2, 3, 1)
            bp = torch.zeros(cpad, dtype=torch.float64)
            bp[:cout] = b
        if res is not None:
            assert res.c == cpad and (res.H, res.W) == (Ho, Wo)
        extra = dict(cpad=cpad)
        if up is not None:
            # half-resolution tensor added before the activation at (h/2, w/2): see upcat_conv()
            assert self.dtype == "bf16" and up.c == cpad and (2 * up.H, 2 * up.W) == (Ho, Wo)
            extra["up"] = up
        self.ops.append(Op("conv", src=src, dst=dst, res=res, w_off=self._add_w(wp), b_off=self._add_b(bp),
                           k=k, stride=stride, act=act, cin=cin, cout=cout, nchw=nchw, extra=extra))
        return dst

    def upcat_fusable(self) -> bool:
        import os
        return self.tensor_core and os.environ.get("LEANYOLO_FUSE_UPCAT", "1") != "0"

    def upcat_conv(self, low: View, skip: View, w: torch.Tensor, b: torch.Tensor, *, act: bool, dst: View) -> View:
        """act(conv1x1(cat[upsample2x(low), skip])) without the upsample or the concat: a 1x1 conv
        commutes with nearest upsampling, so the ``low`` half of the weights is applied at half
        resolution (4x fewer pixels) and added, upsampled on the fly, before the activation of the
        conv over ``skip`` (neck.py:116-121; cat order [up, skip])."""
        c_low = w.shape[1] - skip.c
        assert c_low == low.c and w.shape[2:] == (1, 1) and (2 * low.H, 2 * low.W) == (skip.H, skip.W)
        t = self.conv(low, w[:, :c_low], b if self.dry else torch.zeros_like(b), k=1, stride=1, act=False)
        return self.conv(skip, w[:, c_low:], b, k=1, stride=1, act=act, dst=dst, up=t)

    def dwconv(self, src: View, w: torch.Tensor, b: torch.Tensor, *, k: int, stride: int, act: bool,
               dst: Optional[View] = None, res: Optional[View] = None) -> View:
        c = w.shape[0]
        assert w.shape[1] == 1 and w.shape[2] == w.shape[3] == k and k in (3, 7)
        assert c <= src.c < c + CH_ALIGN
        Ho, Wo = (src.H + stride - 1) // stride, (src.W + stride - 1) // stride
        if dst is None:
            dst = self.buffer(Ho, Wo, c).view()
        assert dst.c == src.c and (dst.H, dst.W) == (Ho, Wo)
        if self.dry:
            wp, bp = k * k * src.c, src.c
        else:
            wp = torch.zeros(k * k, src.c, dtype=torch.float64)
            wp[:, :c] = w.reshape(c, k * k).t()
            bp = torch.zeros(src.c, dtype=torch.float64)
            bp[:c] = b
        if res is not None:
            assert res.c == dst.c and (res.H, res.W) == (Ho, Wo)
        self.ops.append(Op("dw", src=src, dst=dst, res=res, w_off=self._add_w(wp), b_off=self._add_b(bp),
                           k=k, stride=stride, act=act, cin=c, cout=c))
        return dst

    def dwpw_fusable(self, src: View, c: int, cout: int, k: int, stride: int) -> bool:
        """dw k x k -> 1x1 pairs the fused tensor-core kernel takes (csrc/dwpw_tc.cu)."""
        import os
        if os.environ.get("LEANYOLO_FUSE_DWPW", "1") == "0":
            return False
        return (self.tensor_core and k == 3 and stride == 1 and src.c % 64 == 0 and src.c == c
                and _rup(cout, CH_ALIGN) <= 256)

    def dwpw(self, src: View, dw_w: torch.Tensor, dw_b: torch.Tensor, pw_w: torch.Tensor, pw_b: torch.Tensor, *,
             dw_act: bool, act: bool, dst: Optional[View] = None,
             nchw: Optional[Tuple[str, int, int, int, int]] = None) -> Optional[View]:
        """act2(pw(act1(dw3x3(src)))) as ONE launch: the depthwise result feeds the 1x1 GEMM through
        shared memory and never reaches HBM."""
        c = dw_w.shape[0]
        cout, cin = pw_w.shape[0], pw_w.shape[1]
        assert dw_w.shape[1:] == (1, 3, 3) and pw_w.shape[2:] == (1, 1) and cin == c == src.c
        cpad = _rup(cout, CH_ALIGN)
        if nchw is None:
            if dst is None:
                dst = self.buffer(src.H, src.W, cout).view()
            assert dst.c == cpad and (dst.H, dst.W) == (src.H, src.W)
        else:
            name, level, c0, cc, ctot = nchw
            self.outputs[(name, level)] = (ctot, src.H, src.W)
        if self.dry:
            dwp, wp, bp, dw_b = 9 * c, cpad * src.c, cpad, c
        else:
            dwp = dw_w.reshape(c, 9).t().contiguous()                      # [9][C]
            wp = torch.zeros(cpad, 1, 1, src.c, dtype=torch.float64)
            wp[:cout, :, :, :cin] = pw_w.permute(0, 2, 3, 1)
            bp = torch.zeros(cpad, dtype=torch.float64)
            bp[:cout] = pw_b
        self.ops.append(Op("dwpw", src=src, dst=dst, w_off=self._add_w(wp), b_off=self._add_b(bp), k=1, stride=1, act=act,
                           cin=cin, cout=cout, nchw=nchw,
                           extra=dict(cpad=cpad, pre_w_off=self._add_w(dwp), pre_b_off=self._add_b(dw_b), pre_act=dw_act)))
        return dst

    def chain_fusable(self) -> bool:
        """LY_OP_CHAIN (csrc/chain_tc.cu) is a bf16 tensor-core kernel.  Opt-in (LEANYOLO_FUSE_CHAIN=1): measured on
        yolov10s at batch 256 the fused C2f block runs at 2.35 ms against 1.40 ms layer by layer (its N = 32 MMAs
        issue at ~67 cycles each and the per-stage epilogues of a tile serialise), see DESIGN.md."""
        import os
        return self.tensor_core and os.environ.get("LEANYOLO_FUSE_CHAIN", "0") == "1"

    def tail_fusable(self) -> bool:
        """The 3x3 -> 1x1 tail of a regression stack as ONE back-to-back GEMM launch (csrc/conv_b2b.cu, reached through
        LY_OP_CHAIN).  On by default on the bf16 tensor-core path; LEANYOLO_FUSE_TAIL=0 lowers the two convs separately."""
        import os
        return self.tensor_core and os.environ.get("LEANYOLO_FUSE_TAIL", "1") != "0"

    def stem_pair_fusable(self) -> bool:
        """Backbone cv1 (3x3 / s2) -> c2.cv1 (1x1) as one stride-2 back-to-back launch.  Opt-in (LEANYOLO_FUSE_S2=1): bit-correct, but on
        yolov10s at batch 256 the fused launch takes 0.80 ms against 0.41 + 0.33 ms for the two layers: its band of one output row
        (six parity-plane rows of 64-byte pixels, 12 TMA boxes) keeps only ~64 KB per SM in flight, see DESIGN.md."""
        import os
        return self.tail_fusable() and os.environ.get("LEANYOLO_FUSE_S2", "0") == "1"

    def chain(self, src: View, regions: Sequence[int], n_in: int, stages: Sequence[dict], *, dst: Optional[View] = None,
              nchw: Optional[Tuple[str, int, int, int, int]] = None, stride0: int = 1) -> Optional[View]:
        """A chain of dense conv stages executed per spatial tile with every intermediate in shared memory
        (``LY_OP_CHAIN``, include/leanyolo_b200.h).  ``regions``: channels of each shared-memory region (the first
        ``n_in`` are the 64-channel blocks of ``src``); ``stages``: dicts {k, act, w [cout,cin,k,k], b [cout],
        src [(region, c0, c)], dst (region, c0, c) | None, res (region, c0, c) | None}; the last stage writes
        ``dst`` (NHWC slice) or the public NCHW tensor ``nchw``."""
        assert self.dtype == "bf16" and sum(regions[:n_in]) == src.c and 1 <= len(stages) <= 6 and len(regions) <= 8
        packed = []
        for st in stages:
            w, b, k = st["w"], st["b"], st["k"]
            cout, cin = w.shape[0], w.shape[1]
            assert cout % CH_ALIGN == 0 and cin == sum(c for _, _, c in st["src"]) and w.shape[2] == w.shape[3] == k
            packed.append(dict(k=k, act=bool(st["act"]), cout=cout, cin=cin, src=list(st["src"]), dst=st.get("dst"), res=st.get("res"),
                               w_off=self._add_w(w.numel() if self.dry else w.permute(0, 2, 3, 1).contiguous()),
                               b_off=self._add_b(b.numel() if self.dry else b)))
        cout = packed[-1]["cout"]
        assert stride0 in (1, 2) and (stride0 == 1 or (stages[0]["k"] == 3 and src.H % 2 == 0 and src.W % 2 == 0))
        Ho, Wo = src.H // stride0, src.W // stride0       # stride0 = 2: the FIRST stage is a stride-2 3x3 (back-to-back kernel only)
        if nchw is None:
            if dst is None:
                dst = self.buffer(Ho, Wo, cout).view()
            assert dst.c == cout and (dst.H, dst.W) == (Ho, Wo)
        else:
            name, level, c0, c, ctot = nchw
            self.outputs[(name, level)] = (ctot, Ho, Wo)
        self.ops.append(Op("chain", src=src, dst=dst, k=1, stride=1, cin=src.c, cout=cout, nchw=nchw,
                           extra=dict(regions=list(regions), n_in=n_in, stages=packed, cpad=cout, stride0=stride0)))
        return dst

    def sppf_pool(self, cat: Buf, c: int) -> None:
        """cat[..., c:4c] <- three chained 5x5/s1/p2 max-pools of cat[..., 0:c] (windows 5, 9, 13)."""
        assert cat.C == 4 * c and c % 8 == 0
        self.ops.append(Op("pool", src=cat.view(0, c), dst=cat.view(c, 3 * c), cin=c, cout=3 * c))

    def upsample2x(self, src: View, dst: View) -> None:
        assert dst.c == src.c and (dst.H, dst.W) == (2 * src.H, 2 * src.W)
        self.ops.append(Op("up", src=src, dst=dst, cin=src.c, cout=src.c))

    def attention(self, qkv: View, *, nh: int, kdp: int, hd: int, scale: float) -> View:
        assert qkv.c == 2 * nh * kdp + nh * hd
        dst = self.buffer(qkv.H, qkv.W, nh * hd).view()
        self.ops.append(Op("attn", src=qkv, dst=dst, attn=(nh, kdp, hd, scale), cin=qkv.c, cout=nh * hd))
        return dst

    def import_nchw(self, name: str, c: int, H: int, W: int, dst: Optional[View] = None) -> View:
        """Sub-module entry (``model.neck(c3, c4, c5)``, ``model.head(feats)``): a caller's NCHW fp32
        feature map -> NHWC storage (optionally straight into a concat slice)."""
        if dst is None:
            dst = self.buffer(H, W, c).view()
        assert (dst.H, dst.W) == (H, W) and dst.c == _rup(c, CH_ALIGN)
        self.inputs[(name, 0)] = (c, H, W)
        self.ops.append(Op("import", dst=dst, nchw=(name, 0, 0, c, c), cin=c, cout=c))
        return dst

    def export_nchw(self, src: View, name: str, c: int) -> None:
        """Debug / sub-module taps: NHWC storage -> public NCHW fp32 tensor."""
        self.outputs[(name, 0)] = (c, src.H, src.W)
        self.ops.append(Op("export", src=src, nchw=(name, 0, 0, c, c), cin=c, cout=c))

    # ---------------------------------------------------------------- accounting
    def dense_flops(self) -> int:
        """2*MAC of the dense convs per image (the tensor-pipe-eligible work)."""
        f = 0
        for op in self.ops:
            if op.kind in ("conv", "dwpw"):
                Ho, Wo = op.src.H // op.stride, op.src.W // op.stride
                f += 2 * Ho * Wo * op.cout * op.cin * op.k * op.k
            elif op.kind == "chain":
                s0 = op.extra.get("stride0", 1)
                f += sum(2 * (op.src.H // s0) * (op.src.W // s0) * st["cout"] * st["cin"] * st["k"] ** 2 for st in op.extra["stages"])
            elif op.kind == "stem":
                f += 2 * op.dst.H * op.dst.W * op.cout * 27
        return f
