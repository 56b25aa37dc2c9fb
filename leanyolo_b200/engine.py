"""Executes a lowered plan on the GPU through the C ABI.

Host-side runtime of the hot path: owns the device copy of the packed parameters,
one workspace per (batch, H, W) shape, the native ``ly_plan`` (tensor maps encoded
once) and optional CUDA-graph capture of the whole forward.  PyTorch is used only
for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _native as N
from .plan import Op, PlanBuilder, View

_KIND = {"stem": N.OP_STEM, "conv": N.OP_CONV, "dw": N.OP_DW, "dwpw": N.OP_DWPW, "pool": N.OP_POOL, "up": N.OP_UP,
         "attn": N.OP_ATTN, "export": N.OP_EXPORT, "import": N.OP_IMPORT, "chain": N.OP_CHAIN}


def _view(v: Optional[View], base: int) -> N.LyView:
    if v is None:
        return N.LyView(None, 0, 0, 0, 0, 0)
    return N.LyView(base + v.buf.offset, v.H, v.W, v.buf.C, v.c0, v.c)


@dataclass
class Compiled:
    pb: PlanBuilder
    B: int
    workspace: torch.Tensor
    handle: C.c_void_p
    out_keys: List[Tuple[str, int]]
    n_launches: int
    in_keys: List[Tuple[str, int]] = None


def serialise_ops(pb: PlanBuilder, B: int, dtype: str, impl: int, base: int, wbase: int, bbase: int, in_u8=False):
    """The C-ABI description (``ly_op[]``) of a lowered plan: views become (buffer base + offset, H, W, pitch, c0, c),
    parameters become addresses inside the packed weight / bias blobs.  External pointer table of a run:
    [image | named NCHW inputs ... | NCHW outputs ...].  Returns (array, chain structs that must outlive
    ``ly_plan_create``, input keys, output keys).  Pure host code: the CPU tests feed it to ``ly_op_validate``."""
    in_keys = list(pb.inputs.keys())
    out_keys = sorted(pb.outputs.keys())
    in_slots = {k: i + 1 for i, k in enumerate(in_keys)}
    slots = {k: i + 1 + len(in_keys) for i, k in enumerate(out_keys)}
    arr = (N.LyOp * len(pb.ops))()
    esz = pb.esize
    dt = N.LY_BF16 if dtype == "bf16" else N.LY_F32
    chains = []     # ctypes ly_chain structs: must outlive ly_plan_create (which copies them)
    for i, op in enumerate(pb.ops):
        o = arr[i]
        o.kind, o.dtype, o.B = _KIND[op.kind], dt, B
        o.k, o.stride, o.act, o.impl = op.k, op.stride, int(op.act), impl
        o.src, o.dst, o.res = _view(op.src, base), _view(op.dst, base), _view(op.res, base)
        o.ext_slot = -1
        if op.kind == "stem":
            o.w, o.bias = bbase + 4 * op.w_off, bbase + 4 * op.b_off
            for j in range(3):
                o.sub[j], o.div[j] = op.extra["sub"][j], op.extra["div"][j]
            o.ext_slot = 0
            if in_u8 == "lb":     # fused letterbox: ext[0] is the device array of ly_lb_desc, border colour in nh/kdp/hd
                o.impl = N.STEM_IN_LB
                o.nh, o.kdp, o.hd = 114, 114, 114
            elif in_u8:
                o.impl = N.STEM_IN_U8
        elif op.w_off >= 0:
            o.w, o.bias = wbase + esz * op.w_off, bbase + 4 * op.b_off
        if op.extra.get("up") is not None:
            o.up = _view(op.extra["up"], base)
        if op.kind == "dwpw":
            o.pre_w, o.pre_bias = wbase + esz * op.extra["pre_w_off"], bbase + 4 * op.extra["pre_b_off"]
            o.pre_k, o.pre_act = 3, int(op.extra["pre_act"])
        if op.attn is not None:
            o.nh, o.kdp, o.hd, o.scale = op.attn
        if op.nchw is not None:
            name, level, c0, c, ctot = op.nchw
            o.nchw_ctot, o.nchw_c0, o.nchw_c = ctot, c0, c
            o.ext_slot = in_slots[(name, level)] if op.kind == "import" else slots[(name, level)]
        if op.kind == "chain":
            chains.append(Engine._chain_struct(op, wbase, bbase, esz))
            o.chain = C.pointer(chains[-1])
    return arr, chains, in_keys, out_keys


class Engine:
    """One per (model, dtype).  ``emit`` fills a PlanBuilder for a batch of ``B`` images."""

    def __init__(self, emit: Callable[[PlanBuilder], None], device: torch.device, dtype: str = "bf16",
                 conv_impl: Optional[str] = None, pack_cache=None):
        if device.type != "cuda":
            raise RuntimeError("leanyolo_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        self.emit, self.device, self.dtype = emit, device, dtype
        impl = conv_impl or os.environ.get("LEANYOLO_CONV_IMPL", "auto")
        self.impl = N.IMPL_SIMT if impl == "simt" else N.IMPL_AUTO
        self.lib = N.lib()
        with torch.cuda.device(device):
            sms = C.c_int32(0)
            N.check(self.lib.ly_device_check(C.byref(sms)), "ly_device_check")
        self.sm_count = sms.value
        self._plans: Dict[Tuple[int, int, int, bool], Compiled] = {}
        self._workspaces: Dict[Tuple[int, int, int], torch.Tensor] = {}   # shared by the fp32- and uint8-input plans of a shape
        self._w: Optional[torch.Tensor] = None
        self._b: Optional[torch.Tensor] = None
        self._graphs: Dict[Tuple, Tuple[torch.cuda.CUDAGraph, torch.Tensor, Dict]] = {}
        self.pack_cache = pack_cache       # weights.PackCache or None
        self.pack_seconds = 0.0            # host time spent producing the packed parameters (fold + pack, or cache load)

    # ------------------------------------------------------------------ build
    def compile(self, B: int, H: int, W: int, in_u8: bool = False) -> Compiled:
        key = (B, H, W, in_u8)
        if key in self._plans:
            return self._plans[key]
        import time
        tc = self.impl != N.IMPL_SIMT
        pb = None
        t0 = time.perf_counter()
        if self._w is None and self.pack_cache is not None:
            # packed blobs of an earlier process, stored next to the checkpoint (keyed by its sha256 and this lowering)
            dry = PlanBuilder(B, H, W, self.dtype, tensor_core=tc, dry=True)
            self.emit(dry)
            blobs = self.pack_cache.load(dry.signature())
            if blobs is not None and blobs[0].numel() == dry._w_len and blobs[1].numel() == dry._b_len:
                self._w, self._b = blobs[0].to(self.device), blobs[1].to(self.device)
                pb = dry
        if pb is None:
            # parameters are packed ONCE per engine: later shapes lower dry (offsets only, no fp64 folding)
            pb = PlanBuilder(B, H, W, self.dtype, tensor_core=tc, dry=self._w is not None)
            self.emit(pb)
            if self._w is None:
                w, b = pb.finalize_params()
                self._w, self._b = w.to(self.device), b.to(self.device)
                if self.pack_cache is not None:
                    self.pack_cache.save(pb.signature(), w, b)
            elif self._w.numel() != pb._w_len or self._b.numel() != pb._b_len:
                raise RuntimeError("parameter packing must not depend on the input shape")
        if not self._plans:
            self.pack_seconds = time.perf_counter() - t0
        pb.assign_offsets(reuse=os.environ.get("LEANYOLO_WS_REUSE", "1") != "0")
        # the two stem variants (fp32 / uint8 image) differ in one op's flag only: same buffer layout, one workspace
        ws = self._workspaces.get((B, H, W))
        if ws is None or ws.numel() < pb.ws_bytes:
            ws = torch.empty(max(pb.ws_bytes, 1024), dtype=torch.uint8, device=self.device)
            self._workspaces[(B, H, W)] = ws
        base, wbase, bbase = ws.data_ptr(), self._w.data_ptr(), self._b.data_ptr()
        arr, chains, in_keys, out_keys = serialise_ops(pb, B, self.dtype, self.impl, base, wbase, bbase, in_u8)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(self.lib.ly_plan_create(arr, len(pb.ops), C.byref(handle)), "ly_plan_create")
        comp = Compiled(pb, B, ws, handle, out_keys, len(pb.ops), in_keys)
        self._plans[key] = comp
        return comp

    @staticmethod
    def _chain_struct(op: Op, wbase: int, bbase: int, esz: int) -> "N.LyChain":
        ex = op.extra
        ch = N.LyChain()
        ch.n_regions, ch.n_in, ch.n_stages = len(ex["regions"]), ex["n_in"], len(ex["stages"])
        ch.reserved = 2 if ex.get("stride0", 1) == 2 else 0      # stride of the first stage (include/leanyolo_b200.h)
        for i, c in enumerate(ex["regions"]):
            ch.region_c[i] = c
        for i, st in enumerate(ex["stages"]):
            s = ch.st[i]
            s.k, s.act, s.cout, s.n_src = st["k"], int(st["act"]), st["cout"], len(st["src"])
            for j, blk in enumerate(st["src"]):
                s.src[j] = N.LyChainBlk(*blk)
            s.dst = N.LyChainBlk(*(st["dst"] if st["dst"] is not None else (-1, 0, 0)))
            s.res = N.LyChainBlk(*(st["res"] if st["res"] is not None else (-1, 0, 0)))
            s.w, s.bias = wbase + esz * st["w_off"], bbase + 4 * st["b_off"]
        return ch

    # ------------------------------------------------------------------ run
    def _launch(self, comp: Compiled, x: Optional[torch.Tensor], outs: Dict[Tuple[str, int], torch.Tensor], img0: int,
                ins: Optional[Dict[Tuple[str, int], torch.Tensor]] = None) -> None:
        n_in = len(comp.in_keys)
        ext = (C.c_void_p * (1 + n_in + len(comp.out_keys)))()
        ext[0] = x.data_ptr() if x is not None else None
        for i, k in enumerate(comp.in_keys):
            ext[i + 1] = ins[k].data_ptr()
        for i, k in enumerate(comp.out_keys):
            ext[i + 1 + n_in] = outs[k].data_ptr()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.check(self.lib.ly_plan_run(comp.handle, ext, len(ext), img0, C.c_void_p(stream)), "ly_plan_run")

    def alloc_outputs(self, B: int, H: int, W: int, sub: int, in_u8: bool = False) -> Dict[Tuple[str, int], torch.Tensor]:
        comp = self.compile(sub, H, W, in_u8)      # (the plan this run uses anyway: no second plan for the shapes)
        return {k: torch.empty((B, c, h, w), dtype=torch.float32, device=self.device)
                for k, (c, h, w) in comp.pb.outputs.items()}

    def run(self, x: torch.Tensor, sub_batch: Optional[int] = None,
            outs: Optional[Dict[Tuple[str, int], torch.Tensor]] = None) -> Dict[Tuple[str, int], torch.Tensor]:
        """x: [B,3,H,W] fp32 or uint8, contiguous, on the engine's device.  Returns {(name, level): NCHW fp32}."""
        if not (x.is_cuda and x.dtype in (torch.float32, torch.uint8) and x.is_contiguous() and x.dim() == 4 and x.shape[1] == 3):
            raise ValueError("expected a contiguous CUDA tensor [B,3,H,W] of dtype float32 or uint8")
        u8 = x.dtype == torch.uint8
        B, _, H, W = x.shape
        sub = min(B, sub_batch) if sub_batch else B
        with torch.cuda.device(self.device):
            if outs is None:
                outs = self.alloc_outputs(B, H, W, sub, u8)
            for img0 in range(0, B, sub):
                n = min(sub, B - img0)
                self._launch(self.compile(n, H, W, u8), x, outs, img0)
        return outs

    def run_lb(self, descs: torch.Tensor, B: int, H: int, W: int, sub_batch: Optional[int] = None) -> Dict[Tuple[str, int], torch.Tensor]:
        """Forward of B images given as letterbox descriptors (``ly_lb_desc[B]`` on the device, see
        ``preprocess.letterbox_descs``): the stem kernel samples the letterboxed H x W pixels from the source images."""
        rec = C.sizeof(N.LyLbDesc)
        if not (descs.is_cuda and descs.dtype == torch.uint8 and descs.is_contiguous() and descs.numel() == rec * B):
            raise ValueError(f"descs must be a contiguous CUDA uint8 tensor of B ly_lb_desc records ({rec} bytes each)")
        if self.dtype != "bf16":
            raise RuntimeError("the fused letterbox loader exists for the bf16 path only")
        sub = min(B, sub_batch) if sub_batch else B
        with torch.cuda.device(self.device):
            outs = self.alloc_outputs(B, H, W, sub, "lb")
            for img0 in range(0, B, sub):
                self._launch(self.compile(min(sub, B - img0), H, W, "lb"), descs, outs, img0)
        return outs

    def run_named(self, ins: Dict[Tuple[str, int], torch.Tensor], H: int, W: int,
                  x: Optional[torch.Tensor] = None) -> Dict[Tuple[str, int], torch.Tensor]:
        """Sub-module plans: inputs are named NCHW fp32 feature maps (and/or the image ``x``); H, W = image size."""
        ts = list(ins.values()) + ([x] if x is not None else [])
        B = ts[0].shape[0]
        for t in ins.values():
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape[0] == B):
                raise ValueError("sub-module inputs must be contiguous CUDA float32 tensors with the same batch size")
        with torch.cuda.device(self.device):
            comp = self.compile(B, H, W, False)
            outs = {k: torch.empty((B, c, h, w), dtype=torch.float32, device=self.device) for k, (c, h, w) in comp.pb.outputs.items()}
            self._launch(comp, x, outs, 0, ins)
        return outs

    def profile(self, x: torch.Tensor, sub_batch: Optional[int] = None):
        """Per-op device times (ms) of one forward via CUDA events between launches.
        Returns a list of dicts {kind, ms, flops, bytes, tc, k, stride, cin, cout, hw}."""
        B, _, H, W = x.shape
        sub = min(B, sub_batch) if sub_batch else B
        rows = []
        with torch.cuda.device(self.device):
            outs = self.alloc_outputs(B, H, W, sub)
            for img0 in range(0, B, sub):
                n = min(sub, B - img0)
                comp = self.compile(n, H, W)
                ext = (C.c_void_p * (1 + len(comp.out_keys)))()
                ext[0] = x.data_ptr()
                for i, k in enumerate(comp.out_keys):
                    ext[i + 1] = outs[k].data_ptr()
                ms = (C.c_float * comp.n_launches)()
                tc = (C.c_int32 * comp.n_launches)()
                stream = torch.cuda.current_stream(self.device).cuda_stream
                N.check(self.lib.ly_plan_profile(comp.handle, ext, len(ext), img0, C.c_void_p(stream), ms, tc), "ly_plan_profile")
                es = comp.pb.esize
                for i, op in enumerate(comp.pb.ops):
                    flops = byts = 0
                    if op.kind == "chain":
                        s0 = op.extra.get("stride0", 1)
                        hw_ = (op.src.H // s0) * (op.src.W // s0)
                        flops = sum(2 * n * hw_ * st["cout"] * st["cin"] * st["k"] ** 2 for st in op.extra["stages"])
                        byts = n * (op.src.H * op.src.W * op.src.c * es + hw_ * op.cout * (es if op.dst is not None else 4))
                    elif op.kind in ("conv", "dwpw"):
                        Ho, Wo = op.src.H // op.stride, op.src.W // op.stride
                        flops = 2 * n * Ho * Wo * op.cout * op.cin * op.k * op.k
                        byts = n * (op.src.H * op.src.W * op.src.c * es + Ho * Wo * op.extra["cpad"] * (es if op.dst is not None else 4)
                                    + (Ho * Wo * op.extra["cpad"] * es if op.res is not None else 0))
                    elif op.kind == "stem":
                        flops = 2 * n * op.dst.H * op.dst.W * op.cout * 27
                        byts = n * (3 * 4 * op.dst.H * op.dst.W * 4 + op.dst.H * op.dst.W * op.dst.c * es)
                    elif op.kind == "dw":
                        byts = n * es * (op.src.H * op.src.W * op.src.c + op.dst.H * op.dst.W * op.dst.c * (2 if op.res is not None else 1))
                    elif op.kind == "pool":
                        byts = n * es * op.src.H * op.src.W * op.src.c * 4
                    elif op.kind == "up":
                        byts = n * es * op.dst.H * op.dst.W * op.dst.c * 5 // 4
                    elif op.kind == "attn":
                        nh, kdp, hd, _ = op.attn
                        t = op.src.H * op.src.W
                        flops = 2 * n * nh * t * t * (kdp + hd)
                        byts = n * es * t * (op.src.c + op.dst.c)
                    elif op.kind == "export":
                        byts = n * op.src.H * op.src.W * op.cin * (es + 4)
                    rows.append(dict(kind=op.kind, ms=float(ms[i]), flops=flops, bytes=byts, tc=int(tc[i]), k=op.k,
                                     stride=op.stride, cin=op.cin, cout=op.cout,
                                     hw=(op.src.H if op.src is not None else op.dst.H * 2)))
        return rows

    def launches_per_forward(self, B: int, H: int, W: int, sub_batch: Optional[int] = None) -> int:
        sub = min(B, sub_batch) if sub_batch else B
        total = 0
        for img0 in range(0, B, sub):
            total += self.compile(min(sub, B - img0), H, W).n_launches
        return total

    def close(self) -> None:
        for comp in self._plans.values():
            self.lib.ly_plan_destroy(comp.handle)
        self._plans.clear()
        self._workspaces.clear()

    def __deepcopy__(self, memo):
        raise TypeError("Engine owns native plan handles and device memory and cannot be copied; build a new one")

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
