"""YOLOv10 model façade: the reference's ``nn.Module`` surface over the CUDA plan.

``model(x)`` and ``model.decode_forward(raw)`` keep the reference contract
(leanyolo/models/yolov10/yolov10s.py:105-144): eval forward returns the three
contiguous NCHW fp32 one2many head tensors and caches both branches in
``_eval_branches``; ``decode_forward`` runs the top-k decode on the one2one
branch.  All arithmetic happens in hand-written sm_100a kernels
(``csrc/``); there is no PyTorch or CPU compute fallback.

Extras that do not change the reference surface:
* ``model.precision``  "bf16" (tensor-core hot path) or "fp32" (CUDA-core check mode).
* ``model.detect(x)``  forward + GPU top-k decode returning a fixed ``[B,k,6]`` tensor.
* ``model.sub_batch``  images per plan sweep (keeps producer->consumer tensors in L2).
"""
from __future__ import annotations

import os
import weakref
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import postprocess as PP
from .engine import Engine
from .modules import Backbone, Detect, Neck
from .variants import REG_MAX, STRIDES, VARIANTS, Variant


class YOLOv10(nn.Module):
    variant: Variant

    def __init__(self, *, class_names: Sequence[str], in_channels: int = 3,
                 input_norm_subtract: Sequence[float] = (0.0, 0.0, 0.0),
                 input_norm_divide: Sequence[float] = (255.0, 255.0, 255.0)):
        super().__init__()
        if in_channels != 3:
            raise ValueError("leanyolo_b200 supports 3-channel RGB input only")
        self.class_names = list(class_names)
        self.register_buffer("input_subtract", torch.tensor(list(input_norm_subtract), dtype=torch.float32).view(1, 3, 1, 1))
        self.register_buffer("input_divide", torch.tensor(list(input_norm_divide), dtype=torch.float32).view(1, 3, 1, 1))
        v = self.variant
        self.backbone = Backbone(v, in_channels)
        self.neck = Neck(v, *self.backbone.out_c)
        self.head = Detect(len(self.class_names), self.neck.out_c, REG_MAX)
        self.precision = os.environ.get("LEANYOLO_PRECISION", "bf16")
        sb = os.environ.get("LEANYOLO_SUB_BATCH")
        self.sub_batch: Optional[int] = int(sb) if sb else None
        self._engines: Dict[tuple, Engine] = {}
        self._eval_branches: Dict[str, List[torch.Tensor]] = {}
        self._weights_source = None      # (path, sha256) of the checkpoint the parameters came from (registry.get_model)
        # the sub-modules are callable like the reference's (model.backbone(x), model.neck(c3, c4, c5),
        # model.head(feats), head.forward_feat(feats, cv2, cv3)): they run their slice of the plan through this model
        for m in (self.backbone, self.neck, self.head):
            object.__setattr__(m, "_root", weakref.ref(self))

    def __deepcopy__(self, memo):
        """``copy.deepcopy(model)`` (EMA / clone patterns): parameters and buffers are copied, compiled engines are NOT
        (they own native plan handles and device workspaces and are rebuilt on first use), and the sub-modules of the
        copy dispatch to the copy."""
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k in ("_engines", "_eval_branches") else copy.deepcopy(v, memo)
        for m in (new.backbone, new.neck, new.head):
            object.__setattr__(m, "_root", weakref.ref(new))
        return new

    # ---- class tables kept for callers that introspect them (yolov10s.py:62-65)
    @property
    def CH(self):
        return dict(enumerate(self.variant.width))

    @property
    def HCH(self):
        return dict(self.variant.neck)

    @property
    def REPS(self):
        return dict(self.variant.reps)

    # ------------------------------------------------------------------ lowering
    def emit(self, pb, taps: bool = False, head_only: Optional[str] = None) -> None:
        """Lower the whole eval forward (both head branches, or just ``head_only``) into ``pb``."""
        bb, nk, hd = self.backbone, self.neck, self.head
        w0, b0 = bb.cv0.folded(pb)
        x = pb.stem(w0, b0, self.input_subtract.flatten().tolist(), self.input_divide.flatten().tolist())
        cats = nk.concat_buffers(pb, pb.H // 8, pb.W // 8)
        c3w, c4w, c5w, h13, h16, h19, _ = nk.widths
        c3, c4, c5 = bb.emit(pb, x, cats["cat_p3"].view(h13, c3w), cats["cat_p4"].view(c5w, c4w),
                             cats["cat_n5"].view(h19, c5w))
        p3, p4, p5 = nk.emit(pb, cats)
        if taps:
            for name, v, c in (("c3", c3, c3w), ("c4", c4, c4w), ("c5", c5, c5w), ("p3", p3, nk.out_c[0]),
                               ("p4", p4, nk.out_c[1]), ("p5", p5, nk.out_c[2])):
                pb.export_nchw(v, name, c)
        hd.emit(pb, (p3, p4, p5), only=head_only)

    # ------------------------------------------------------------------ sub-module plans
    def _emit_part(self, pb, part: str) -> None:
        bb, nk, hd = self.backbone, self.neck, self.head
        c3w, c4w, c5w, h13, h16, h19, _ = nk.widths
        h8, w8 = pb.H // 8, pb.W // 8
        if part == "backbone":
            # the reference's backbone takes the already-normalised image (yolov10s.py:107-114)
            w0, b0 = bb.cv0.folded(pb)
            x = pb.stem(w0, b0, [0.0, 0.0, 0.0], [1.0, 1.0, 1.0])
            for name, v, c in zip(("c3", "c4", "c5"), bb.emit(pb, x), bb.out_c):
                pb.export_nchw(v, name, c)
        elif part == "neck":
            cats = nk.concat_buffers(pb, h8, w8)
            pb.import_nchw("c3", c3w, h8, w8, cats["cat_p3"].view(h13, c3w))
            pb.import_nchw("c4", c4w, h8 // 2, w8 // 2, cats["cat_p4"].view(c5w, c4w))
            pb.import_nchw("c5", c5w, h8 // 4, w8 // 4, cats["cat_n5"].view(h19, c5w))
            for name, v, c in zip(("p3", "p4", "p5"), nk.emit(pb, cats), nk.out_c):
                pb.export_nchw(v, name, c)
        elif part == "head":
            feats = [pb.import_nchw(f"p{3 + i}", c, h8 >> i, w8 >> i) for i, c in enumerate(nk.out_c)]
            hd.emit(pb, feats)
        else:
            raise ValueError(part)

    def _run_part(self, part: str, ins: Dict[str, torch.Tensor], hw, x: Optional[torch.Tensor] = None):
        if self.training:
            raise NotImplementedError("leanyolo_b200 is inference-only: call model.eval() first")
        dev = self.input_subtract.device
        ts = list(ins.values()) + ([x] if x is not None else [])
        if dev.type != "cuda" or any(not t.is_cuda for t in ts):
            raise RuntimeError("leanyolo_b200 runs on CUDA (sm_100a) only: move the model and the inputs to the GPU; "
                               "there is no CPU fallback")
        prec = "f32" if self.precision in ("fp32", "f32", "float32") else "bf16"
        key = (str(dev), prec, "part:" + part)
        if key not in self._engines:
            self._engines[key] = Engine(lambda pb: self._emit_part(pb, part), dev, prec)
        named = {(k, 0): v.detach().to(torch.float32).contiguous() for k, v in ins.items()}
        xin = x.detach().to(torch.float32).contiguous() if x is not None else None
        return self._engines[key].run_named(named, hw[0], hw[1], xin)

    def invalidate(self, params_changed: bool = True) -> None:
        """Drop packed weights / plans (call after mutating parameters in place).  ``params_changed`` also forgets
        which checkpoint the parameters came from (the on-disk pack cache is keyed by it)."""
        for e in self._engines.values():
            e.close()
        self._engines.clear()
        if params_changed:
            self._weights_source = None

    def load_state_dict(self, *args, **kwargs):
        ret = super().load_state_dict(*args, **kwargs)
        self.invalidate()
        return ret

    def _apply(self, fn, *args, **kwargs):
        ret = super()._apply(fn, *args, **kwargs)
        self.invalidate(params_changed=False)     # a device move keeps the values (a dtype change is caught in _pack_cache)
        return ret

    def _pack_cache(self, prec: str):
        """weights.PackCache for this model's checkpoint, or None (random init, mutated or non-fp32 parameters,
        LEANYOLO_PACK_CACHE=0)."""
        src = getattr(self, "_weights_source", None)
        if src is None or os.environ.get("LEANYOLO_PACK_CACHE", "1") == "0":
            return None
        if any(p.dtype != torch.float32 for p in self.parameters()):
            return None
        from .weights import PackCache
        return PackCache(src[0], src[1], f"{self.variant.name}.{prec}.nc{len(self.class_names)}")

    def engine(self, device: torch.device, taps: bool = False, head_only: Optional[str] = None) -> Engine:
        prec = "f32" if self.precision in ("fp32", "f32", "float32") else "bf16"
        key = (str(device), prec, taps) if head_only is None else (str(device), prec, taps, head_only)
        if key not in self._engines:
            self._engines[key] = Engine(lambda pb: self.emit(pb, taps, head_only), device, prec,
                                        pack_cache=None if (taps or head_only) else self._pack_cache(prec))
        return self._engines[key]

    # ------------------------------------------------------------------ reference surface
    def _run(self, x: torch.Tensor, taps: bool = False, head_only: Optional[str] = None):
        if self.training:
            raise NotImplementedError("leanyolo_b200 is inference-only: call model.eval() first "
                                      "(training stays with the reference implementation)")
        dev = self.input_subtract.device
        if dev.type != "cuda" or not x.is_cuda:
            raise RuntimeError("leanyolo_b200 runs on CUDA (sm_100a) only: move the model and the input to the GPU; "
                               "there is no CPU fallback")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] % 32 or x.shape[3] % 32:
            raise ValueError("expected input [B,3,H,W] with H and W multiples of 32")
        x = x.detach()
        if x.dtype != torch.uint8:     # uint8 images go straight to the stem kernel (x.float() happens in its loader)
            x = x.to(dtype=torch.float32)
        x = x.contiguous()
        return self.engine(dev, taps, head_only).run(x, self.sub_batch)

    @torch.no_grad()
    def detect_letterboxed(self, descs: torch.Tensor, n_images: int, hw, meta: Optional[torch.Tensor] = None, max_det: int = 300,
                           one2one_only: bool = False) -> torch.Tensor:
        """forward + top-k decode of ``n_images`` SOURCE images of any sizes described by ``descs``
        (``preprocess.letterbox_descs``): the letterbox (utils/letterbox.py:9-91) runs inside the stem kernel's loader
        and, with ``meta``, the unletterbox (utils/box_ops.py:96-124) inside the decode kernel -- the letterboxed batch
        never exists in memory.  bf16 path only."""
        if self.training:
            raise NotImplementedError("leanyolo_b200 is inference-only: call model.eval() first")
        dev = self.input_subtract.device
        if dev.type != "cuda":
            raise RuntimeError("leanyolo_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        H, W = (int(hw), int(hw)) if isinstance(hw, int) else (int(hw[0]), int(hw[1]))
        if H % 32 or W % 32:
            raise ValueError("the letterbox target must be a multiple of 32")
        outs = self.engine(dev, False, "one2one" if one2one_only else None).run_lb(descs, n_images, H, W, self.sub_batch)
        keys = ("one2one",) if one2one_only else ("one2many", "one2one")
        self._eval_branches = {k: [outs[(k, i)] for i in range(self.head.nl)] for k in keys}
        out, _, _ = PP.topk_raw(self._eval_branches["one2one"], num_classes=len(self.class_names), strides=STRIDES,
                                max_det=max_det, lb_meta=meta)
        return out

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        outs = self._run(x)
        self._eval_branches = {k: [outs[(k, i)] for i in range(self.head.nl)] for k in ("one2many", "one2one")}
        return self._eval_branches["one2many"]

    @torch.no_grad()
    def forward_with_taps(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Forward that also exports c3,c4,c5,p3,p4,p5 as NCHW fp32 (the reference fidelity taps)."""
        outs = self._run(x, taps=True)
        res = {k: outs[(k, 0)] for k in ("c3", "c4", "c5", "p3", "p4", "p5")}
        for k in ("one2many", "one2one"):
            res[k] = [outs[(k, i)] for i in range(self.head.nl)]
        return res

    @torch.no_grad()
    def decode_forward(self, raw):
        if isinstance(raw, dict):
            seq = raw.get("one2one", raw.get("one2many"))
        else:
            seq = getattr(self, "_eval_branches", {}).get("one2one", raw)
        return PP.decode_v10_official_topk(seq, num_classes=len(self.class_names), strides=STRIDES)

    @torch.no_grad()
    def detect(self, x: torch.Tensor, max_det: int = 300, one2one_only: bool = False,
               lb_meta: Optional[torch.Tensor] = None) -> torch.Tensor:
        """forward + top-k decode, detections ``[B, min(max_det, A), 6]`` left on the device.

        Default: exactly ``decode_forward(model(x))`` -- both head branches run and are cached in ``_eval_branches``
        like the reference's eval forward (yolov10s.py:105-122).  ``one2one_only=True`` is the opt-in fused path
        (SURVEY hard part 6): the one-to-many branch, which the top-k decode never reads, is not computed at all
        (-3.2 GFLOP per image on yolov10s, no one-to-many NCHW tensors written).  Same arithmetic per output; the first
        regression conv runs as its own N = c2 GEMM instead of the merged two-branch one, so its fp32 summation order
        (hence a bf16 rounding here and there) may differ: detections agree to ~1e-2 px / 1e-4 in score, bit for bit
        when both lower to the same conv mode.  ``_eval_branches`` then holds only ``one2one``.  ``lb_meta`` (from ``preprocess.letterbox_batch``): boxes are
        returned in each source image's own coordinates (unletterbox fused into the decode kernel)."""
        if one2one_only:
            outs = self._run(x, head_only="one2one")
            self._eval_branches = {"one2one": [outs[("one2one", i)] for i in range(self.head.nl)]}
        else:
            self.forward(x)
        out, _, _ = PP.topk_raw(self._eval_branches["one2one"], num_classes=len(self.class_names), strides=STRIDES,
                                max_det=max_det, lb_meta=lb_meta)
        return out


def _make(name: str):
    return type("YOLOv10" + name[-1], (YOLOv10,), {"variant": VARIANTS[name], "__doc__": f"{name} (see variants.py)"})


YOLOv10n, YOLOv10s, YOLOv10m = _make("yolov10n"), _make("yolov10s"), _make("yolov10m")
YOLOv10b, YOLOv10l, YOLOv10x = _make("yolov10b"), _make("yolov10l"), _make("yolov10x")
MODEL_CLASSES = {"yolov10n": YOLOv10n, "yolov10s": YOLOv10s, "yolov10m": YOLOv10m,
                 "yolov10b": YOLOv10b, "yolov10l": YOLOv10l, "yolov10x": YOLOv10x}
