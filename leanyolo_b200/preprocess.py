"""GPU pre/post-processing on either side of the hot path.

Drop-ins for the reference's ``letterbox`` (leanyolo/utils/letterbox.py:9-91) and
``unletterbox_coords`` (leanyolo/utils/box_ops.py:96-124) on CUDA tensors, plus the batched forms
the serving loop of ``tools/infer.py:110-138`` needs at tens of thousands of images per second:

    batch, meta = letterbox_batch(images, 640)        # list of uint8 HWC cuda tensors -> [B,3,640,640] uint8
    dets = model.detect(batch)                        # the stem kernel consumes uint8 NCHW directly
    unletterbox_dets(dets, meta)                      # in place, back to each image's own coordinates

The geometry (scale, rounded new size, pad split) is host arithmetic restated from the reference;
the resize is cv2's 8-bit fixed-point bilinear, bit for bit, in ``csrc/preprocess.cu``.  CUDA only:
CPU tensors raise (no fallback).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import torch

from . import _native as N


def letterbox_params(orig_h: int, orig_w: int, new_shape=640, auto: bool = False, scale_fill: bool = False,
                     scaleup: bool = True, stride: int = 32):
    """letterbox.py:44-80 -> (new_w, new_h, left, top, right, bottom, gain_w, gain_h)."""
    if isinstance(new_shape, int):
        tgt_h, tgt_w = new_shape, new_shape
    else:
        tgt_h, tgt_w = int(new_shape[0]), int(new_shape[1])
    if scale_fill:
        gain_w, gain_h = tgt_w / max(orig_w, 1), tgt_h / max(orig_h, 1)
        new_w, new_h, pad_w, pad_h = tgt_w, tgt_h, 0.0, 0.0
    else:
        r = min(tgt_w / max(orig_w, 1), tgt_h / max(orig_h, 1))
        if not scaleup:
            r = min(r, 1.0)
        new_w, new_h = int(round(orig_w * r)), int(round(orig_h * r))
        gain_w = gain_h = r
        pad_w, pad_h = float(tgt_w - new_w), float(tgt_h - new_h)
        if auto and stride > 1:
            pad_w, pad_h = pad_w % stride, pad_h % stride
    left = int(round(pad_w / 2.0)); right = int(round(pad_w - left))
    top = int(round(pad_h / 2.0)); bottom = int(round(pad_h - top))
    return new_w, new_h, left, top, right, bottom, float(gain_w), float(gain_h)


def letterbox_params_batch(hs, ws, new_shape=640, scale_fill: bool = False, scaleup: bool = True):
    """``letterbox_params`` for a whole batch at once (numpy, no Python loop over images): arrays
    (new_w, new_h, left, top, gain_w, gain_h).  Python's round() and numpy's rint() both round half to even."""
    import numpy as np
    hs, ws = np.asarray(hs, dtype=np.float64), np.asarray(ws, dtype=np.float64)
    tgt_h, tgt_w = (new_shape, new_shape) if isinstance(new_shape, int) else (int(new_shape[0]), int(new_shape[1]))
    if scale_fill:
        gw, gh = tgt_w / np.maximum(ws, 1), tgt_h / np.maximum(hs, 1)
        new_w, new_h = np.full(ws.shape, tgt_w, dtype=np.int64), np.full(hs.shape, tgt_h, dtype=np.int64)
        pad_w, pad_h = np.zeros_like(ws), np.zeros_like(hs)
    else:
        r = np.minimum(tgt_w / np.maximum(ws, 1), tgt_h / np.maximum(hs, 1))
        if not scaleup:
            r = np.minimum(r, 1.0)
        new_w, new_h = np.rint(ws * r).astype(np.int64), np.rint(hs * r).astype(np.int64)
        gw = gh = r
        pad_w, pad_h = (tgt_w - new_w).astype(np.float64), (tgt_h - new_h).astype(np.float64)
    left, top = np.rint(pad_w / 2.0).astype(np.int64), np.rint(pad_h / 2.0).astype(np.int64)
    return new_w, new_h, left, top, gw, gh


def letterbox_descs(images: Sequence[torch.Tensor], new_shape=640, scale_fill: bool = False, scaleup: bool = True):
    """Descriptors for the fused path (``model.detect_letterboxed``): (descs uint8 on the device = ly_lb_desc[B],
    meta [B,6] fp32 on the device).  One host->device copy each; the geometry is vectorised."""
    import numpy as np
    if len(images) == 0:
        raise ValueError("empty batch")
    for img in images:
        _check_img(img)
    dev = images[0].device
    hs = np.array([img.shape[0] for img in images], dtype=np.int64)
    ws = np.array([img.shape[1] for img in images], dtype=np.int64)
    new_w, new_h, left, top, gw, gh = letterbox_params_batch(hs, ws, new_shape, scale_fill, scaleup)
    a = np.zeros(len(images), dtype=np.dtype([("src", "<u8"), ("pitch", "<i8"), ("sh", "<i4"), ("sw", "<i4"), ("nh", "<i4"),
                                              ("nw", "<i4"), ("top", "<i4"), ("left", "<i4")]))
    a["src"] = [img.data_ptr() for img in images]
    a["pitch"] = [img.stride(0) for img in images]
    a["sh"], a["sw"], a["nh"], a["nw"], a["top"], a["left"] = hs, ws, new_h, new_w, top, left
    meta = np.stack([gw, gh, left, top, hs, ws], 1).astype(np.float32)
    return torch.from_numpy(a.view(np.uint8).copy()).to(dev), torch.from_numpy(meta).to(dev)


def _check_img(img: torch.Tensor) -> None:
    if not isinstance(img, torch.Tensor) or not img.is_cuda:
        raise RuntimeError("leanyolo_b200 letterbox runs on CUDA tensors only (no CPU fallback)")
    if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3 or img.stride(2) != 1 or img.stride(1) != 3:
        raise ValueError("expected a uint8 HWC RGB image tensor with packed pixels")
    if img.shape[0] < 1 or img.shape[1] < 1 or img.shape[0] > 32767 or img.shape[1] > 32767:
        raise ValueError("image size out of range")   # cv2's fixed-point tables are 16-bit offsets as well


def _descs(images: Sequence[torch.Tensor], geo, dev) -> torch.Tensor:
    """The DEVICE array of ``ly_lb_desc`` for a batch (one small host->device copy)."""
    import numpy as np
    a = np.zeros(len(images), dtype=np.dtype([("src", "<u8"), ("pitch", "<i8"), ("sh", "<i4"), ("sw", "<i4"), ("nh", "<i4"),
                                              ("nw", "<i4"), ("top", "<i4"), ("left", "<i4")]))
    a["src"] = [img.data_ptr() for img in images]
    a["pitch"] = [img.stride(0) for img in images]
    a["sh"] = [img.shape[0] for img in images]
    a["sw"] = [img.shape[1] for img in images]
    a["nw"], a["nh"], a["left"], a["top"] = [g[0] for g in geo], [g[1] for g in geo], [g[2] for g in geo], [g[3] for g in geo]
    assert a.dtype.itemsize == C.sizeof(N.LyLbDesc)
    return torch.from_numpy(a.view(np.uint8)).to(dev)


def _run(d_descs: torch.Tensor, B: int, out: torch.Tensor, chw: bool, color) -> None:
    dev = out.device
    fill = (C.c_uint8 * 3)(*[int(c) for c in color])
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        N.check(N.lib().ly_letterbox_u8(d_descs.data_ptr(), B, out.data_ptr(), out.shape[-2] if chw else out.shape[-3],
                                        out.shape[-1] if chw else out.shape[-2], int(chw), fill, stream), "ly_letterbox_u8")
        d_descs.record_stream(torch.cuda.current_stream(dev))


def _launch(images: Sequence[torch.Tensor], geo, out: torch.Tensor, chw: bool, color) -> None:
    _run(_descs(images, geo, out.device), len(images), out, chw, color)


@torch.no_grad()
def letterbox(img: torch.Tensor, new_shape=640, color: Tuple[int, int, int] = (114, 114, 114), auto: bool = False,
              scale_fill: bool = False, scaleup: bool = True, stride: int = 32):
    """Reference signature and return value, on a CUDA uint8 HWC tensor: (img_out HWC, (gain_w, gain_h), (left, top))."""
    _check_img(img)
    g = letterbox_params(img.shape[0], img.shape[1], new_shape, auto, scale_fill, scaleup, stride)
    new_w, new_h, left, top, right, bottom, gw, gh = g
    out = torch.empty((1, new_h + top + bottom, new_w + left + right, 3), dtype=torch.uint8, device=img.device)
    _launch([img], [g], out, False, color)
    return out[0], (gw, gh), (left, top)


@torch.no_grad()
def letterbox_batch(images: Sequence[torch.Tensor], new_shape=640, color: Tuple[int, int, int] = (114, 114, 114),
                    scale_fill: bool = False, scaleup: bool = True):
    """B images of any sizes -> one uint8 NCHW batch ``[B,3,H,W]`` (ONE launch) + meta ``[B,6]`` fp32 on the device =
    (gain_w, gain_h, pad_left, pad_top, orig_h, orig_w) per image, the argument of ``unletterbox_dets``."""
    if len(images) == 0:
        raise ValueError("empty batch")
    for img in images:
        _check_img(img)
    tgt = (new_shape, new_shape) if isinstance(new_shape, int) else (int(new_shape[0]), int(new_shape[1]))
    geo = [letterbox_params(img.shape[0], img.shape[1], tgt, False, scale_fill, scaleup) for img in images]
    dev = images[0].device
    out = torch.empty((len(images), 3, tgt[0], tgt[1]), dtype=torch.uint8, device=dev)
    _launch(images, geo, out, True, color)
    meta = torch.tensor([[g[6], g[7], g[2], g[3], img.shape[0], img.shape[1]] for g, img in zip(geo, images)],
                        dtype=torch.float32).to(dev)
    return out, meta


@torch.no_grad()
def unletterbox_dets(dets: torch.Tensor, meta: torch.Tensor) -> torch.Tensor:
    """In place: xyxy columns of ``dets [B,K,>=4]`` fp32 back to original-image coordinates (box_ops.py:96-124)."""
    if not dets.is_cuda or not meta.is_cuda:
        raise RuntimeError("leanyolo_b200 unletterbox runs on CUDA tensors only (no CPU fallback)")
    if not (dets.dtype == torch.float32 and dets.dim() == 3 and dets.shape[2] >= 4 and dets.is_contiguous()):
        raise ValueError("dets must be a contiguous float32 tensor [B, K, >=4]")
    if not (meta.dtype == torch.float32 and tuple(meta.shape) == (dets.shape[0], 6) and meta.is_contiguous()):
        raise ValueError("meta must be a contiguous float32 tensor [B, 6]")
    if dets.shape[1] == 0:
        return dets
    with torch.cuda.device(dets.device):
        stream = C.c_void_p(torch.cuda.current_stream(dets.device).cuda_stream)
        N.check(N.lib().ly_unletterbox(dets.data_ptr(), dets.shape[0], dets.shape[1], dets.shape[2], meta.data_ptr(), stream),
                "ly_unletterbox")
    return dets


@torch.no_grad()
def unletterbox_coords(boxes: torch.Tensor, gain: Tuple[float, float], pad: Tuple[int, int], to_shape: Tuple[int, int]) -> torch.Tensor:
    """Reference signature (box_ops.py:96-101): boxes ``[N,4]`` xyxy -> new tensor in original-image coordinates."""
    if not isinstance(boxes, torch.Tensor) or not boxes.is_cuda:
        raise RuntimeError("leanyolo_b200 unletterbox runs on CUDA tensors only (no CPU fallback)")
    out = boxes.detach().to(torch.float32).reshape(1, -1, 4).contiguous().clone()
    meta = torch.tensor([[gain[0], gain[1], pad[0], pad[1], to_shape[0], to_shape[1]]], dtype=torch.float32).to(boxes.device)
    return unletterbox_dets(out, meta).reshape(boxes.shape)


@torch.no_grad()
def detect_images(model, images: Sequence[torch.Tensor], imgsz: int = 640, max_det: int = 300) -> List[torch.Tensor]:
    """The loop body of tools/infer.py:110-138 for a batch: letterbox -> forward -> top-k decode -> unletterbox.
    Returns one ``[k,6]`` tensor per image in ITS OWN pixel coordinates."""
    if getattr(model, "precision", "bf16") == "bf16" and _fused_ok(model):
        # letterbox inside the stem loader, unletterbox inside the decode kernel: no letterboxed batch in memory
        descs, meta = letterbox_descs(images, imgsz)
        dets = model.detect_letterboxed(descs, len(images), imgsz, meta, max_det=max_det)
        for img in images:                     # the source images must outlive the asynchronous forward
            img.record_stream(torch.cuda.current_stream(img.device))
        return list(dets.unbind(0))
    batch, meta = letterbox_batch(images, imgsz)
    dets = model.detect(batch, max_det=max_det, lb_meta=meta)      # unletterbox happens inside the decode kernel
    return list(dets.unbind(0))


def _fused_ok(model) -> bool:
    """The fused letterbox loader lives in the tensor-core stem kernel: bf16, first conv <= 80 channels."""
    import os
    return (os.environ.get("LEANYOLO_FUSE_LETTERBOX", "1") != "0" and os.environ.get("LEANYOLO_CONV_IMPL", "auto") != "simt"
            and model.backbone.cv0.conv.out_channels <= 80)
