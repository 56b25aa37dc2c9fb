"""GPU-resident detection tail behind the reference's decode functions.

Mirrors ``leanyolo.models.yolov10.postprocess`` (same names, keyword arguments,
return structure and edge-case behaviour) but every stage — DFL expectation,
anchor decode, sigmoid, two-stage top-k, greedy IoU NMS — runs in the CUDA
kernels of ``csrc/decode.cu`` through the C ABI.  No CPU fallback: CPU tensors
are rejected.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _native as N

MAX_DET_LIMIT = 1024  # per-CTA sort buffers in decode.cu


def _levels(preds: Sequence[torch.Tensor], num_classes: int, strides: Sequence[int], img_size=None, nms_path=False):
    if len(preds) != len(strides):
        raise ValueError("preds and strides length mismatch")
    if len(preds) > 4:
        raise ValueError("at most 4 pyramid levels are supported")
    p0 = preds[0]
    if not p0.is_cuda:
        raise RuntimeError("leanyolo_b200 decodes on CUDA only (no CPU fallback); move the head tensors to the GPU")
    keep = []
    lv = N.LyLevels()
    b, c = p0.shape[0], p0.shape[1]
    direct = bool(nms_path and c == 4 + num_classes)
    reg_max = 1 if direct else (c - num_classes) // 4
    if not direct and 4 * reg_max + num_classes != c:
        raise ValueError("Invalid channel layout for v10 head")
    for i, (p, s) in enumerate(zip(preds, strides)):
        if p.shape[0] != b or p.shape[1] != c or p.device != p0.device:
            raise ValueError("head tensors must share batch size, channel count and device")
        q = p.detach()
        if q.dtype != torch.float32 or not q.is_contiguous():
            q = q.float().contiguous()
        keep.append(q)
        lv.preds[i] = q.data_ptr()
        lv.H[i], lv.W[i], lv.stride[i] = q.shape[2], q.shape[3], int(s)
    lv.n_levels, lv.B, lv.nc, lv.reg_max = len(preds), b, int(num_classes), int(reg_max)
    lv.direct = int(direct)
    if direct and img_size is not None:
        lv.clamp_h, lv.clamp_w = int(img_size[0]), int(img_size[1])
    return lv, keep, sum(int(p.shape[2] * p.shape[3]) for p in preds)


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


@torch.no_grad()
def topk_raw(preds: Sequence[torch.Tensor], *, num_classes: int, strides: Sequence[int] = (8, 16, 32),
             max_det: int = 300, lb_meta: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Fixed-shape result: (dets [B,k,6], anchor [B,k] int32, cls [B,k] int32).  ``lb_meta`` ([B,6] fp32 on the
    device, from ``preprocess.letterbox_batch``): the boxes are mapped back to each source image's coordinates inside
    the decode kernel (``unletterbox_coords`` fused into its epilogue)."""
    if not 1 <= max_det <= MAX_DET_LIMIT:
        raise ValueError(f"max_det must be in 1..{MAX_DET_LIMIT}")
    lib = N.lib()
    lv, keep, A = _levels(preds, num_classes, strides)
    dev = keep[0].device
    k = min(max_det, A)
    with torch.cuda.device(dev):
        nbytes = lib.ly_decode_scratch_bytes(C.byref(lv), max_det)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.empty((lv.B, k, 6), dtype=torch.float32, device=dev)
        anchor = torch.empty((lv.B, k), dtype=torch.int32, device=dev)
        cls = torch.empty((lv.B, k), dtype=torch.int32, device=dev)
        if lb_meta is not None:
            if not (lb_meta.is_cuda and lb_meta.dtype == torch.float32 and tuple(lb_meta.shape) == (lv.B, 6) and lb_meta.is_contiguous()):
                raise ValueError("lb_meta must be a contiguous CUDA float32 tensor [B, 6]")
            N.check(lib.ly_decode_topk_lb(C.byref(lv), max_det, lb_meta.data_ptr(), out.data_ptr(), anchor.data_ptr(), cls.data_ptr(),
                                          scratch.data_ptr(), nbytes, _stream(dev)), "ly_decode_topk_lb")
        else:
            N.check(lib.ly_decode_topk(C.byref(lv), max_det, out.data_ptr(), anchor.data_ptr(), cls.data_ptr(),
                                       scratch.data_ptr(), nbytes, _stream(dev)), "ly_decode_topk")
        scratch.record_stream(torch.cuda.current_stream(dev))
    return out, anchor, cls


@torch.no_grad()
def decode_v10_official_topk(
    preds: Sequence[torch.Tensor],
    *,
    num_classes: int,
    strides: Sequence[int] = (8, 16, 32),
    max_det: int = 300,
    conf_thresh: Optional[float] = None,   # accepted and ignored, like the reference
    iou_thresh: Optional[float] = None,
    img_size: Optional[Tuple[int, int]] = None,
) -> List[List[torch.Tensor]]:
    """Drop-in for ``decode_v10_official_topk`` (postprocess.py:166-261): always
    ``min(max_det, A)`` rows per image, score-descending, no threshold, no clamp."""
    out, _, _ = topk_raw(preds, num_classes=num_classes, strides=strides, max_det=max_det)
    return [[out[i]] for i in range(out.shape[0])]


@torch.no_grad()
def export_decode(preds: Sequence[torch.Tensor], *, num_classes: int, strides: Sequence[int] = (8, 16, 32), imgsz=640,
                  max_dets: int = 300, conf: float = 0.25, nms: bool = False, iou: float = 0.45, pre_topk: int = 1000,
                  img0: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """The decode of the reference's ONNX wrapper (``YOLOv10ONNXExport.forward``, models/yolov10/export.py:97-198) on
    the GPU: fixed-shape ``detections [B, k, 6]`` and ``num_dets [B]`` (int64).  ``nms=False``: top-k anchors by best
    class score with the confidence mask and the clamp to the image; ``nms=True``: class-wise NMS over the
    ``pre_topk`` best (anchor, class) pairs, bit-compatible with the reference's class-offset + torchvision NMS
    (``img0`` = global index of the first image, which enters the reference's fp32 offset arithmetic when a batch
    is sharded).  ``imgsz``: int (square, like the reference) or (H, W).  max_dets, pre_topk <= 1024."""
    if not 1 <= max_dets <= MAX_DET_LIMIT or not 1 <= pre_topk <= MAX_DET_LIMIT:
        raise ValueError(f"max_dets and pre_topk must be in 1..{MAX_DET_LIMIT}")
    ih, iw = (int(imgsz), int(imgsz)) if isinstance(imgsz, (int, float)) else (int(imgsz[0]), int(imgsz[1]))
    lib = N.lib()
    lv, keep, A = _levels(preds, num_classes, strides)
    dev = keep[0].device
    k = min(max_dets, A) if not nms else min(max_dets, pre_topk, A * int(num_classes))
    with torch.cuda.device(dev):
        nbytes = lib.ly_decode_export_scratch_bytes(C.byref(lv), max_dets, pre_topk)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.empty((lv.B, k, 6), dtype=torch.float32, device=dev)
        num = torch.empty((lv.B,), dtype=torch.int32, device=dev)
        N.check(lib.ly_decode_export(C.byref(lv), float(conf), max_dets, int(bool(nms)), float(iou), pre_topk, ih, iw, int(img0),
                                     out.data_ptr(), num.data_ptr(), scratch.data_ptr(), nbytes, _stream(dev)), "ly_decode_export")
        scratch.record_stream(torch.cuda.current_stream(dev))
    return out, num.to(torch.int64)


@torch.no_grad()
def nms_raw(preds: Sequence[torch.Tensor], *, num_classes: int, strides: Sequence[int] = (8, 16, 32),
            conf_thresh: float = 0.25, iou_thresh: float = 0.45, max_det: int = 300,
            img_size: Optional[Tuple[int, int]] = None, classwise: bool = False):
    """Fixed-shape result: (dets [B,max_det,6] zero padded, count [B] int32, anchor [B,max_det] int32, -1 padded)."""
    if not 1 <= max_det <= MAX_DET_LIMIT:
        raise ValueError(f"max_det must be in 1..{MAX_DET_LIMIT}")
    lib = N.lib()
    lv, keep, A = _levels(preds, num_classes, strides, img_size, nms_path=True)
    dev = keep[0].device
    with torch.cuda.device(dev):
        nbytes = lib.ly_decode_scratch_bytes(C.byref(lv), max_det)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        out = torch.empty((lv.B, max_det, 6), dtype=torch.float32, device=dev)
        count = torch.empty((lv.B,), dtype=torch.int32, device=dev)
        anchor = torch.empty((lv.B, max_det), dtype=torch.int32, device=dev)
        N.check(lib.ly_decode_nms(C.byref(lv), float(conf_thresh), float(iou_thresh), max_det, int(classwise),
                                  out.data_ptr(), count.data_ptr(), anchor.data_ptr(), scratch.data_ptr(), nbytes,
                                  _stream(dev)), "ly_decode_nms")
        scratch.record_stream(torch.cuda.current_stream(dev))
    return out, count, anchor


@torch.no_grad()
def decode_v10_predictions(
    preds: Sequence[torch.Tensor],
    *,
    num_classes: int,
    strides: Sequence[int] = (8, 16, 32),
    conf_thresh: float = 0.25,
    iou_thresh: float = 0.45,
    max_det: int = 300,
    img_size: Optional[Tuple[int, int]] = None,
    classwise: bool = False,
) -> List[List[torch.Tensor]]:
    """Drop-in for ``decode_v10_predictions`` (postprocess.py:47-163): ragged
    ``[n_i, 6]`` per image, empty images give ``torch.empty((0, 6))``.  The
    default is class-agnostic suppression, which is what the reference code does;
    ``classwise=True`` restricts suppression to equal labels."""
    out, count, _ = nms_raw(preds, num_classes=num_classes, strides=strides, conf_thresh=conf_thresh,
                            iou_thresh=iou_thresh, max_det=max_det, img_size=img_size, classwise=classwise)
    counts = count.tolist()  # the only device->host sync of the decode
    return [[out[i, :n]] if n > 0 else [torch.empty((0, 6), device=out.device)] for i, n in enumerate(counts)]


@torch.no_grad()
def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_thresh: float, *, labels: Optional[torch.Tensor] = None,
        max_keep: Optional[int] = None) -> torch.Tensor:
    """Drop-in for ``leanyolo.utils.box_ops.nms`` on CUDA tensors: keep indices
    (int64) into the input order, score-descending.  ``labels`` switches to
    class-wise suppression.  ``max_keep`` (<= 1024) truncates.  The kernel keeps at most 1024 boxes (its kept set lives
    in shared memory): with ``max_keep=None`` a result that reaches that cap raises instead of silently truncating
    what the reference would return in full."""
    if not boxes.is_cuda:
        raise RuntimeError("leanyolo_b200 nms runs on CUDA only (no CPU fallback)")
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros((0,), dtype=torch.long, device=boxes.device)
    mk = min(max_keep or MAX_DET_LIMIT, MAX_DET_LIMIT)
    lib = N.lib()
    dev = boxes.device
    with torch.cuda.device(dev):
        b = boxes.detach().float().contiguous()
        s = scores.detach().float().contiguous()
        l = labels.detach().to(torch.int32).contiguous() if labels is not None else None
        nbytes = lib.ly_nms_scratch_bytes(1, n)
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        keep = torch.empty((1, mk), dtype=torch.int32, device=dev)
        cnt = torch.empty((1,), dtype=torch.int32, device=dev)
        N.check(lib.ly_nms(b.data_ptr(), s.data_ptr(), l.data_ptr() if l is not None else None, None, 1, n,
                           float(iou_thresh), mk, int(l is not None), keep.data_ptr(), cnt.data_ptr(),
                           scratch.data_ptr(), nbytes, _stream(dev)), "ly_nms")
        k = int(cnt.item())
    if max_keep is None and k >= MAX_DET_LIMIT and n > MAX_DET_LIMIT:
        raise RuntimeError(f"nms: {MAX_DET_LIMIT} or more boxes survive; leanyolo_b200 keeps at most {MAX_DET_LIMIT} "
                           "(pass max_keep to truncate explicitly)")
    return keep[0, :k].to(torch.long)
