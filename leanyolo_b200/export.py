"""Export-style fixed-shape detections on the GPU.

Mirrors the reference's ONNX wrapper ``leanyolo.models.yolov10.export.YOLOv10ONNXExport`` (export.py:34-198: same
constructor arguments, ``forward(images) -> (detections [B,N,6], num_dets [B] int64)``) with the model forward and the
whole decode -- DFL, sigmoid, confidence mask, top-k or class-wise pre-top-k NMS, clamp -- in the CUDA kernels of this
package.  ONNX serialisation itself (``export_onnx``) is out of scope: there is no ONNX runtime for hand-written
sm_100a kernels.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.nn as nn

from . import postprocess as PP


class YOLOv10ONNXExport(nn.Module):
    def __init__(self, model: nn.Module, *, imgsz: int = 640, max_dets: int = 300, conf: float = 0.25,
                 strides: Sequence[int] = (8, 16, 32), nms: bool = False, iou: float = 0.45, pre_topk: int = 1000) -> None:
        super().__init__()
        self.model = model.eval()
        self.imgsz = int(imgsz)
        self.max_dets = int(max(1, max_dets))
        self.conf = float(conf)
        self.strides = tuple(int(s) for s in strides)
        self.nms = bool(nms)
        self.iou = float(iou)
        self.pre_topk = int(max(1, pre_topk))
        head = getattr(self.model, "head", None)
        if head is None or not hasattr(head, "nc") or not hasattr(head, "reg_max"):
            raise ValueError("Provided model does not appear to be a YOLOv10 model with V10Detect head.")
        self.num_classes = int(head.nc)
        self.reg_max = int(head.reg_max)

    @torch.no_grad()
    def forward(self, images: torch.Tensor, img0: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        preds = self.model(images)
        if not isinstance(preds, (list, tuple)):
            raise RuntimeError("Unexpected model output; expected list of tensors per scale.")
        if len(preds) != len(self.strides):
            raise ValueError("preds/strides mismatch")
        return PP.export_decode(preds, num_classes=self.num_classes, strides=self.strides, imgsz=self.imgsz,
                                max_dets=self.max_dets, conf=self.conf, nms=self.nms, iou=self.iou, pre_topk=self.pre_topk,
                                img0=img0)


def build_export_wrapper(model: nn.Module, *, imgsz: int = 640, max_dets: int = 300, conf: float = 0.25, decode: str = "topk",
                         iou: float = 0.45, pre_topk: int = 1000) -> YOLOv10ONNXExport:
    """export.py:201-221."""
    return YOLOv10ONNXExport(model, imgsz=int(imgsz), max_dets=int(max_dets), conf=float(conf), nms=decode.lower() == "nms",
                             iou=float(iou), pre_topk=int(pre_topk))
