"""leanyolo_b200 — B200-native YOLOv10 inference, drop-in for leanyolo's public API.

    from leanyolo_b200 import get_model
    model = get_model("yolov10s", weights="PRETRAINED_COCO", class_names=names).to("cuda").eval()
    dets = model.decode_forward(model(x))          # x: [B,3,H,W] float, RGB, 0..255

Same surface as ``leanyolo`` (reference: leanyolo/__init__.py:1-7,
leanyolo/models/__init__.py:1-19); the compute is hand-written sm_100a CUDA behind
the C ABI in ``include/leanyolo_b200.h``.
"""
from .registry import get_model, get_model_weights, list_models
from .model import YOLOv10b, YOLOv10l, YOLOv10m, YOLOv10n, YOLOv10s, YOLOv10x
from .postprocess import decode_v10_official_topk, decode_v10_predictions

__all__ = [
    "get_model", "get_model_weights", "list_models",
    "YOLOv10n", "YOLOv10s", "YOLOv10m", "YOLOv10b", "YOLOv10l", "YOLOv10x",
    "decode_v10_official_topk", "decode_v10_predictions",
]
__version__ = "0.1.0"
