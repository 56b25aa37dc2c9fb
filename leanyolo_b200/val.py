"""Batched COCO-style evaluation loop on the GPU.

Mirrors the reference's ``tools/val.py:90-248`` (``validate_coco``: image -> letterbox -> model -> top-k or NMS
decode -> ``unletterbox_coords`` -> COCO-format result rows -> mAP) with the per-image Python loop replaced by
batches that stay on the device: ONE letterbox launch per batch of arbitrary image sizes, the plan forward, the GPU
decode with the unletterbox fused into its epilogue, one device->host copy of the fixed-shape detections per batch.

mAP: ``pycocotools`` is what the reference uses (``tools/val.py:236-247``); it is not in this image, so
``coco_bbox_map`` restates the published COCOeval bbox protocol (IoU .50:.05:.95, 101 recall points, maxDets 100,
crowd regions as ignore).  PARITY UNPINNED against pycocotools (absent here: no fixture could be generated); when
pycocotools is importable ``validate_coco`` uses it instead, exactly like the reference.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import postprocess as PP
from .preprocess import letterbox_batch, unletterbox_dets
from .variants import STRIDES


def _to_cuda_u8(img, dev) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(img)) if isinstance(img, np.ndarray) else img
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError("images must be uint8 HWC RGB")
    return t.to(dev, non_blocking=True).contiguous()


@torch.no_grad()
def detect_dataset(model, samples: Iterable[Tuple[int, "np.ndarray | torch.Tensor"]], *, imgsz: int = 640, decode: str = "topk",
                   conf: float = 0.25, iou: float = 0.65, max_dets: int = 300, batch_size: int = 64,
                   cat_ids: Optional[Sequence[int]] = None) -> List[dict]:
    """``samples``: (image_id, uint8 HWC RGB image) pairs of any sizes.  Returns COCO result rows
    ``{"image_id", "category_id", "bbox": [x, y, w, h], "score"}`` in each image's own pixel coordinates
    (tools/val.py:217-231).  decode="topk": one2one branch, official top-k (no threshold); "nms": one2many branch,
    ``decode_v10_predictions`` semantics (score > conf, greedy IoU NMS)."""
    if decode not in ("topk", "nms"):
        raise ValueError("decode must be 'topk' or 'nms'")
    dev = model.input_subtract.device
    if dev.type != "cuda":
        raise RuntimeError("leanyolo_b200 runs on CUDA (sm_100a) only: move the model to the GPU")
    nc = len(model.class_names)
    results: List[dict] = []

    def flush(ids: List[int], imgs: List[torch.Tensor]) -> None:
        batch, meta = letterbox_batch(imgs, imgsz)
        if decode == "topk":
            dets = model.detect(batch, max_det=max_dets, lb_meta=meta)
            counts = [dets.shape[1]] * len(ids)
        else:
            raw = model(batch)
            dets, cnt, _ = PP.nms_raw(raw, num_classes=nc, strides=STRIDES, conf_thresh=conf, iou_thresh=iou, max_det=max_dets)
            unletterbox_dets(dets, meta)
            counts = cnt.tolist()
        rows = dets.cpu().numpy()          # the one device->host copy of the batch
        for i, image_id in enumerate(ids):
            for x1, y1, x2, y2, score, cls in rows[i, :counts[i]]:
                c = int(cls)
                cat = c if cat_ids is None else (cat_ids[c] if c < len(cat_ids) else cat_ids[-1])
                results.append({"image_id": int(image_id), "category_id": int(cat),
                                "bbox": [float(x1), float(y1), float(x2 - x1), float(y2 - y1)], "score": float(score)})

    ids: List[int] = []
    imgs: List[torch.Tensor] = []
    for image_id, img in samples:
        ids.append(int(image_id))
        imgs.append(_to_cuda_u8(img, dev))
        if len(ids) == batch_size:
            flush(ids, imgs)
            ids, imgs = [], []
    if ids:
        flush(ids, imgs)
    return results


# ---------------------------------------------------------------------------------------------------- mAP
def _iou_xywh(d: np.ndarray, g: np.ndarray, crowd: np.ndarray) -> np.ndarray:
    """IoU of detections [n,4] and ground truths [m,4] (xywh); for crowd gts the union is the detection's area."""
    if len(d) == 0 or len(g) == 0:
        return np.zeros((len(d), len(g)))
    dx2, dy2, gx2, gy2 = d[:, 0] + d[:, 2], d[:, 1] + d[:, 3], g[:, 0] + g[:, 2], g[:, 1] + g[:, 3]
    iw = np.clip(np.minimum(dx2[:, None], gx2[None]) - np.maximum(d[:, None, 0], g[None, :, 0]), 0, None)
    ih = np.clip(np.minimum(dy2[:, None], gy2[None]) - np.maximum(d[:, None, 1], g[None, :, 1]), 0, None)
    inter = iw * ih
    da, ga = (d[:, 2] * d[:, 3])[:, None], (g[:, 2] * g[:, 3])[None]
    union = np.where(crowd[None], da, da + ga - inter)
    return inter / np.maximum(union, 1e-12)


def coco_bbox_map(gts: Sequence[dict], dts: Sequence[dict], max_dets: int = 100) -> Dict[str, float]:
    """COCO bbox AP restated from the published COCOeval protocol (see the module docstring: parity unpinned).
    ``gts``: {"image_id", "category_id", "bbox" [x,y,w,h], optional "iscrowd"}; ``dts``: + "score"."""
    thrs = np.linspace(0.5, 0.95, 10)
    rec_pts = np.linspace(0.0, 1.0, 101)
    cats = sorted({g["category_id"] for g in gts})
    by_g: Dict[Tuple[int, int], List[dict]] = {}
    by_d: Dict[Tuple[int, int], List[dict]] = {}
    for g in gts:
        by_g.setdefault((g["image_id"], g["category_id"]), []).append(g)
    for d in dts:
        by_d.setdefault((d["image_id"], d["category_id"]), []).append(d)
    images = sorted({k[0] for k in by_g} | {k[0] for k in by_d})
    aps = np.full((len(thrs), len(cats)), -1.0)
    for ci, cat in enumerate(cats):
        scores, matched, ignored, n_gt = [], [], [], 0
        for img in images:
            g = sorted(by_g.get((img, cat), []), key=lambda x: int(x.get("iscrowd", 0)))      # ignore regions last
            d = sorted(by_d.get((img, cat), []), key=lambda x: -x["score"])[:max_dets]
            if not g and not d:
                continue
            g_ign = np.array([bool(x.get("iscrowd", 0)) for x in g], dtype=bool)
            n_gt += int((~g_ign).sum())
            if not d:
                continue
            ious = _iou_xywh(np.array([x["bbox"] for x in d], dtype=np.float64).reshape(-1, 4),
                             np.array([x["bbox"] for x in g], dtype=np.float64).reshape(-1, 4), g_ign)
            dm = np.zeros((len(thrs), len(d)), dtype=bool)
            di = np.zeros((len(thrs), len(d)), dtype=bool)
            for ti, t in enumerate(thrs):
                gm = np.zeros(len(g), dtype=bool)
                for j in range(len(d)):
                    best, m = min(t, 1 - 1e-10), -1
                    for k in range(len(g)):
                        if gm[k] and not g_ign[k]:
                            continue
                        if m > -1 and not g_ign[m] and g_ign[k]:
                            break                      # a real match is already found: do not trade it for an ignore region
                        if ious[j, k] < best:
                            continue
                        best, m = ious[j, k], k
                    if m > -1:
                        dm[ti, j], di[ti, j], gm[m] = True, g_ign[m], True
            scores.append(np.array([x["score"] for x in d]))
            matched.append(dm)
            ignored.append(di)
        if n_gt == 0:
            continue
        if not scores:
            aps[:, ci] = 0.0
            continue
        sc = np.concatenate(scores)
        order = np.argsort(-sc, kind="mergesort")
        dm, di = np.concatenate(matched, 1)[:, order], np.concatenate(ignored, 1)[:, order]
        for ti in range(len(thrs)):
            tp = np.cumsum(dm[ti] & ~di[ti])
            fp = np.cumsum(~dm[ti] & ~di[ti])
            rc = tp / n_gt
            pr = tp / np.maximum(tp + fp, np.spacing(1))
            for i in range(len(pr) - 1, 0, -1):      # precision envelope
                pr[i - 1] = max(pr[i - 1], pr[i])
            idx = np.searchsorted(rc, rec_pts, side="left")
            q = np.zeros(len(rec_pts))
            ok = idx < len(pr)
            q[ok] = pr[idx[ok]]
            aps[ti, ci] = q.mean()
    valid = aps[:, (aps > -1).all(0)]
    if valid.size == 0:
        return {"mAP50-95": 0.0, "mAP50": 0.0, "mAP75": 0.0}
    return {"mAP50-95": float(valid.mean()), "mAP50": float(valid[0].mean()), "mAP75": float(valid[5].mean())}


@torch.no_grad()
def validate_coco(*, model=None, model_name: str = "yolov10s", weights: Optional[str] = "PRETRAINED_COCO", data_root: str = "data/coco",
                  imgsz: int = 640, conf: float = 0.25, iou: float = 0.65, decode: str = "topk", max_dets: int = 300,
                  device: str = "cuda", max_images: Optional[int] = None, save_json: Optional[str] = None,
                  images_dir: Optional[str] = None, ann_json: Optional[str] = None, batch_size: int = 64) -> Dict[str, float]:
    """``tools/val.py:90-248`` with the batched GPU loop.  Dataset layout as the reference resolves it
    (``<root>/images`` + ``<root>/annotations.json``, or explicit ``images_dir`` / ``ann_json``)."""
    import cv2
    from .registry import get_model
    root = Path(data_root)
    img_dir = Path(images_dir) if images_dir else root / "images"
    ann_p = Path(ann_json) if ann_json else root / "annotations.json"
    data = json.loads(ann_p.read_text(encoding="utf-8"))
    cats = sorted(data.get("categories", []), key=lambda c: c.get("id", 0))
    cat_ids = [int(c["id"]) for c in cats]
    names = [c.get("name", str(i)) for i, c in enumerate(cats)]
    if model is None:
        model = get_model(model_name, weights=weights, class_names=names, input_norm_subtract=[0.0, 0.0, 0.0],
                          input_norm_divide=[255.0, 255.0, 255.0]).to(device).eval()
    infos = sorted(data.get("images", []), key=lambda i: i["file_name"])
    if max_images is not None:
        infos = infos[:max_images]

    def samples():
        for info in infos:
            bgr = cv2.imread(str(img_dir / info["file_name"]), cv2.IMREAD_COLOR)
            if bgr is None:
                continue
            yield int(info["id"]), cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)

    results = detect_dataset(model, samples(), imgsz=imgsz, decode=decode, conf=conf, iou=iou, max_dets=max_dets,
                             batch_size=batch_size, cat_ids=cat_ids)
    if not results:
        return {"mAP50-95": 0.0}
    if save_json:
        Path(save_json).parent.mkdir(parents=True, exist_ok=True)
        Path(save_json).write_text(json.dumps(results))
    keep = {int(i["id"]) for i in infos}
    try:
        from pycocotools.coco import COCO
        from pycocotools.cocoeval import COCOeval
        coco = COCO(str(ann_p))
        ev = COCOeval(coco, coco.loadRes(results), iouType="bbox")
        ev.params.imgIds = sorted(keep)
        ev.evaluate(); ev.accumulate(); ev.summarize()
        return {"mAP50-95": float(ev.stats[0]), "mAP50": float(ev.stats[1]), "mAP75": float(ev.stats[2])}
    except ImportError:
        gts = [a for a in data.get("annotations", []) if int(a["image_id"]) in keep]
        return coco_bbox_map(gts, results)
