"""CPU oracle for the YOLOv10 inference hot path (TEST INFRASTRUCTURE ONLY).

This file is a from-scratch, *functional* restatement (plain torch fp32 on CPU,
driven purely by a reference-format ``state_dict``) of the algorithm that
jremillard/leanyolo runs for ``model(x)`` and the two decoders.  It is the
checker for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package ``leanyolo_b200`` never does.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the real reference
from ``/root/reference`` (possible only in the build container), runs it on
seeded weights/inputs and commits its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors
and against restated versions of the reference's own known-answer tests.

Reference lines each function follows (paths relative to the reference root):

* ``conv_bn_act``      leanyolo/models/yolov10/layers.py:51-88   (BN eps=1e-3)
* ``bottleneck``       layers.py:91-126
* ``cib``              layers.py:243-300   (RepVGGDW = SiLU(BN7(dw7)+BN3(dw3)))
* ``c2f_family``       layers.py:129-173, 303-335
* ``sppf``             layers.py:176-217
* ``attention``/``psa`` layers.py:338-425
* ``scdown``           layers.py:428-458
* ``backbone``         leanyolo/models/yolov10/backbone.py:88-106
* ``neck``             leanyolo/models/yolov10/neck.py:102-129
* ``head_branch``      leanyolo/models/yolov10/head.py:83-122
* ``forward``          leanyolo/models/yolov10/yolov10s.py:105-122 (same in n/m/b/l/x)
* ``decode_topk``      leanyolo/models/yolov10/postprocess.py:166-261,
                       leanyolo/utils/tal.py:10-46
* ``decode_nms``       postprocess.py:47-163 (DFL branch 103-139)
* ``box_iou``/``nms``  leanyolo/utils/box_ops.py:31-78
* ``nms_classwise``    leanyolo/models/yolov10/export.py:145-198 (semantics only)

Block structure (C2f vs C2fCIB, long-kernel branch, repeat counts, head widths)
is inferred from which keys exist in the ``state_dict`` — no variant table is
needed, which also makes the oracle independent of the product's tables.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-3  # layers.py:84


# --------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------
_CALIB = None  # set by calibrate_bn(): a torch.Generator while BN statistics are being re-estimated


def conv_bn_act(sd: SD, p: str, x: torch.Tensor, *, s: int = 1, act: bool = True) -> torch.Tensor:
    """``act(BN(conv(x)))``; kernel size and groups come from the weight shape."""
    w = sd[p + ".conv.weight"]
    k = w.shape[-1]
    g = x.shape[1] // w.shape[1]
    y = F.conv2d(x, w, None, s, k // 2, 1, g)
    if _CALIB is not None:
        # test-data preparation only: make this BN see roughly unit-variance input, with a
        # deliberate mismatch so that folding the statistics is still a non-trivial transform
        m, v = y.mean((0, 2, 3)), y.var((0, 2, 3), unbiased=False) + 1e-6
        sd[p + ".bn.running_mean"] = m + 0.2 * v.sqrt() * torch.randn(m.shape, generator=_CALIB)
        sd[p + ".bn.running_var"] = v * torch.empty(v.shape).uniform_(0.7, 1.4, generator=_CALIB)
    y = F.batch_norm(y, sd[p + ".bn.running_mean"], sd[p + ".bn.running_var"],
                     sd[p + ".bn.weight"], sd[p + ".bn.bias"], False, 0.0, BN_EPS)
    return F.silu(y) if act else y


def bottleneck(sd: SD, p: str, x: torch.Tensor, shortcut: bool) -> torch.Tensor:
    y = conv_bn_act(sd, p + ".cv2", conv_bn_act(sd, p + ".cv1", x))
    return x + y if shortcut and x.shape[1] == y.shape[1] else y


def cib(sd: SD, p: str, x: torch.Tensor, shortcut: bool) -> torch.Tensor:
    q = p + ".cv1"
    y = conv_bn_act(sd, q + ".0", x)
    y = conv_bn_act(sd, q + ".1", y)
    if (q + ".2.conv1.conv.weight") in sd:  # long-kernel RepVGGDW branch
        y = F.silu(conv_bn_act(sd, q + ".2.conv", y, act=False) + conv_bn_act(sd, q + ".2.conv1", y, act=False))
    else:
        y = conv_bn_act(sd, q + ".2", y)
    y = conv_bn_act(sd, q + ".3", y)
    y = conv_bn_act(sd, q + ".4", y)
    return x + y if shortcut and x.shape[1] == y.shape[1] else y


def _count(sd: SD, prefix: str) -> int:
    n = 0
    while any(k.startswith(f"{prefix}.{n}.") for k in sd):
        n += 1
    return n


def c2f_family(sd: SD, p: str, x: torch.Tensor, *, c2f_shortcut: bool) -> torch.Tensor:
    """C2f or C2fCIB, decided by the inner block's keys.  C2fCIB always uses the
    residual (backbone.py:75,80; neck.py:85,93,98); C2f uses ``c2f_shortcut``."""
    y = conv_bn_act(sd, p + ".cv1", x)
    y1, y2 = y.chunk(2, 1)
    parts = [y1, y2]
    for i in range(_count(sd, p + ".m")):
        q = f"{p}.m.{i}"
        if (q + ".cv1.0.conv.weight") in sd:
            y2 = cib(sd, q, y2, True)
        else:
            y2 = bottleneck(sd, q, y2, c2f_shortcut)
        parts.append(y2)
    return conv_bn_act(sd, p + ".cv2", torch.cat(parts, 1))


def sppf(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    x = conv_bn_act(sd, p + ".cv1", x)
    y1 = F.max_pool2d(x, 5, 1, 2)
    y2 = F.max_pool2d(y1, 5, 1, 2)
    y3 = F.max_pool2d(y2, 5, 1, 2)
    return conv_bn_act(sd, p + ".cv2", torch.cat([x, y1, y2, y3], 1))


def attention(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    b, c, h, w = x.shape
    n = h * w
    nh = max(1, c // 64)
    hd = c // nh
    kd = int(hd * 0.5)
    qkv = conv_bn_act(sd, p + ".qkv", x, act=False).view(b, nh, 2 * kd + hd, n)
    q, k, v = qkv.split([kd, kd, hd], dim=2)
    att = ((q.transpose(-2, -1) @ k) * (kd ** -0.5)).softmax(dim=-1)
    o = (v @ att.transpose(-2, -1)).view(b, c, h, w)
    o = o + conv_bn_act(sd, p + ".pe", v.reshape(b, c, h, w), act=False)
    return conv_bn_act(sd, p + ".proj", o, act=False)


def psa(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    y = conv_bn_act(sd, p + ".cv1", x)
    c = y.shape[1] // 2
    a, b = y.split((c, c), dim=1)
    b = b + attention(sd, p + ".attn", b)
    b = b + conv_bn_act(sd, p + ".ffn.1", conv_bn_act(sd, p + ".ffn.0", b), act=False)
    return conv_bn_act(sd, p + ".cv2", torch.cat((a, b), 1))


def scdown(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return conv_bn_act(sd, p + ".cv2", conv_bn_act(sd, p + ".cv1", x), s=2, act=False)


# --------------------------------------------------------------------------
# graph
# --------------------------------------------------------------------------
def backbone(sd: SD, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    p = "backbone."
    x = conv_bn_act(sd, p + "cv0", x, s=2)
    x = conv_bn_act(sd, p + "cv1", x, s=2)
    x = c2f_family(sd, p + "c2", x, c2f_shortcut=True)
    x = conv_bn_act(sd, p + "cv3", x, s=2)
    c3 = c2f_family(sd, p + "c4", x, c2f_shortcut=True)
    x = scdown(sd, p + "sc5", c3)
    c4 = c2f_family(sd, p + "c6", x, c2f_shortcut=True)
    x = scdown(sd, p + "sc7", c4)
    x = c2f_family(sd, p + "c8", x, c2f_shortcut=True)
    x = sppf(sd, p + "sppf9", x)
    c5 = psa(sd, p + "psa10", x)
    return c3, c4, c5


def _up2(x: torch.Tensor) -> torch.Tensor:
    return F.interpolate(x, scale_factor=2.0, mode="nearest")


def neck(sd: SD, c3, c4, c5):
    p = "neck."
    p4 = c2f_family(sd, p + "p5_p4_c2f", torch.cat([_up2(c5), c4], 1), c2f_shortcut=False)
    p3 = c2f_family(sd, p + "p4_p3_c2f", torch.cat([_up2(p4), c3], 1), c2f_shortcut=False)
    d3 = conv_bn_act(sd, p + "p3_down", p3, s=2)
    p4 = c2f_family(sd, p + "p3_p4_c2f", torch.cat([d3, p4], 1), c2f_shortcut=False)
    d4 = scdown(sd, p + "p4_down", p4)
    p5 = c2f_family(sd, p + "p4_p5_c2f", torch.cat([d4, c5], 1), c2f_shortcut=False)
    return p3, p4, p5


def head_branch(sd: SD, feats: Sequence[torch.Tensor], reg: str, cls: str) -> List[torch.Tensor]:
    """``reg``/``cls`` are 'head.cv2'/'head.cv3' or the one2one twins."""
    out = []
    for i, f in enumerate(feats):
        r = conv_bn_act(sd, f"{reg}.{i}.1", conv_bn_act(sd, f"{reg}.{i}.0", f))
        c = conv_bn_act(sd, f"{cls}.{i}.0.1", conv_bn_act(sd, f"{cls}.{i}.0.0", f))
        c = conv_bn_act(sd, f"{cls}.{i}.1.1", conv_bn_act(sd, f"{cls}.{i}.1.0", c))
        if _CALIB is not None:
            # test-data preparation only: rescale the two final 1x1 convs so that logits are
            # O(1) (DFL bins spread, class scores mostly low with a tail) instead of saturated
            for key, t, shift in ((f"{reg}.{i}.2", r, 0.0), (f"{cls}.{i}.2", c, -3.0)):
                std = F.conv2d(t, sd[key + ".weight"]).std().clamp(min=1e-6)
                sd[key + ".weight"] = sd[key + ".weight"] * (1.5 / std)
                sd[key + ".bias"] = sd[key + ".bias"] + shift
        r = F.conv2d(r, sd[f"{reg}.{i}.2.weight"], sd[f"{reg}.{i}.2.bias"])
        c = F.conv2d(c, sd[f"{cls}.{i}.2.weight"], sd[f"{cls}.{i}.2.bias"])
        out.append(torch.cat((r, c), 1))
    return out


@torch.no_grad()
def forward(sd: SD, x: torch.Tensor, *, taps: dict | None = None) -> Dict[str, List[torch.Tensor]]:
    """Eval forward: returns {'one2many': [3 tensors], 'one2one': [3 tensors]}.

    ``taps`` (optional dict) receives c3,c4,c5,p3,p4,p5 — the same taps the
    reference's fidelity suite compares."""
    sub, div = sd["input_subtract"], sd["input_divide"]
    skip_sub, skip_div = bool((sub == 0).all()), bool((div == 1).all())
    if not (skip_sub and skip_div):
        x = x.float()
    if not skip_sub:
        x = x - sub
    if not skip_div:
        x = x / div
    c3, c4, c5 = backbone(sd, x)
    p3, p4, p5 = neck(sd, c3, c4, c5)
    if taps is not None:
        taps.update(c3=c3, c4=c4, c5=c5, p3=p3, p4=p4, p5=p5)
    return {
        "one2many": head_branch(sd, (p3, p4, p5), "head.cv2", "head.cv3"),
        "one2one": head_branch(sd, (p3, p4, p5), "head.one2one_cv2", "head.one2one_cv3"),
    }


def calibrate_bn(sd: SD, x: torch.Tensor, seed: int = 0) -> SD:
    """Return a copy of ``sd`` whose BN running statistics are re-estimated layer by
    layer on ``x`` (then perturbed), so that synthetic weights keep activations O(1) at
    any depth and outputs depend strongly on the input.  Test-data preparation, not part
    of the reference algorithm."""
    global _CALIB
    out = {k: v.clone() for k, v in sd.items()}
    _CALIB = torch.Generator().manual_seed(seed)
    try:
        forward(out, x)
    finally:
        _CALIB = None
    return out


# --------------------------------------------------------------------------
# decode
# --------------------------------------------------------------------------
def _dfl_boxes_scores(preds: Sequence[torch.Tensor], nc: int, strides: Sequence[int]):
    """Per-level DFL expectation + anchor decode.  Returns boxes [B,A,4] (xyxy,
    pixels) and sigmoid scores [B,A,nc], anchors ordered level-major, row-major."""
    boxes, scores = [], []
    for p, s in zip(preds, strides):
        b, c, h, w = p.shape
        reg_max = (c - nc) // 4
        assert 4 * reg_max + nc == c
        p = p.reshape(b, c, h * w)
        bins = torch.arange(reg_max, dtype=p.dtype)
        dist = (p[:, : 4 * reg_max].reshape(b, 4, reg_max, h * w).softmax(2) * bins.view(1, 1, -1, 1)).sum(2)
        ax = (torch.arange(w, dtype=p.dtype) + 0.5).repeat(h)
        ay = (torch.arange(h, dtype=p.dtype) + 0.5).repeat_interleave(w)
        x1y1 = torch.stack((ax, ay), 0)[None] - dist[:, :2]
        x2y2 = torch.stack((ax, ay), 0)[None] + dist[:, 2:]
        boxes.append((torch.cat((x1y1, x2y2), 1) * float(s)).permute(0, 2, 1))
        scores.append(p[:, 4 * reg_max:].sigmoid().permute(0, 2, 1))
    return torch.cat(boxes, 1), torch.cat(scores, 1)


def topk_canonical(v: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k of each row with the canonical tie rule (value desc, index asc).
    ``torch.topk`` tie order is unspecified; a stable descending sort is the
    deterministic restatement both sides are compared under."""
    vals, idx = torch.sort(v, dim=1, descending=True, stable=True)
    return vals[:, :k], idx[:, :k]


@torch.no_grad()
def decode_topk(preds: Sequence[torch.Tensor], *, num_classes: int,
                strides: Sequence[int] = (8, 16, 32), max_det: int = 300,
                return_indices: bool = False):
    """Two-stage top-k, no NMS (postprocess.py:166-261)."""
    boxes, scores = _dfl_boxes_scores(preds, num_classes, strides)
    B, A, nc = scores.shape
    k = min(max_det, A)
    _, top_anchor = topk_canonical(scores.amax(-1), k)                 # stage 1
    sel = scores.gather(1, top_anchor[..., None].expand(B, k, nc)).reshape(B, k * nc)
    vals, flat = topk_canonical(sel, k)                                # stage 2
    anchor = top_anchor.gather(1, flat // nc)
    cls = flat % nc
    out = torch.cat((boxes.gather(1, anchor[..., None].expand(B, k, 4)), vals[..., None], cls[..., None].to(boxes.dtype)), -1)
    dets = [[out[i]] for i in range(B)]
    return (dets, anchor, cls) if return_indices else dets


def box_iou(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """box_ops.py:31-46 — exact operation order matters for keep-set parity."""
    area_a = (a[:, 2] - a[:, 0]).clamp(min=0) * (a[:, 3] - a[:, 1]).clamp(min=0)
    area_b = (b[:, 2] - b[:, 0]).clamp(min=0) * (b[:, 3] - b[:, 1]).clamp(min=0)
    lt = torch.max(a[:, None, :2], b[:, :2])
    rb = torch.min(a[:, None, 2:], b[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b - inter + 1e-9)


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_thresh: float, max_keep: int | None = None) -> torch.Tensor:
    """Greedy class-agnostic NMS (box_ops.py:49-78): visit by (score desc, index
    asc), drop j when IoU(kept, j) > thr.  Vectorised as one IoU matrix plus a
    sequential scan; ``max_keep`` only truncates (the reference truncates after
    running to exhaustion, postprocess.py:158-159 — same result)."""
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros((0,), dtype=torch.long)
    order = torch.sort(scores, descending=True, stable=True)[1]
    bs = boxes[order]
    alive = torch.ones(n, dtype=torch.bool)
    keep: List[int] = []
    for i in range(n):
        if not alive[i]:
            continue
        keep.append(i)
        if max_keep is not None and len(keep) >= max_keep:
            break
        if i + 1 < n:
            iou = box_iou(bs[i:i + 1], bs[i + 1:])[0]
            alive[i + 1:] &= iou <= iou_thresh
    return order[torch.tensor(keep, dtype=torch.long)]


@torch.no_grad()
def decode_nms(preds: Sequence[torch.Tensor], *, num_classes: int, strides: Sequence[int] = (8, 16, 32),
               conf_thresh: float = 0.25, iou_thresh: float = 0.45, max_det: int = 300,
               classwise: bool = False, return_indices: bool = False):
    """DFL layout of ``decode_v10_predictions`` (postprocess.py:103-161): one
    (max-class score, label) candidate per anchor, ``score > conf`` (strict),
    class-agnostic greedy NMS, first ``max_det`` survivors.  ``classwise=True``
    is the north-star variant: suppression only between equal labels
    (export.py:165-176 offset trick ≡ per-class NMS)."""
    boxes, scores = _dfl_boxes_scores(preds, num_classes, strides)
    best, label = scores.max(-1)
    dets, kept = [], []
    for i in range(boxes.shape[0]):
        cand = torch.nonzero(best[i] > conf_thresh).flatten()
        if cand.numel() == 0:
            dets.append([torch.empty((0, 6))])
            kept.append(cand)
            continue
        bi, si, li = boxes[i, cand], best[i, cand], label[i, cand]
        if classwise:
            k = nms_classwise(bi, si, li, iou_thresh, max_det)
        else:
            k = nms(bi, si, iou_thresh, max_det)[:max_det]
        dets.append([torch.cat((bi[k], si[k, None], li[k, None].to(bi.dtype)), 1)])
        kept.append(cand[k])
    return (dets, kept) if return_indices else dets


def nms_classwise(boxes, scores, labels, iou_thresh: float, max_keep: int | None = None) -> torch.Tensor:
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros((0,), dtype=torch.long)
    order = torch.sort(scores, descending=True, stable=True)[1]
    bs, ls = boxes[order], labels[order]
    alive = torch.ones(n, dtype=torch.bool)
    keep: List[int] = []
    for i in range(n):
        if not alive[i]:
            continue
        keep.append(i)
        if max_keep is not None and len(keep) >= max_keep:
            break
        if i + 1 < n:
            iou = box_iou(bs[i:i + 1], bs[i + 1:])[0]
            alive[i + 1:] &= ~((iou > iou_thresh) & (ls[i + 1:] == ls[i]))
    return order[torch.tensor(keep, dtype=torch.long)]


def nms_torchvision(boxes: torch.Tensor, scores: torch.Tensor, iou_thresh: float) -> torch.Tensor:
    """``torchvision.ops.nms`` restated (third-party dependency of export.py:30-33,182; torchvision is not under the
    reference tree, unpinned in requirements.txt, 0.26.0 in this image; CPU kernel ``nms_kernel_impl``): areas
    ``(x2-x1)*(y2-y1)`` WITHOUT clamping, candidates visited by a stable descending sort of the scores,
    ``ovr = inter / (area_i + area_j - inter)`` in fp32 (no epsilon) and ``ovr > thr`` with the threshold held
    as a double.  Returns kept indices in visiting order."""
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros((0,), dtype=torch.long)
    order = torch.sort(scores, descending=True, stable=True)[1]
    b = boxes[order]
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    alive = torch.ones(n, dtype=torch.bool)
    keep: List[int] = []
    for i in range(n):
        if not alive[i]:
            continue
        keep.append(i)
        if i + 1 < n:
            w = (torch.min(b[i, 2], b[i + 1:, 2]) - torch.max(b[i, 0], b[i + 1:, 0])).clamp(min=0)
            h = (torch.min(b[i, 3], b[i + 1:, 3]) - torch.max(b[i, 1], b[i + 1:, 1])).clamp(min=0)
            inter = w * h
            ovr = inter / (area[i] + area[i + 1:] - inter)
            alive[i + 1:] &= ~(ovr.double() > float(iou_thresh))
    return order[torch.tensor(keep, dtype=torch.long)]


@torch.no_grad()
def decode_export(preds: Sequence[torch.Tensor], *, num_classes: int, strides: Sequence[int] = (8, 16, 32),
                  imgsz: int = 640, max_dets: int = 300, conf: float = 0.25, nms: bool = False, iou: float = 0.45,
                  pre_topk: int = 1000, img0: int = 0):
    """``YOLOv10ONNXExport.forward`` after the model call (export.py:97-198): fixed-shape detections
    ``[B, k, 6]`` + ``num_dets [B]`` (int64).

    nms=False (:126-144): top-k anchors by best class score (scores below ``conf`` masked to -1), argmax class,
    boxes clamped to the image, ``num_dets = #(score >= conf)``; rows past ``num_dets`` are whatever anchors the
    tie order of ``torch.topk`` at -1 picks (canonical rule here: index ascending).
    nms=True (:145-198): the ``pre_topk`` best (anchor, class) pairs, boxes offset by ``(image*C + class) * 10*imgsz``
    IN FP32 (the rounding of that addition is part of the reference's result), ONE torchvision NMS over all
    images, then the first ``max_dets`` survivors per image; entries below ``conf`` are zeroed.  ``img0`` = global
    index of image 0 of this batch (the offset arithmetic depends on it)."""
    boxes, scores = _dfl_boxes_scores(preds, num_classes, strides)
    B, A, C = scores.shape
    H = W = float(imgsz)
    lo = torch.tensor(0.0)

    def clamp_boxes(bx):
        bx = bx.clone()
        bx[..., 0].clamp_(0, W); bx[..., 2].clamp_(0, W); bx[..., 1].clamp_(0, H); bx[..., 3].clamp_(0, H)
        return bx

    conf_t = torch.tensor(conf, dtype=boxes.dtype)
    if not nms:
        best, cls = scores.max(dim=2)
        masked = torch.where(best >= conf_t, best, torch.full_like(best, -1.0))
        k = min(max_dets, A)
        _, idx = topk_canonical(masked, k)
        sel_b = clamp_boxes(boxes.gather(1, idx[..., None].expand(B, k, 4)))
        sel_s = best.gather(1, idx).clamp(min=0.0)
        sel_c = cls.gather(1, idx).to(boxes.dtype)
        return torch.cat((sel_b, sel_s[..., None], sel_c[..., None]), -1), (sel_s >= conf_t).sum(1).to(torch.int64)
    flat = scores.reshape(B, A * C)
    k_pre = min(pre_topk, A * C)
    vals, pidx = topk_canonical(flat, k_pre)
    anc, cls = pidx // C, pidx % C
    cand = boxes.gather(1, anc[..., None].expand(B, k_pre, 4))
    img = (torch.arange(B) + img0).view(B, 1).to(boxes.dtype)
    off = ((img * float(C) + cls.to(boxes.dtype)) * float(max(H, W) * 10.0)).unsqueeze(-1)
    cand_off = cand + off                                        # fp32 addition: rounds the coordinates
    keep = nms_torchvision(cand_off.reshape(B * k_pre, 4), vals.reshape(B * k_pre), iou)
    kmax = min(max_dets, k_pre)
    det = torch.zeros(B, kmax, 6)
    num = torch.zeros(B, dtype=torch.int64)
    cand_c = clamp_boxes(cand)
    for b in range(B):
        kb = keep[(keep >= b * k_pre) & (keep < (b + 1) * k_pre)] - b * k_pre      # already score-descending
        kb = kb[:kmax]
        kb = kb[vals[b, kb] >= conf_t]
        n = kb.numel()
        det[b, :n, :4] = cand_c[b, kb]
        det[b, :n, 4] = vals[b, kb]
        det[b, :n, 5] = cls[b, kb].to(boxes.dtype)
        num[b] = n
    return det, num


@torch.no_grad()
def decode_forward(sd: SD, x: torch.Tensor, *, max_det: int = 300):
    """``model.decode_forward(model(x))`` (yolov10s.py:124-144): top-k on one2one."""
    nc = sd["head.cv3.0.2.weight"].shape[0]
    return decode_topk(forward(sd, x)["one2one"], num_classes=nc, max_det=max_det)
