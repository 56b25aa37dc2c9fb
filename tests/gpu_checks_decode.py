"""GPU parity checks of the detection tail against the CPU oracle / committed goldens."""
from __future__ import annotations

import os

import torch

from leanyolo_b200 import postprocess as PP
from leanyolo_b200.synth import synth_head_logits
from oracle import yolov10_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"
HW640 = [(80, 80), (40, 40), (20, 20)]


def _solid_rows(scores: torch.Tensor, gap_ulps: float = 16.0) -> torch.Tensor:
    s = scores.double()
    eps = torch.finfo(torch.float32).eps * s.abs().clamp(min=1e-30)
    g = (s[:-1] - s[1:]) / eps[:-1]
    big = torch.tensor([1e9], dtype=torch.float64)
    return (torch.cat((big, g)) > gap_ulps) & (torch.cat((g, big)) > gap_ulps)


def _compare_topk(out, anchor, cls, ref_rows, ref_anchor=None, ref_cls=None):
    out = out.cpu()
    assert out.shape == ref_rows.shape, (out.shape, ref_rows.shape)
    assert torch.allclose(out[:, 4], ref_rows[:, 4], atol=2e-6, rtol=0), "scores differ"
    assert bool((out[:-1, 4] >= out[1:, 4]).all()), "rows not score-descending"
    solid = _solid_rows(ref_rows[:, 4])
    assert torch.equal(out[solid, 5], ref_rows[solid, 5]), "class indices differ on tie-free rows"
    assert torch.allclose(out[solid, :4], ref_rows[solid, :4], atol=2e-3, rtol=1e-5), "boxes differ"
    if ref_anchor is not None:
        assert torch.equal(anchor.cpu().long()[solid], ref_anchor[solid]), "anchor indices differ on tie-free rows"
        assert torch.equal(cls.cpu().long()[solid], ref_cls[solid])
    return float(solid.float().mean())


def check_topk_golden():
    g = torch.load(os.path.join(G, "decode_topk.pt"))
    logits = [t.to(DEV) for t in synth_head_logits(2, g["nc"], g["hw"], seed=g["seed"])]
    out, anchor, cls = PP.topk_raw(logits, num_classes=g["nc"])
    fr = [_compare_topk(out[i], anchor[i], cls[i], g["out"][i]) for i in range(2)]
    return {"solid_fraction": min(fr)}


def check_topk_vs_oracle(B=4, seed=3, hw=None, nc=80, reg_max=16, max_det=300, strides=(8, 16, 32)):
    hw = hw or HW640
    logits = synth_head_logits(B, nc, hw, reg_max=reg_max, seed=seed)
    dets, ra, rc = O.decode_topk(logits, num_classes=nc, strides=strides, max_det=max_det, return_indices=True)
    out, anchor, cls = PP.topk_raw([t.to(DEV) for t in logits], num_classes=nc, strides=strides, max_det=max_det)
    fr = [_compare_topk(out[i], anchor[i], cls[i], dets[i][0], ra[i], rc[i]) for i in range(B)]
    # the list-of-lists drop-in wrapper returns the same rows
    wrapped = PP.decode_v10_official_topk([t.to(DEV) for t in logits], num_classes=nc, strides=strides, max_det=max_det)
    assert len(wrapped) == B and torch.equal(wrapped[0][0], out[0])
    return {"solid_fraction": min(fr)}


def _boxes(n, seed):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(n, 2, generator=g) * 600
    wh = torch.rand(n, 2, generator=g) * 120 + 4
    boxes = torch.cat((xy, xy + wh), 1)
    scores = (torch.randperm(n, generator=g).float() + 0.5) / n
    labels = torch.randint(0, 5, (n,), generator=g)
    return boxes, scores, labels


def check_nms_exact(n=3000, thr=0.5, classwise=False, seed=41):
    boxes, scores, labels = _boxes(n, seed)
    if classwise:
        ref = O.nms_classwise(boxes, scores, labels, thr, 1024)
    else:
        ref = O.nms(boxes, scores, thr, 1024)
    got = PP.nms(boxes.to(DEV), scores.to(DEV), thr, labels=labels.to(DEV) if classwise else None, max_keep=1024).cpu()
    assert torch.equal(got, ref[:1024]), f"keep-set differs: {got[:8].tolist()} vs {ref[:8].tolist()} (len {len(got)} vs {len(ref)})"
    return {"kept": int(got.numel())}


def check_nms_golden():
    g = torch.load(os.path.join(G, "nms.pt"))
    boxes, scores, _ = _boxes(g["n"], g["seed"])
    kept = {}
    for thr, keep in g["keep"].items():
        got = PP.nms(boxes.to(DEV), scores.to(DEV), float(thr), max_keep=1024).cpu()
        assert torch.equal(got, keep[:1024]), f"thr {thr}"
        kept[thr] = int(got.numel())
    return kept


def _canon(d):
    if d.numel() == 0:
        return d
    idx = sorted(range(d.shape[0]), key=lambda i: (-float(d[i, 4]), float(d[i, 0]), float(d[i, 1])))
    return d[idx]


def flip_count(a, b):
    """Rows of `a` without a counterpart in `b` plus rows of `b` without one in `a` (counterpart: same label,
    |score difference| <= 1e-5, box corners within 1e-2 px).  Both sides decode the same logits; the GPU's DFL
    (CUDA expf) and torch's CPU softmax differ by ulps, so a greedy-NMS decision whose IoU sits within an ulp of
    the threshold may flip, and one flip can change a few later rows."""
    if a.numel() == 0 or b.numel() == 0:
        return int(a.shape[0] + b.shape[0])
    same = ((a[:, None, :4] - b[None, :, :4]).abs().amax(-1) < 1e-2) & ((a[:, None, 4] - b[None, :, 4]).abs() <= 1e-5) & \
           (a[:, None, 5] == b[None, :, 5])
    return int((~same.any(1)).sum() + (~same.any(0)).sum())


def nms_on_oracle_candidates(logits, conf, iou, max_det, classwise=False):
    """The exact half of the end-to-end claim: the oracle decodes the candidates (boxes, best score, label) and BOTH
    sides run NMS on those identical tensors: keep-sets must be equal, element for element."""
    boxes, scores = O._dfl_boxes_scores(logits, 80, (8, 16, 32))
    best, label = scores.max(-1)
    kept = 0
    for i in range(boxes.shape[0]):
        cand = torch.nonzero(best[i] > conf).flatten()
        if cand.numel() == 0:
            continue
        bi, si, li = boxes[i, cand].contiguous(), best[i, cand].contiguous(), label[i, cand].contiguous()
        ref = O.nms_classwise(bi, si, li, iou, max_det) if classwise else O.nms(bi, si, iou, max_det)[:max_det]
        got = PP.nms(bi.to(DEV), si.to(DEV), iou, labels=li.to(DEV) if classwise else None, max_keep=max_det).cpu()
        assert torch.equal(got, ref[:max_det]), f"image {i}: keep-set on identical candidates differs ({len(got)} vs {len(ref)})"
        kept += int(got.numel())
    return kept


def check_decode_nms(conf, iou, cls_mean, seed, B=2, classwise=False):
    logits = synth_head_logits(B, 80, HW640, seed=seed, cls_mean=cls_mean)
    kept_exact = nms_on_oracle_candidates(logits, conf, iou, 300, classwise)
    ref = O.decode_nms(logits, num_classes=80, conf_thresh=conf, iou_thresh=iou, max_det=300, classwise=classwise)
    got = PP.decode_v10_predictions([t.to(DEV) for t in logits], num_classes=80, conf_thresh=conf, iou_thresh=iou,
                                    max_det=300, classwise=classwise)
    rows = flips = 0
    for r, g_ in zip(ref, got):
        a, b = _canon(r[0]), _canon(g_[0].cpu())
        assert a.shape == b.shape, (a.shape, b.shape)
        rows += a.shape[0]
        flips += flip_count(a, b)
    # end to end the boxes are decoded on each side (ulp-level differences): report the flips, bound them at 1 %
    assert flips <= max(2, rows // 100), f"{flips} of {rows} rows differ end to end"
    return {"rows": rows, "flips_end_to_end": flips, "kept_exact_on_identical_candidates": kept_exact}


def check_decode_nms_direct():
    """Legacy direct-offset layout + clamp (leanyolo/tests/test_postprocess_v10_ext.py:56-98)."""
    nc = 3
    p = torch.zeros(1, 4 + nc, 2, 2)
    p[:, 4:] = -10.0
    p[0, 4, 0, 0], p[0, 4, 0, 1] = 8.0, 6.0
    p[0, 2, 0, 0] = p[0, 3, 0, 0] = p[0, 2, 0, 1] = p[0, 3, 0, 1] = 3.0
    neg = torch.zeros_like(p)
    neg[:, 4:] = -10.0
    out = PP.decode_v10_predictions([p.to(DEV), neg.to(DEV), neg.to(DEV)], num_classes=nc, strides=(8, 8, 8), conf_thresh=0.25,
                                    iou_thresh=0.5, max_det=10, img_size=(64, 64))[0][0].cpu()
    assert out.shape == (1, 6)
    assert bool(((out[0, :4] >= 0) & (out[0, :4] <= 64)).all())
    p2 = torch.zeros(2, 6, 1, 1)
    p2[0, 5, 0, 0], p2[1, 5, 0, 0] = 2.0, -5.0
    z = torch.zeros_like(p2)
    d = PP.decode_v10_predictions([p2.to(DEV), z.to(DEV), z.to(DEV)], num_classes=2, strides=(16, 16, 16), conf_thresh=0.5,
                                  iou_thresh=0.5, max_det=5, img_size=(64, 64))
    assert d[0][0].shape[0] >= 1 and d[1][0].shape[0] == 0
    return {}


# ------------------------------------------------------------------ export-style outputs (export.py:126-198)
def _compare_export(det, num, rdet, rnum, nms):
    det, num = det.cpu(), num.cpu()
    assert det.shape == rdet.shape, (det.shape, rdet.shape)
    assert torch.equal(num.to(torch.int64), rnum.to(torch.int64)), (num.tolist(), rnum.tolist())
    frac = 1.0
    for b in range(det.shape[0]):
        n = int(rnum[b])
        o, r = det[b, :n], rdet[b, :n]
        assert torch.allclose(o[:, 4], r[:, 4], atol=2e-6, rtol=0), "scores differ"
        if n > 2:
            solid = _solid_rows(r[:, 4])
            frac = min(frac, float(solid.float().mean()))
            assert torch.equal(o[solid, 5], r[solid, 5]), "class indices differ on tie-free rows"
            assert torch.allclose(o[solid, :4], r[solid, :4], atol=2e-3, rtol=1e-5), "boxes differ"
        if nms and n < det.shape[1]:
            assert float(det[b, n:].abs().max()) == 0.0, "rows past num_dets must be zero"
    return frac


def check_export_golden():
    """CUDA export decode against outputs of the reference's own YOLOv10ONNXExport (real torchvision NMS)."""
    g = torch.load(os.path.join(G, "decode_export.pt"))
    worst = 1.0
    for tag, c in g["cases"].items():
        lg = [t.to(DEV) for t in synth_head_logits(g["B"], g["nc"], g["hw"], seed=c["seed"], cls_mean=c["cls_mean"])]
        kw = c["kw"]
        det, num = PP.export_decode(lg, num_classes=g["nc"], imgsz=640, max_dets=kw["max_dets"], conf=kw["conf"], nms=kw["nms"],
                                    iou=kw.get("iou", 0.45), pre_topk=kw.get("pre_topk", 1000))
        worst = min(worst, _compare_export(det, num, c["det"], c["num"], kw["nms"]))
    return {"solid_fraction": worst}


def check_export_vs_oracle(B=3, seed=51, nms=True, conf=0.2, iou=0.6, max_dets=120, pre_topk=700, img0=37, hw=None, nc=80,
                           cls_mean=-2.5, strides=(8, 16, 32), imgsz=640):
    """... and against the oracle, incl. a non-zero first-image index (the reference's fp32 class-offset arithmetic
    depends on the global image index) and small pyramids where A < pre_topk."""
    hw = hw or HW640
    lg = synth_head_logits(B, nc, hw, seed=seed, cls_mean=cls_mean)
    rdet, rnum = O.decode_export(lg, num_classes=nc, strides=strides, imgsz=imgsz, max_dets=max_dets, conf=conf, nms=nms, iou=iou,
                                 pre_topk=pre_topk, img0=img0)
    det, num = PP.export_decode([t.to(DEV) for t in lg], num_classes=nc, strides=strides, imgsz=imgsz, max_dets=max_dets, conf=conf,
                                nms=nms, iou=iou, pre_topk=pre_topk, img0=img0)
    return {"solid_fraction": _compare_export(det, num, rdet, rnum, nms), "num": num.tolist()}
