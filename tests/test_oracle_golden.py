"""Pin the CPU oracle (oracle/yolov10_oracle.py) to the reference.

1. against outputs of the real reference committed under tests/golden/ (made by
   oracle/make_golden.py in the build container);
2. against restated versions of the reference's own known-answer tests
   (leanyolo/tests/test_postprocess_v10_ext.py, test_box_ops_extra.py,
   test_postprocess.py, test_head_v10.py:41-51).
"""
import json
import os

import pytest
import torch

from leanyolo_b200 import get_model
from leanyolo_b200.synth import synth_head_logits, synth_images, synth_state_dict
from oracle import yolov10_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")
VARIANTS = ["yolov10n", "yolov10s", "yolov10m", "yolov10b", "yolov10l", "yolov10x"]
NAMES = [f"class{i}" for i in range(80)]


def _relmax(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))


@pytest.mark.parametrize("name", VARIANTS)
def test_forward_matches_reference_golden(name):
    g = torch.load(os.path.join(G, f"forward_{name}.pt"))
    model = get_model(name, weights=None, class_names=NAMES)
    sd = synth_state_dict(model.state_dict(), seed=g["seed_weights"], gain=g["gain"])
    x = synth_images(1, g["hw"], g["hw"], seed=g["seed_input"])
    taps = {}
    out = O.forward(sd, x, taps=taps)
    # fp32 CPU conv kernels may differ between hosts (ISA-dependent blocking): 1e-4 of the
    # tensor's max, the reference's own fidelity rubric (tests/fidelity/rubric.py:28-31)
    for k, ref in g["taps"].items():
        assert _relmax(taps[k], ref) < 1e-4, k
    for br in ("one2many", "one2one"):
        for i in range(3):
            assert out[br][i].shape == g[br][i].shape
            assert _relmax(out[br][i], g[br][i]) < 1e-4, (br, i)
    scores = O.decode_forward(sd, x)[0][0][:, 4]
    assert torch.allclose(scores, g["dets_scores"], atol=1e-5)


def test_forward_640_matches_reference_golden_samples():
    """Headline resolution: the oracle against the reference's yolov10s @640x640 head tensors (2048 sampled positions
    and the double-precision checksums of each of the six tensors; oracle/make_golden.py::make_640)."""
    g = torch.load(os.path.join(G, "forward640_yolov10s.pt"))
    model = get_model("yolov10s", weights=None, class_names=NAMES)
    sd = synth_state_dict(model.state_dict(), seed=g["seed_weights"], gain=g["gain"])
    x = synth_images(1, g["hw"], g["hw"], seed=g["seed_input"])
    out = O.forward(sd, x)
    for br in ("one2many", "one2one"):
        for i in range(3):
            ref = g[f"{br}{i}"]
            t = out[br][i]
            assert list(t.shape) == ref["shape"]
            err = float((t.flatten()[ref["idx"]] - ref["val"]).abs().max()) / ref["abs_max"]
            assert err < 1e-4, (br, i, err)
            assert abs(float(t.double().abs().sum()) - ref["abs_sum"]) < 1e-4 * ref["abs_sum"], (br, i)


def test_state_dict_keys_match_reference_order_and_shapes():
    ref = json.load(open(os.path.join(G, "state_keys.json")))
    for name in VARIANTS:
        sd = get_model(name, weights=None, class_names=NAMES).state_dict()
        assert [[k, list(v.shape)] for k, v in sd.items()] == ref[name], name


def _rows_match(out, ref, gap_ulps=16.0):
    """Score vectors equal to 1e-6; rows whose score is separated from both neighbours
    must match exactly in (box, cls) (tie order of torch.topk is unspecified)."""
    assert out.shape == ref.shape
    assert torch.allclose(out[:, 4], ref[:, 4], atol=2e-6, rtol=0)
    s = ref[:, 4].double()
    eps = torch.finfo(torch.float32).eps * s.abs().clamp(min=1e-30)
    gap_prev = torch.cat((torch.tensor([1e9], dtype=torch.float64), (s[:-1] - s[1:]) / eps[:-1]))
    gap_next = torch.cat(((s[:-1] - s[1:]) / eps[:-1], torch.tensor([1e9], dtype=torch.float64)))
    solid = (gap_prev > gap_ulps) & (gap_next > gap_ulps)
    assert solid.float().mean() > 0.9
    assert torch.equal(out[solid, 5], ref[solid, 5])
    assert torch.allclose(out[solid, :4], ref[solid, :4], atol=1e-3, rtol=1e-5)


def test_decode_topk_matches_reference_golden():
    g = torch.load(os.path.join(G, "decode_topk.pt"))
    logits = synth_head_logits(2, g["nc"], g["hw"], seed=g["seed"])
    dets = O.decode_topk(logits, num_classes=g["nc"])
    for i in range(2):
        _rows_match(dets[i][0], g["out"][i])
    sm = g["small"]
    d = O.decode_topk(synth_head_logits(1, sm["nc"], sm["hw"], reg_max=sm["reg_max"], seed=sm["seed"]),
                      num_classes=sm["nc"], strides=(8,), max_det=sm["max_det"])[0][0]
    _rows_match(d, sm["out"], gap_ulps=4.0)
    r1 = g["regmax1"]
    d = O.decode_topk(synth_head_logits(1, r1["nc"], r1["hw"], reg_max=1, seed=r1["seed"]), num_classes=r1["nc"])[0][0]
    assert d.shape == r1["out"].shape
    assert torch.allclose(d[:, 4], r1["out"][:, 4], atol=2e-6)


def _canon(d):
    """Order rows by (score desc, x1, y1): removes the arbitrary order the reference's
    unstable argsort gives to exactly tied scores (content must still be identical)."""
    if d.numel() == 0:
        return d
    key = torch.stack((-d[:, 4].double(), d[:, 0].double(), d[:, 1].double()), 1)
    idx = sorted(range(d.shape[0]), key=lambda i: tuple(key[i].tolist()))
    return d[idx]


def test_decode_nms_matches_reference_golden():
    g = torch.load(os.path.join(G, "decode_nms.pt"))
    for tag, case in g["cases"].items():
        logits = synth_head_logits(len(case["out"]), g["nc"], g["hw"], seed=case["seed"], cls_mean=case["cls_mean"])
        dets = O.decode_nms(logits, num_classes=g["nc"], conf_thresh=case["conf"], iou_thresh=case["iou"], max_det=300)
        for d, ref in zip(dets, case["out"]):
            assert d[0].shape == ref.shape, tag
            if ref.numel():
                assert torch.allclose(_canon(d[0]), _canon(ref), atol=1e-4, rtol=1e-6), tag


def test_nms_keep_matches_reference_golden():
    g = torch.load(os.path.join(G, "nms.pt"))
    gen = torch.Generator().manual_seed(g["seed"])
    n = g["n"]
    xy = torch.rand(n, 2, generator=gen) * 600
    wh = torch.rand(n, 2, generator=gen) * 120 + 4
    boxes = torch.cat((xy, xy + wh), 1)
    scores = (torch.randperm(n, generator=gen).float() + 0.5) / n
    for thr, keep in g["keep"].items():
        assert torch.equal(O.nms(boxes, scores, float(thr)), keep)


# ---------------------------------------------------------------- restated reference KATs
def _one_anchor(b, h, w, reg_max, nc, sel, bins, cls_idx, cls_logit=8.0, bin_logit=20.0):
    x = torch.zeros(b, 4 * reg_max + nc, h, w)
    for side, tb in enumerate(bins):
        x[0, side * reg_max + int(tb), sel[0], sel[1]] = bin_logit
    x[0, 4 * reg_max + cls_idx, sel[0], sel[1]] = cls_logit
    return x


def test_kat_dfl_geometry():  # test_postprocess_v10_ext.py:21-53
    s, nc, reg_max, sel, bins, ci = 8, 5, 8, (1, 2), (2, 3, 4, 5), 2
    out = O.decode_topk([_one_anchor(1, 3, 4, reg_max, nc, sel, bins, ci)], num_classes=nc, strides=(s,), max_det=10)[0][0]
    lg = torch.zeros(4, reg_max)
    for i, tb in enumerate(bins):
        lg[i, tb] = 20.0
    dist = (lg.softmax(1) * torch.arange(reg_max).float()).sum(1) * s
    cx, cy = (sel[1] + 0.5) * s, (sel[0] + 0.5) * s
    exp = torch.tensor([cx - dist[0], cy - dist[1], cx + dist[2], cy + dist[3]])
    row = out[(out[:, 5].long() == ci).nonzero()[0, 0]]
    assert torch.allclose(row[:4], exp, atol=1e-5)


def test_kat_topk_respects_max_det():  # test_postprocess_v10_ext.py:101-108
    p = torch.zeros(1, 4 * 8 + 4, 2, 2)
    p[:, 32] = 5.0
    assert O.decode_topk([p], num_classes=4, strides=(8,), max_det=3)[0][0].shape == (3, 6)


def test_kat_regmax1_does_not_crash():  # test_postprocess.py:6-25
    p = torch.zeros(1, 7, 2, 2)
    p[0, 4, 0, 0] = 10.0
    d = O.decode_topk([p, torch.zeros_like(p), torch.zeros_like(p)], num_classes=3)[0][0]
    assert d.shape[1] == 6 and d.shape[0] >= 1
    assert (d[:, :4] >= 0).all() and (d[:, :4] <= 64).all()


def test_kat_iou_and_nms():  # test_box_ops_extra.py:25-47
    a = torch.tensor([[0.0, 0.0, 2.0, 2.0]])
    b = torch.tensor([[0.0, 0.0, 2.0, 2.0], [10.0, 10.0, 12.0, 12.0]])
    assert torch.allclose(O.box_iou(a, b), torch.tensor([[1.0, 0.0]]), atol=1e-6)
    assert O.nms(torch.zeros((0, 4)), torch.zeros((0,)), 0.5).numel() == 0
    boxes = torch.tensor([[0.0, 0.0, 10.0, 10.0], [1.0, 1.0, 9.0, 9.0], [20.0, 20.0, 21.0, 21.0]])
    assert set(O.nms(boxes, torch.tensor([0.9, 0.8, 0.1]), 0.5).tolist()) == {0, 2}


def test_kat_nms_decode_conf_filter_per_image():
    lg = synth_head_logits(2, 4, [(4, 4)], reg_max=4, seed=5, cls_mean=-8.0)
    lg[0][0, 16 + 1, 1, 1] = 4.0
    dets = O.decode_nms(lg, num_classes=4, strides=(8,), conf_thresh=0.5, iou_thresh=0.5, max_det=5)
    assert dets[0][0].shape == (1, 6) and dets[1][0].shape == (0, 6)


def test_classwise_nms_only_suppresses_equal_labels():
    boxes = torch.tensor([[0.0, 0.0, 10.0, 10.0], [1.0, 1.0, 9.0, 9.0], [0.5, 0.5, 9.5, 9.5]])
    scores = torch.tensor([0.9, 0.8, 0.7])
    labels = torch.tensor([0, 1, 0])
    assert O.nms(boxes, scores, 0.5).tolist() == [0]
    assert O.nms_classwise(boxes, scores, labels, 0.5).tolist() == [0, 1]


# ------------------------------------------------------------------ export-style outputs (export.py:126-198)
def _export_rows_match(det, num, rdet, rnum, gap_ulps=16.0):
    """num_dets equal; valid rows: scores equal to 2e-6, and rows whose score is separated from both neighbours
    (torch.topk tie order is unspecified) equal in box (1e-3 px) and class.  nms=True zeroes the invalid rows
    (checked); nms=False leaves whatever the -1 ties pick there (not compared)."""
    assert torch.equal(num.to(torch.int64), rnum.to(torch.int64)), (num.tolist(), rnum.tolist())
    for b in range(det.shape[0]):
        n = int(rnum[b])
        o, r = det[b, :n], rdet[b, :n]
        assert torch.allclose(o[:, 4], r[:, 4], atol=2e-6, rtol=0)
        s = r[:, 4].double()
        if n > 2:
            ulp = torch.finfo(torch.float32).eps * s.abs()
            sep = torch.ones(n, dtype=torch.bool)
            sep[1:] &= (s[:-1] - s[1:]) > gap_ulps * ulp[1:]
            sep[:-1] &= (s[:-1] - s[1:]) > gap_ulps * ulp[:-1]
            assert sep.float().mean() > 0.5
            assert torch.allclose(o[sep][:, :4], r[sep][:, :4], atol=1e-3, rtol=0)
            assert torch.equal(o[sep][:, 5], r[sep][:, 5])


@pytest.mark.parametrize("tag", ["topk_default", "topk_lowconf", "nms_default", "nms_stress", "nms_small_k", "topk_sparse",
                                 "nms_sparse", "topk_none", "nms_none"])
def test_export_decode_matches_reference_golden(tag):
    """Pins decode_export (incl. the restated torchvision NMS and the fp32 class-offset arithmetic) to outputs of the
    reference's own YOLOv10ONNXExport wrapper running the real torchvision.ops.nms (oracle/make_golden.py)."""
    g = torch.load(os.path.join(G, "decode_export.pt"))
    c = g["cases"][tag]
    lg = synth_head_logits(g["B"], g["nc"], g["hw"], seed=c["seed"], cls_mean=c["cls_mean"])
    det, num = O.decode_export(lg, num_classes=g["nc"], imgsz=640, **c["kw"])
    assert det.shape == c["det"].shape
    _export_rows_match(det, num, c["det"], c["num"])
    if c["kw"]["nms"]:
        for b in range(det.shape[0]):
            assert float(det[b, int(num[b]):].abs().max()) == 0.0 if int(num[b]) < det.shape[1] else True
            assert float(c["det"][b, int(num[b]):].abs().max()) == 0.0 if int(num[b]) < det.shape[1] else True
