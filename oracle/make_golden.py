"""Generate tests/golden/* by running the REAL reference (build container only).

    python oracle/make_golden.py            # needs /root/reference

The reference (jremillard/leanyolo) is pure Python, importable here but absent on the
GPU box, so its outputs on seeded inputs are committed as small fixtures.  Weights and
inputs are regenerated from seeds by ``leanyolo_b200.synth`` on both sides, so only the
reference's OUTPUTS are stored.  Nothing under tests/, smoke() or bench.py reads
/root/reference at run time.

Fixtures
  state_keys.json            state_dict key order + shapes of the six reference variants
  forward_<variant>.pt       reference eval forward @64x64, batch 1: c3..p5 taps (module
                             outputs) and both head branches
  decode_topk.pt             reference decode_v10_official_topk on seeded logits (2 images,
                             640x640 pyramid, nc=80) + reg_max=1 / max_det edge cases
  decode_nms.pt              reference decode_v10_predictions (two threshold settings)
  nms.pt                     reference box_ops.nms keep indices on seeded boxes
  forward640_yolov10s.pt     reference eval forward of yolov10s @640x640, batch 1 (the headline resolution): 2048 sampled
                             positions + double-precision sum / abs-sum of each of the six head tensors
  decode_export.pt           reference YOLOv10ONNXExport.forward (export.py:126-198: top-k with conf mask + clamp, and the
                             class-wise pre-top-k NMS through the real torchvision.ops.nms) on seeded head logits
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("LEANYOLO_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from leanyolo.models import get_model as ref_get_model  # noqa: E402  (the reference)
from leanyolo.models.yolov10.postprocess import decode_v10_official_topk, decode_v10_predictions  # noqa: E402
from leanyolo.utils.box_ops import nms as ref_nms  # noqa: E402

from leanyolo_b200.synth import synth_head_logits, synth_images, synth_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
VARIANTS = ["yolov10n", "yolov10s", "yolov10m", "yolov10b", "yolov10l", "yolov10x"]
NAMES = [f"class{i}" for i in range(80)]
GAIN = 1.25


def min_rel_gap(v: torch.Tensor) -> float:
    """Smallest gap between consecutive (descending) values, in units of fp32 ulps at that value."""
    d = (v[:-1] - v[1:]).double()
    ulp = torch.finfo(torch.float32).eps * v[:-1].abs().clamp(min=1e-30).double()
    return float((d / ulp).min())


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    keys = {}
    for vi, name in enumerate(VARIANTS):
        model = ref_get_model(name, weights=None, class_names=NAMES).eval()
        keys[name] = [[k, list(v.shape)] for k, v in model.state_dict().items()]
        sd = synth_state_dict(model.state_dict(), seed=100 + vi, gain=GAIN)
        model.load_state_dict(sd, strict=True)
        x = synth_images(1, 64, 64, seed=200 + vi)
        taps = {}
        with torch.no_grad():
            xin = x.float() / model.input_divide
            c3, c4, c5 = model.backbone(xin)
            p3, p4, p5 = model.neck(c3, c4, c5)
            taps.update(c3=c3, c4=c4, c5=c5, p3=p3, p4=p4, p5=p5)
            one2many = model(x)
            one2one = model._eval_branches["one2one"]
            dets = model.decode_forward(one2many)
        torch.save({"seed_weights": 100 + vi, "seed_input": 200 + vi, "gain": GAIN, "hw": 64,
                    "taps": {k: v.clone() for k, v in taps.items()},
                    "one2many": [t.clone() for t in one2many], "one2one": [t.clone() for t in one2one],
                    "dets_scores": dets[0][0][:, 4].clone()},
                   os.path.join(OUT, f"forward_{name}.pt"))
        print(name, "ok", [tuple(t.shape) for t in one2many])
    with open(os.path.join(OUT, "state_keys.json"), "w") as f:
        json.dump(keys, f)

    # ---- decode: top-k
    hw = [(80, 80), (40, 40), (20, 20)]
    logits = synth_head_logits(2, 80, hw, seed=11)
    topk = decode_v10_official_topk(logits, num_classes=80, strides=(8, 16, 32))
    out = torch.stack([d[0] for d in topk])
    gaps = [min_rel_gap(out[i, :, 4]) for i in range(2)]
    small = synth_head_logits(1, 5, [(6, 8)], reg_max=8, seed=12)
    small_out = decode_v10_official_topk(small, num_classes=5, strides=(8,), max_det=10)[0][0]
    rm1 = synth_head_logits(1, 3, [(4, 4), (2, 2), (1, 1)], reg_max=1, seed=13)
    rm1_out = decode_v10_official_topk(rm1, num_classes=3, strides=(8, 16, 32))[0][0]
    torch.save({"seed": 11, "nc": 80, "hw": hw, "out": out, "min_gap_ulps": gaps,
                "small": {"seed": 12, "nc": 5, "hw": [(6, 8)], "reg_max": 8, "max_det": 10, "out": small_out,
                          "min_gap_ulps": min_rel_gap(small_out[:, 4])},
                "regmax1": {"seed": 13, "nc": 3, "hw": [(4, 4), (2, 2), (1, 1)], "out": rm1_out}},
               os.path.join(OUT, "decode_topk.pt"))
    print("topk gaps (ulps):", gaps)

    # ---- decode: NMS (class-agnostic, as the reference code does)
    nms_cases = {}
    for tag, conf, iou, seed, mean in (("default", 0.25, 0.45, 21, -3.0), ("stress", 0.001, 0.7, 22, -2.0)):
        lg = synth_head_logits(2, 80, hw, seed=seed, cls_mean=mean)
        res = decode_v10_predictions(lg, num_classes=80, strides=(8, 16, 32), conf_thresh=conf, iou_thresh=iou, max_det=300)
        nms_cases[tag] = {"seed": seed, "cls_mean": mean, "conf": conf, "iou": iou, "out": [r[0] for r in res]}
        print("nms", tag, [tuple(r[0].shape) for r in res])
    empty = decode_v10_predictions(synth_head_logits(1, 80, hw, seed=23, cls_mean=-12.0), num_classes=80,
                                   strides=(8, 16, 32), conf_thresh=0.25, iou_thresh=0.45)
    nms_cases["empty"] = {"seed": 23, "cls_mean": -12.0, "conf": 0.25, "iou": 0.45, "out": [r[0] for r in empty]}
    torch.save({"hw": hw, "nc": 80, "cases": nms_cases}, os.path.join(OUT, "decode_nms.pt"))

    # ---- plain NMS on explicit boxes
    g = torch.Generator().manual_seed(31)
    n = 3000
    xy = torch.rand(n, 2, generator=g) * 600
    wh = torch.rand(n, 2, generator=g) * 120 + 4
    boxes = torch.cat((xy, xy + wh), 1)
    scores = (torch.randperm(n, generator=g).float() + 0.5) / n   # all distinct: no tie ambiguity
    keep = {str(thr): ref_nms(boxes, scores, thr) for thr in (0.3, 0.5, 0.7)}
    torch.save({"seed": 31, "n": n, "keep": keep}, os.path.join(OUT, "nms.pt"))
    print("nms keep sizes", {k: int(v.numel()) for k, v in keep.items()})
    # ---- export-style fixed-shape outputs (export.py:126-198), through the reference's own wrapper.  The wrapper
    # calls self.model(images): a stub with the reference's V10Detect head hands it the seeded logits.
    from leanyolo.models.yolov10.export import YOLOv10ONNXExport
    from leanyolo.models.yolov10.head import V10Detect

    class Stub(torch.nn.Module):
        def __init__(self, preds, nc):
            super().__init__()
            self.head = V10Detect(nc=nc, ch=(16, 16, 16), reg_max=16)
            self.preds = preds

        def forward(self, x):
            return self.preds

    exp_cases = {}
    for tag, seed, mean, kw in (("topk_default", 41, -3.0, dict(nms=False, conf=0.25, max_dets=300)),
                                ("topk_lowconf", 42, -2.0, dict(nms=False, conf=0.05, max_dets=100)),
                                ("nms_default", 43, -3.0, dict(nms=True, conf=0.25, iou=0.45, max_dets=300, pre_topk=1000)),
                                ("nms_stress", 44, -1.0, dict(nms=True, conf=0.001, iou=0.7, max_dets=300, pre_topk=1000)),
                                ("nms_small_k", 45, -2.0, dict(nms=True, conf=0.1, iou=0.5, max_dets=50, pre_topk=200)),
                                ("topk_sparse", 46, -6.5, dict(nms=False, conf=0.25, max_dets=300)),
                                ("nms_sparse", 46, -6.5, dict(nms=True, conf=0.25, iou=0.45, max_dets=300, pre_topk=1000)),
                                ("topk_none", 47, -12.0, dict(nms=False, conf=0.25, max_dets=300)),
                                ("nms_none", 47, -12.0, dict(nms=True, conf=0.25, iou=0.45, max_dets=300, pre_topk=1000))):
        lg = synth_head_logits(3, 80, hw, seed=seed, cls_mean=mean)
        wrap = YOLOv10ONNXExport(Stub(lg, 80), imgsz=640, **kw)
        det, num = wrap(torch.zeros(3, 3, 640, 640))
        exp_cases[tag] = dict(seed=seed, cls_mean=mean, kw=kw, det=det.clone(), num=num.clone())
        print("export", tag, tuple(det.shape), num.tolist())
    torch.save({"hw": hw, "nc": 80, "B": 3, "cases": exp_cases}, os.path.join(OUT, "decode_export.pt"))
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"golden dir: {total / 1e6:.2f} MB")


def make_640() -> None:
    """One full-resolution reference fixture (round-1 verdict: the 64x64 fixtures pin the oracle, 640x640 parity rested
    on the oracle alone).  The six head tensors are 8400 x 144 floats each: store seeded samples and checksums."""
    name, sw, si = "yolov10s", 301, 302
    model = ref_get_model(name, weights=None, class_names=NAMES).eval()
    model.load_state_dict(synth_state_dict(model.state_dict(), seed=sw, gain=GAIN), strict=True)
    x = synth_images(1, 640, 640, seed=si)
    with torch.no_grad():
        one2many = model(x)
        one2one = model._eval_branches["one2one"]
    g = torch.Generator().manual_seed(303)
    out = {"seed_weights": sw, "seed_input": si, "gain": GAIN, "hw": 640, "n_samples": 2048, "seed_samples": 303}
    for br, ts in (("one2many", one2many), ("one2one", one2one)):
        for i, t in enumerate(ts):
            idx = torch.randint(0, t.numel(), (2048,), generator=g)
            out[f"{br}{i}"] = {"shape": list(t.shape), "idx": idx, "val": t.flatten()[idx].clone(),
                               "sum": float(t.double().sum()), "abs_sum": float(t.double().abs().sum()),
                               "abs_max": float(t.abs().max())}
            print(name, 640, br, i, tuple(t.shape), out[f"{br}{i}"]["abs_sum"])
    torch.save(out, os.path.join(OUT, "forward640_yolov10s.pt"))


if __name__ == "__main__":
    if "--only-640" in sys.argv:
        make_640()
    elif "--only-export" in sys.argv:      # regenerate decode_export.pt alone (the other fixtures are unchanged)
        _src = open(__file__).read()
        _body = _src[_src.index("    # ---- export-style fixed-shape outputs"):_src.index("    total = sum(")]
        hw = [(80, 80), (40, 40), (20, 20)]
        exec(compile("if True:\n" + _body, __file__, "exec"))
    else:
        main()
        make_640()
