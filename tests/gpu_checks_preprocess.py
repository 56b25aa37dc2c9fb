"""GPU letterbox / unletterbox against the reference goldens (cv2) and the numpy oracle: bit-exact."""
from __future__ import annotations

import os

import numpy as np
import torch

from leanyolo_b200 import preprocess as P
from oracle import letterbox_oracle as L

G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def check_letterbox_golden():
    """Every case of tests/golden/letterbox.pt (outputs of the reference's cv2 letterbox): identical bytes, gain, pad."""
    gold = torch.load(os.path.join(G, "letterbox.pt"), weights_only=False)
    for g in gold:
        out, gain, pad = P.letterbox(g["img"].to(DEV), new_shape=g["new_shape"], **g["kwargs"])
        assert tuple(out.shape) == tuple(g["out"].shape), (tuple(g["img"].shape), g["new_shape"], g["kwargs"])
        assert torch.equal(out.cpu(), g["out"]), (tuple(g["img"].shape), g["new_shape"], g["kwargs"])
        assert gain == g["gain"] and tuple(pad) == tuple(g["pad"])
        un = P.unletterbox_coords(g["boxes"].to(DEV), g["gain"], g["pad"], tuple(g["img"].shape[:2]))
        assert torch.equal(un.cpu(), g["unletterboxed"])
    return {"cases": len(gold)}


def check_letterbox_batch(S=640, seed=0):
    """A mixed-size batch in ONE launch (down-scale, up-scale, exact 2x, copy, extreme aspect) vs the oracle, CHW layout."""
    rng = np.random.default_rng(seed)
    shapes = [(375, 500), (720, 1280), (1080, 1920), (640, 640), (1280, 1280), (37, 53), (480, 640), (2000, 3000), (5, 900), (641, 3)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    batch, meta = P.letterbox_batch([torch.from_numpy(i).to(DEV) for i in imgs], S)
    assert batch.shape == (len(imgs), 3, S, S) and batch.dtype == torch.uint8
    got = batch.cpu().numpy()
    for i, img in enumerate(imgs):
        ref, gain, pad = L.letterbox(img, S)
        assert np.array_equal(got[i], ref.transpose(2, 0, 1)), shapes[i]
        m = meta[i].cpu().tolist()
        assert m[0] == np.float32(gain[0]) and m[2] == pad[0] and m[3] == pad[1] and m[4:] == list(shapes[i])
    return {"images": len(imgs)}


def check_unletterbox_batch(seed=1):
    rng = np.random.default_rng(seed)
    B, K = 5, 300
    dets = torch.tensor(rng.uniform(-50, 700, (B, K, 6)), dtype=torch.float32)
    shapes = [(375, 500), (720, 1280), (64, 64), (1080, 1920), (33, 900)]
    geo = [P.letterbox_params(h, w, 640) for h, w in shapes]
    meta = torch.tensor([[g[6], g[7], g[2], g[3], h, w] for g, (h, w) in zip(geo, shapes)], dtype=torch.float32)
    got = P.unletterbox_dets(dets.clone().to(DEV), meta.to(DEV)).cpu()
    for b in range(B):
        ref = L.unletterbox_coords(dets[b, :, :4].numpy(), (geo[b][6], geo[b][7]), (geo[b][2], geo[b][3]), shapes[b])
        assert np.array_equal(got[b, :, :4].numpy(), ref)
        assert torch.equal(got[b, :, 4:], dets[b, :, 4:])
    return {}


def check_detect_images():
    """tools/infer.py:110-138 as one GPU pipeline == oracle letterbox -> our model -> oracle unletterbox."""
    from gpu_checks_model import build
    m, _ = build("yolov10n")
    rng = np.random.default_rng(3)
    shapes = [(90, 160), (200, 120), (128, 128)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    got = P.detect_images(m, [torch.from_numpy(i).to(DEV) for i in imgs], imgsz=128)
    # the same batch built with the oracle's letterbox (identical bytes -> identical kernel results), then the oracle's unletterbox
    lbs = [L.letterbox(img, 128) for img in imgs]
    x = torch.from_numpy(np.stack([lb.transpose(2, 0, 1) for lb, _, _ in lbs])).to(DEV)
    d = m.detect(x).cpu()
    for i, (lb, gain, pad) in enumerate(lbs):
        ref = d[i].clone()
        ref[:, :4] = torch.from_numpy(L.unletterbox_coords(d[i, :, :4].numpy(), gain, pad, shapes[i]))
        assert got[i].shape == ref.shape
        assert torch.equal(got[i].cpu(), ref), f"image {i}"
        assert float(got[i][:, 0].min()) >= 0 and float(got[i][:, 2].max()) <= shapes[i][1] and float(got[i][:, 3].max()) <= shapes[i][0]
    return {}


def check_val_loop(seed=5):
    """leanyolo_b200.val (tools/val.py:90-248 batched): the batched loop's COCO rows == the per-image path (letterbox ->
    forward -> decode -> unletterbox) for both decodes, and validate_coco end to end on a temporary dataset whose
    ground truth is the model's own five best detections per image (stats must equal the evaluator run on the same rows;
    the evaluator itself has known-answer tests on the CPU side)."""
    import json
    import tempfile
    import cv2
    from gpu_checks_model import build
    from leanyolo_b200 import postprocess as PP
    from leanyolo_b200 import val as V
    m, _ = build("yolov10n")
    rng = np.random.default_rng(seed)
    shapes = [(90, 160), (200, 120), (128, 128), (60, 75), (300, 33)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    ids = [11, 7, 42, 3, 19]
    cat_ids = list(range(100, 180))
    rows = V.detect_dataset(m, zip(ids, imgs), imgsz=128, decode="topk", max_dets=50, batch_size=2, cat_ids=cat_ids)
    assert len(rows) == 5 * 50, len(rows)
    # (the reference path is run on the SAME batch compositions: the conv kernels pick their tiling per problem size, so a
    #  different batch size may sum in a different order and move a bf16 rounding)
    refs = []
    for lo in range(0, 5, 2):
        refs += P.detect_images(m, [torch.from_numpy(img).to(DEV) for img in imgs[lo:lo + 2]], imgsz=128, max_det=50)
    for i, (iid, img) in enumerate(zip(ids, imgs)):
        ref = refs[i].cpu().numpy()
        mine = [r for r in rows if r["image_id"] == iid]
        for r, (x1, y1, x2, y2, s, c) in zip(mine, ref):
            assert r["category_id"] == cat_ids[int(c)] and r["score"] == float(s), ("topk row", iid, r, float(s), int(c))
            assert r["bbox"] == [float(x1), float(y1), float(x2 - x1), float(y2 - y1)], ("topk bbox", iid, r["bbox"], x1, y1, x2, y2)
    # NMS decode: rows == decode_v10_predictions on the one2many branch + unletterbox, image by image
    rows = V.detect_dataset(m, zip(ids, imgs), imgsz=128, decode="nms", conf=0.3, iou=0.5, max_dets=40, batch_size=1)
    for iid, img in zip(ids, imgs):
        batch, meta = P.letterbox_batch([torch.from_numpy(img).to(DEV)], 128)
        d = PP.decode_v10_predictions(m(batch), num_classes=80, strides=(8, 16, 32), conf_thresh=0.3, iou_thresh=0.5, max_det=40)[0][0]
        mine = [r for r in rows if r["image_id"] == iid]
        assert len(mine) == d.shape[0], ("nms count", iid, len(mine), d.shape)
        if d.shape[0]:
            d = P.unletterbox_dets(d.clone()[None].contiguous(), meta)[0].cpu().numpy()
            for r, (x1, y1, x2, y2, s, c) in zip(mine, d):
                assert r["score"] == float(s) and r["category_id"] == int(c) and abs(r["bbox"][0] - float(x1)) < 1e-4, ("nms row", iid, r, s, c, x1)
    # end to end from files
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "images"))
        infos = []
        for iid, img in zip(ids, imgs):
            cv2.imwrite(os.path.join(tmp, "images", f"{iid}.png"), cv2.cvtColor(img, cv2.COLOR_RGB2BGR))
            infos.append({"id": iid, "file_name": f"{iid}.png", "height": img.shape[0], "width": img.shape[1]})
        cats = [{"id": 100 + i, "name": f"class{i}"} for i in range(80)]
        order = sorted(range(5), key=lambda i: f"{ids[i]}.png")       # validate_coco walks the files in name order, batches of 4
        top = V.detect_dataset(m, [(ids[i], imgs[i]) for i in order], imgsz=128, decode="topk", max_dets=50, cat_ids=cat_ids, batch_size=4)
        anns = []
        for iid in ids:
            for r in [r for r in top if r["image_id"] == iid][:5]:
                if r["bbox"][2] > 1 and r["bbox"][3] > 1:
                    anns.append({"id": len(anns) + 1, "image_id": iid, "category_id": r["category_id"], "bbox": r["bbox"],
                                 "area": r["bbox"][2] * r["bbox"][3], "iscrowd": 0})
        json.dump({"images": infos, "annotations": anns, "categories": cats}, open(os.path.join(tmp, "annotations.json"), "w"))
        stats = V.validate_coco(model=m, data_root=tmp, imgsz=128, decode="topk", max_dets=50, batch_size=4,
                                save_json=os.path.join(tmp, "out", "dets.json"))
        assert os.path.exists(os.path.join(tmp, "out", "dets.json"))
        expect = V.coco_bbox_map(anns, top)
        assert stats == expect and 0.0 < stats["mAP50-95"] <= 1.0, (stats, expect)
    return {"rows": len(rows), "gt": len(anns), **stats}


def check_fused_letterbox(seed=9):
    """Letterbox fused into the stem loader (LY_STEM_IN_LB) == letterbox kernel -> uint8 batch -> stem, bit for bit, on the
    reference-golden source images (down/up-scale, exact 2x decimation, plain copy, 1-pixel-high) and random sizes; with
    the unletterbox fused into the decode kernel the detections equal the two-step pipeline's too."""
    from gpu_checks_model import build
    m, _ = build("yolov10n")
    gold = torch.load(os.path.join(G, "letterbox.pt"), weights_only=False)
    rng = np.random.default_rng(seed)
    imgs = [g["img"] for g in gold] + [torch.from_numpy(rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
                                       for h, w in ((256, 320), (64, 80), (128, 160), (37, 411), (500, 20))]
    imgs = [i.to(DEV).contiguous() for i in imgs]
    n_modes = set()
    for tgt, kw in (((128, 160), {}), (96, {"scaleup": False}), ((64, 96), {"scale_fill": True})):
        batch, meta = P.letterbox_batch(imgs, tgt, **kw)
        two = m.detect(batch, max_det=100, lb_meta=meta)
        ref = [t.clone() for t in m._eval_branches["one2one"]] + [t.clone() for t in m._eval_branches["one2many"]]
        descs, meta2 = P.letterbox_descs(imgs, tgt, **kw)
        assert torch.equal(meta, meta2), "descriptor meta differs from letterbox_batch's"
        one = m.detect_letterboxed(descs, len(imgs), tgt, meta2, max_det=100)
        got = list(m._eval_branches["one2one"]) + list(m._eval_branches["one2many"])
        for a, b in zip(got, ref):
            assert torch.equal(a, b), f"fused letterbox head tensors differ (target {tgt}, {kw})"
        assert torch.equal(one, two), f"detections differ (target {tgt}): max abs {float((one - two).abs().max())}"
        fused_o2o = m.detect_letterboxed(descs, len(imgs), tgt, meta2, max_det=100, one2one_only=True)
        two_o2o = m.detect(batch, max_det=100, lb_meta=meta, one2one_only=True)     # same engine, two-step letterbox
        assert torch.equal(fused_o2o, two_o2o), f"one2one-only fused detections differ (target {tgt}): max abs {float((fused_o2o - two_o2o).abs().max())}"
        d = np.frombuffer(descs.cpu().numpy().tobytes(), dtype=np.dtype([("src", "<u8"), ("pitch", "<i8"), ("sh", "<i4"), ("sw", "<i4"),
                                                                         ("nh", "<i4"), ("nw", "<i4"), ("top", "<i4"), ("left", "<i4")]))
        for r in d:
            n_modes.add(0 if (r["nw"] == r["sw"] and r["nh"] == r["sh"]) else (1 if (r["sw"] == 2 * r["nw"] and r["sh"] == 2 * r["nh"]) else 2))
    assert n_modes == {0, 1, 2}, f"resize classes exercised: {n_modes}"
    return {"images": len(imgs)}
