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
