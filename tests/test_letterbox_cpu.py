"""Letterbox / unletterbox: the numpy oracle against the REFERENCE's own outputs
(tests/golden/letterbox.pt, generated with cv2 by oracle/make_golden_letterbox.py) and the host-side geometry."""
import os

import numpy as np
import torch

from oracle import letterbox_oracle as L

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "letterbox.pt"), weights_only=False)


def test_oracle_letterbox_is_bit_exact_with_the_reference():
    for g in G:
        out, gain, pad = L.letterbox(g["img"].numpy(), new_shape=g["new_shape"], **g["kwargs"])
        assert out.shape == tuple(g["out"].shape), (g["new_shape"], g["kwargs"])
        assert np.array_equal(out, g["out"].numpy()), (tuple(g["img"].shape), g["new_shape"], g["kwargs"])
        assert gain == g["gain"] and tuple(pad) == tuple(g["pad"])


def test_oracle_unletterbox_is_bit_exact_with_the_reference():
    for g in G:
        H, W = g["img"].shape[:2]
        got = L.unletterbox_coords(g["boxes"].numpy(), g["gain"], g["pad"], (H, W))
        assert np.array_equal(got, g["unletterboxed"].numpy())


def test_host_geometry_matches_the_oracle():
    from leanyolo_b200 import preprocess as P
    rng = np.random.default_rng(0)
    for _ in range(300):
        H, W = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
        ns = int(rng.choice([320, 640, 1280])) if rng.random() < 0.7 else (int(rng.choice([384, 640])), int(rng.choice([640, 512])))
        kw = dict(auto=bool(rng.random() < 0.2), scale_fill=bool(rng.random() < 0.1), scaleup=bool(rng.random() < 0.8))
        assert P.letterbox_params(H, W, ns, **kw) == L.letterbox_params(H, W, ns, **kw)


def test_gpu_entry_points_refuse_cpu_tensors():
    from leanyolo_b200 import preprocess as P
    import pytest
    with pytest.raises(RuntimeError):
        P.letterbox(torch.zeros(8, 8, 3, dtype=torch.uint8), 32)
    with pytest.raises(RuntimeError):
        P.unletterbox_coords(torch.zeros(2, 4), (1.0, 1.0), (0, 0), (8, 8))
