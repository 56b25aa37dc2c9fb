"""Generate tests/golden/letterbox.pt from the REAL reference (runs only in the build container:
needs /root/reference and cv2).  Small seeded uint8 images through leanyolo.utils.letterbox.letterbox
and boxes through leanyolo.utils.box_ops.unletterbox_coords."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from leanyolo.utils.box_ops import unletterbox_coords  # noqa: E402
from leanyolo.utils.letterbox import letterbox  # noqa: E402

CASES = [  # (H, W, new_shape, kwargs)
    (37, 53, 96, {}), (120, 90, 96, {}), (64, 64, 96, {}), (192, 128, 96, {}), (192, 192, 96, {}),        # down / up / exact 2x
    (50, 96, 96, {}), (96, 96, 96, {}), (30, 200, (64, 96), {}), (200, 30, (64, 96), {"scaleup": False}),
    (20, 20, 96, {"scaleup": False}), (45, 77, 96, {"auto": True}), (45, 77, (64, 96), {"scale_fill": True}),
    (1, 5, 32, {}), (333, 517, 160, {}), (3, 5, 64, {}),
]


def main():
    out = []
    rng = np.random.default_rng(1234)
    for H, W, ns, kw in CASES:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        lb, gain, pad = letterbox(img, new_shape=ns, **kw)
        boxes = torch.tensor(rng.uniform(-20, max(lb.shape[:2]) + 20, (7, 4)), dtype=torch.float32)
        un = unletterbox_coords(boxes, gain=gain, pad=pad, to_shape=(H, W))
        out.append(dict(img=torch.from_numpy(img), new_shape=ns, kwargs=kw, out=torch.from_numpy(np.ascontiguousarray(lb)),
                        gain=gain, pad=pad, boxes=boxes, unletterboxed=un))
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "letterbox.pt")
    torch.save(out, p)
    print("wrote", p, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
