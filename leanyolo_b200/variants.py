"""Graph configuration of the six YOLOv10 variants.

One table instead of the reference's six near-identical classes
(leanyolo/models/yolov10/yolov10{n,s,m,b,l,x}.py: ``CH/HCH/REPS/TYPES`` and the
``use_lk_*`` constructor flags, e.g. yolov10s.py:62-65,83,93-94).

``width``  backbone widths CH[0..10]; ``neck`` HCH{13,16,19,22};
``reps``   REPS{2,4,6,8,13,16,19,22}; ``cib`` the merge nodes built as C2fCIB;
``lk``     the C2fCIB nodes whose CIB uses the long-kernel (7x7+3x3) branch.
p4_p5 is always C2fCIB (neck.py:98); p4_p3 is always C2f (neck.py:89).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, FrozenSet, Tuple


@dataclass(frozen=True)
class Variant:
    name: str
    width: Tuple[int, ...]
    neck: Dict[int, int]
    reps: Dict[int, int]
    cib: FrozenSet[str]
    lk: FrozenSet[str]


def _v(name, width, neck, reps, cib, lk=()):
    keys = (2, 4, 6, 8, 13, 16, 19, 22)
    return Variant(name, tuple(width), dict(zip((13, 16, 19, 22), neck)), dict(zip(keys, reps)),
                   frozenset(cib) | {"p4_p5"}, frozenset(lk))


VARIANTS: Dict[str, Variant] = {
    v.name: v
    for v in (
        _v("yolov10n", (16, 32, 32, 64, 64, 128, 128, 256, 256, 256, 256), (128, 64, 128, 256),
           (1, 2, 2, 1, 1, 1, 1, 1), (), ("p4_p5",)),
        _v("yolov10s", (32, 64, 64, 128, 128, 256, 256, 512, 512, 512, 512), (256, 128, 256, 512),
           (1, 2, 2, 1, 1, 1, 1, 1), ("c8",), ("c8", "p4_p5")),
        _v("yolov10m", (48, 96, 96, 192, 192, 384, 384, 576, 576, 576, 576), (384, 192, 384, 576),
           (2, 4, 4, 2, 2, 2, 2, 2), ("c8", "p3_p4")),
        _v("yolov10b", (64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512), (512, 256, 512, 512),
           (2, 4, 4, 2, 2, 2, 2, 2), ("c8", "p5_p4", "p3_p4")),
        _v("yolov10l", (64, 128, 128, 256, 256, 512, 512, 512, 512, 512, 512), (512, 256, 512, 512),
           (3, 6, 6, 3, 3, 3, 3, 3), ("c8", "p5_p4", "p3_p4")),
        _v("yolov10x", (80, 160, 160, 320, 320, 640, 640, 640, 640, 640, 640), (640, 320, 640, 640),
           (3, 6, 6, 3, 3, 3, 3, 3), ("c6", "c8", "p5_p4", "p3_p4")),
    )
}

STRIDES = (8, 16, 32)
REG_MAX = 16
