"""Deterministic synthetic weights and inputs (there is no network for checkpoints).

``synth_state_dict`` fills a reference-format ``state_dict`` from a seed with values
that keep activations O(1) through ~60 layers and make BN folding non-trivial
(running_mean~N(0,0.1), running_var~U(0.5,1.5), gamma~U(0.8,1.2),
beta~N(0,0.1)).  The same function is used by the golden-vector generator (reference
side), the oracle tests and the bench, so all of them see identical parameters.
"""
from __future__ import annotations

import math
from typing import Dict

import torch


def synth_state_dict(template: Dict[str, torch.Tensor], seed: int = 0, gain: float = 1.7) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for k, v in template.items():
        shape = tuple(v.shape)
        if k in ("input_subtract", "input_divide", "head.dfl.bins"):
            out[k] = v.detach().clone()
        elif k.endswith("num_batches_tracked"):
            out[k] = torch.zeros_like(v)
        elif k.endswith("bn.weight"):
            out[k] = torch.empty(shape).uniform_(0.8, 1.2, generator=g)
        elif k.endswith("bn.bias") or k.endswith("running_mean"):
            out[k] = torch.empty(shape).normal_(0.0, 0.1, generator=g)
        elif k.endswith("running_var"):
            out[k] = torch.empty(shape).uniform_(0.5, 1.5, generator=g)
        elif k.endswith(".bias"):        # final head convs
            out[k] = torch.empty(shape).normal_(0.0, 0.5, generator=g)
        elif k.endswith(".weight") and v.dim() == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            out[k] = torch.empty(shape).normal_(0.0, gain / math.sqrt(fan_in), generator=g)
        else:
            raise KeyError(f"synth_state_dict: unexpected key {k}")
    return out


def synth_images(batch: int, h: int, w: int, seed: int = 0) -> torch.Tensor:
    """RGB images in [0,255] (the reference's input contract), fp32 NCHW."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 3, h, w, generator=g) * 255.0


def synth_head_logits(batch: int, nc: int, hw, reg_max: int = 16, seed: int = 0, cls_mean: float = -4.0,
                      cls_std: float = 1.5):
    """Head tensors with spread-out logits for decode tests (random-init models give
    scores in 0.48-0.52, useless for top-k / NMS parity)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for (h, w) in hw:
        reg = torch.empty(batch, 4 * reg_max, h, w).normal_(0.0, 2.0, generator=g)
        cls = torch.empty(batch, nc, h, w).normal_(cls_mean, cls_std, generator=g)
        out.append(torch.cat((reg, cls), 1).contiguous())
    return out
