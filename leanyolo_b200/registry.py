"""``get_model`` / ``list_models`` / ``get_model_weights`` — the public entry points.

Signatures, defaults, error types and messages follow the reference
(leanyolo/models/registry.py:171-354) so callers and the reference's own API tests
(tests/test_get_model_api.py, test_get_model_local_weights.py) carry over unchanged.
"""
from __future__ import annotations

import os
import warnings
from typing import Iterable, Optional, Sequence, Type

import torch
import torch.nn as nn

from .model import MODEL_CLASSES
from .weights import (YOLOv10Weights, adapt_state_dict_for_lean, extract_state_dict, file_sha256,
                      remap_official_yolov10_to_lean)


def list_models() -> Iterable[str]:
    return tuple(MODEL_CLASSES.keys())


def _norm3(x: Sequence[float]) -> Sequence[float]:
    if len(x) == 1:
        return [float(x[0])] * 3
    if len(x) != 3:
        raise ValueError("subtract_mean/divide must have length 1 or 3")
    return [float(v) for v in x]


def get_model(
    name: str,
    *,
    weights: Optional[str],
    class_names: Sequence[str],
    input_norm_subtract: Optional[Sequence[float]] = None,
    input_norm_divide: Optional[Sequence[float]] = None,
) -> nn.Module:
    """Build a YOLOv10 variant, optionally loading weights.

    ``weights``: ``None`` (random init) | path to a ``.pt`` holding a plain
    ``state_dict`` or ``{"state_dict": ...}`` (strict load, failure -> ``ValueError``)
    | ``"PRETRAINED_COCO"`` (official THU-MIG v1.1 file resolved from
    ``$LEANYOLO_WEIGHTS_DIR`` / cache / download; failure -> ``RuntimeWarning`` and
    random init, exactly like the reference).  The model is returned on the CPU in
    train mode; callers do ``.to("cuda").eval()``.
    """
    if name not in MODEL_CLASSES:
        raise ValueError(f"Unknown model '{name}'. Available: {list_models()}")
    sub3 = _norm3((0.0, 0.0, 0.0) if input_norm_subtract is None else input_norm_subtract)
    div3 = _norm3((255.0, 255.0, 255.0) if input_norm_divide is None else input_norm_divide)
    model = MODEL_CLASSES[name](class_names=class_names, in_channels=3, input_norm_subtract=sub3, input_norm_divide=div3)
    if weights is None:
        return model
    if isinstance(weights, str) and os.path.isfile(weights):
        try:
            _load_local_pt_into_model(weights, model)
            _note_source(model, weights)
        except Exception as e:
            raise ValueError(f"Failed to load local weights '{weights}': {e}. "
                             "Provide a state_dict compatible with this library version.")
        return model
    if weights != "PRETRAINED_COCO":
        raise ValueError("weights must be a filename, 'PRETRAINED_COCO', or None")
    try:
        _load_official_pretrained_into_model(name, model)
    except Exception as e:  # environment dependent (offline, missing file, bad hash)
        warnings.warn(f"Could not load weights '{weights}' for '{name}': {e}. "
                      "Proceeding with randomly initialized weights.", RuntimeWarning)
    return model


def get_model_weights(name: str) -> Type[YOLOv10Weights]:
    if name not in MODEL_CLASSES:
        raise ValueError(f"Unknown model '{name}'. Available: {list_models()}")
    return YOLOv10Weights


def _load_local_pt_into_model(path: str, model: nn.Module) -> None:
    """Strict load of a plain ``state_dict`` (or ``{'state_dict': ...}``); no remapping."""
    ckpt = torch.load(path, map_location="cpu", weights_only=True)
    sd = None
    if isinstance(ckpt, dict):
        inner = ckpt["state_dict"] if "state_dict" in ckpt else ckpt
        if isinstance(inner, dict):
            sd = {k: v for k, v in inner.items() if isinstance(v, torch.Tensor)} or None
    if sd is None and callable(getattr(ckpt, "state_dict", None)):
        cand = ckpt.state_dict()
        if isinstance(cand, dict) and all(isinstance(v, torch.Tensor) for v in cand.values()):
            sd = cand
    if sd is None:
        raise ValueError("expected a plain state_dict or a dict with 'state_dict'.")
    ret = model.load_state_dict(sd, strict=True)
    if ret is not None and (getattr(ret, "missing_keys", []) or getattr(ret, "unexpected_keys", [])):
        raise RuntimeError("state_dict keys mismatch for this model version")


def _note_source(model: nn.Module, path: Optional[str]) -> None:
    """Remember which file the parameters came from: the pre-packed weight cache (weights.PackCache) is keyed by its
    sha256.  Best effort: an unreadable file just means no cache."""
    try:
        if path and os.path.isfile(path):
            model._weights_source = (os.path.abspath(path), file_sha256(path))
    except OSError:
        pass


def _load_official_pretrained_into_model(model_name: str, model: nn.Module) -> None:
    entry = YOLOv10Weights().get(model_name, "PRETRAINED_COCO")
    loaded = entry.get_state_dict(progress=True)
    mapped = remap_official_yolov10_to_lean(loaded, model)
    state = mapped if mapped else adapt_state_dict_for_lean(loaded)
    missing, unexpected = model.load_state_dict(state, strict=False)
    try:
        src = extract_state_dict(loaded)
        n_src = sum(1 for v in src.values() if isinstance(v, torch.Tensor)) or 1
        n_used = sum(1 for v in state.values() if isinstance(v, torch.Tensor))
        n_dst = len(model.state_dict()) or 1
        n_filled = n_dst - len(missing)
        warnings.warn(f"Weights loaded: {n_used}/{n_src} from file ({100.0 * n_used / n_src:.1f}%), "
                      f"filled model: {n_filled}/{n_dst} params ({100.0 * n_filled / n_dst:.1f}%).", RuntimeWarning)
    except Exception:
        pass
    if unexpected:
        warnings.warn(f"Unexpected keys when loading weights: {sorted(unexpected)[:10]}...", RuntimeWarning)
    if missing:
        warnings.warn(f"Missing keys when loading weights: {sorted(missing)[:10]}...", RuntimeWarning)
    else:
        _note_source(model, entry.resolved_path)
