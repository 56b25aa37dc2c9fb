"""Weight-file resolution and checkpoint adaptation (host side of the boundary).

Same observable behaviour as the reference (leanyolo/utils/weights.py:140-207,
leanyolo/utils/remap.py, leanyolo/models/yolov10/{remap,keymap}.py):

resolution order   explicit ``local_path`` -> ``$LEANYOLO_WEIGHTS_DIR/<file>`` (no hash
                   check) -> cache dir (``$LEANYOLO_CACHE_DIR`` or ``~/.cache/leanyolo``)
                   with sha256 verification, downloading when absent/corrupt.
safe loading       ``torch.load(weights_only=True)``; classes a pickled official
                   checkpoint names (``ultralytics.nn.tasks.YOLOv10DetectionModel`` ...)
                   are replaced by inert stubs so no third-party code is imported.
official -> lean   ``model.{idx}.`` prefixes map to our attribute names by index, the
                   fused RepVGGDW naming is aliased, anything left is filled in
                   state_dict order by shape, and a missing 3x3 RepVGGDW branch is
                   zero-filled with an identity BN so the re-parameterised sum is exact.
"""
from __future__ import annotations

import hashlib
import os
import re
import sys
import tempfile
import types
from dataclasses import dataclass
from typing import Any, Dict, Iterable, Optional
from urllib.parse import urlparse
from urllib.request import urlopen

import torch

RELEASE = "https://github.com/THU-MIG/yolov10/releases/download/v1.1/"

# sha256 of the official THU-MIG v1.1 checkpoints (reference: models/registry.py:104-159)
SHA256 = {
    "yolov10n": "61b91ffc99b284792dca49bf40216945833cc2a515e1a742954e6e9327cfc19e",
    "yolov10s": "96af3fc7c7169abcc4867f3e3088b761bb33cf801283c2ec05f9703d63a0ba77",
    "yolov10m": "ff2c559f11d13701abc4e0345f82851d146ecfe7035efaafcc08475cfd8b5f2d",
    "yolov10b": "3846434cbf0016b663a1ccd6d843c48468f6852f4feeddcb9f67f9182168c142",
    "yolov10l": "83769ec3cbc61f18113f612f8bdcf922396628d620682bb72966e9b148004b8b",
    "yolov10x": "6e6eae65e6c268c49a25849922e0c75a5c707d626d67170d16a97813b0f8eb79",
}


def _sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def _stub_global(mod_path: str, cls_name: str) -> None:
    """Register an inert class ``mod_path.cls_name`` as a safe unpickling global."""
    parent = None
    parts = mod_path.split(".")
    for i, part in enumerate(parts):
        full = ".".join(parts[: i + 1])
        mod = sys.modules.get(full)
        if mod is None:
            mod = types.ModuleType(full)
            sys.modules[full] = mod
            if parent is not None:
                setattr(parent, part, mod)
        parent = mod
    if not hasattr(parent, cls_name):
        setattr(parent, cls_name, type(cls_name, (object,), {"__module__": mod_path, "state_dict": lambda self: {}}))
    torch.serialization.add_safe_globals([getattr(parent, cls_name)])


def safe_load(path: str, map_location="cpu"):
    """weights_only load; unknown pickled classes become stubs (never imported)."""
    tried = set()
    for _ in range(64):
        try:
            return torch.load(path, map_location=map_location, weights_only=True)
        except Exception as ex:  # UnpicklingError naming the unsupported global
            m = re.search(r"Unsupported global: (?:GLOBAL\s+)?([\w\.]+)\.(\w+)", str(ex))
            if not m or m.group(0) in tried:
                raise
            tried.add(m.group(0))
            _stub_global(m.group(1), m.group(2))
    raise RuntimeError("Failed to safely load checkpoint with dynamic stubs")


@dataclass
class WeightsEntry:
    name: str
    url: Optional[str]
    filename: Optional[str] = None
    metadata: Optional[Dict[str, Any]] = None
    sha256: Optional[str] = None
    resolved_path: Optional[str] = None

    def _target_filename(self) -> str:
        if self.filename:
            return self.filename
        if self.url:
            return os.path.basename(urlparse(self.url).path) or f"{self.name}.pt"
        return f"{self.name}.pt"

    def _default_cache_dir(self) -> str:
        return os.environ.get("LEANYOLO_CACHE_DIR", os.path.join(os.path.expanduser("~"), ".cache", "leanyolo"))

    def get_state_dict(self, *, progress: bool = True, map_location="cpu", local_path: Optional[str] = None,
                       cache_dir: Optional[str] = None, verify_hash: bool = True):
        """Resolution order of the reference (utils/weights.py:140-207).  ``self.resolved_path`` records the file
        that was read (the pre-packed weight cache is keyed by its sha256)."""
        if local_path is not None:
            self.resolved_path = local_path
            return safe_load(local_path, map_location)
        fname = self._target_filename()
        env_dir = os.environ.get("LEANYOLO_WEIGHTS_DIR")
        if env_dir and os.path.exists(os.path.join(env_dir, fname)):
            self.resolved_path = os.path.join(env_dir, fname)
            return safe_load(os.path.join(env_dir, fname), map_location)
        cache_dir = cache_dir or self._default_cache_dir()
        os.makedirs(cache_dir, exist_ok=True)
        path = os.path.join(cache_dir, fname)

        def ok(p: str) -> bool:
            if not (verify_hash and self.sha256):
                return True
            try:
                return _sha256(p) == self.sha256
            except FileNotFoundError:
                return False

        self.resolved_path = path
        if os.path.exists(path) and ok(path):
            return safe_load(path, map_location)
        if not self.url:
            raise FileNotFoundError(f"Weights not found locally ('{path}') and no URL provided. "
                                    "Place the file in LEANYOLO_WEIGHTS_DIR or pass local_path.")
        with tempfile.NamedTemporaryFile(delete=False, dir=cache_dir) as tmp:
            with urlopen(self.url) as r:  # nosec - URL comes from the registry / caller
                for chunk in iter(lambda: r.read(1 << 20), b""):
                    tmp.write(chunk)
        os.replace(tmp.name, path)
        if not ok(path):
            try:
                os.remove(path)
            finally:
                raise RuntimeError(f"Downloaded file hash mismatch for weights '{fname}'.")
        return safe_load(path, map_location)


class WeightsResolver:
    def list(self, model_name: str) -> Iterable[str]:  # pragma: no cover - interface
        raise NotImplementedError

    def get(self, model_name: str, key: str) -> WeightsEntry:  # pragma: no cover - interface
        raise NotImplementedError


class YOLOv10Weights(WeightsResolver):
    MODEL_TO_WEIGHTS: Dict[str, Dict[str, WeightsEntry]] = {
        n: {"PRETRAINED_COCO": WeightsEntry(name=f"{n}.PRETRAINED_COCO", url=f"{RELEASE}{n}.pt", filename=f"{n}.pt",
                                            sha256=h, metadata={"task": "detection", "dataset": "coco",
                                                                "source": "THU-MIG/yolov10@v1.1"})}
        for n, h in SHA256.items()
    }

    def list(self, model_name: str) -> Iterable[str]:
        return self.MODEL_TO_WEIGHTS.get(model_name, {}).keys()

    def get(self, model_name: str, key: str) -> WeightsEntry:
        mapping = self.MODEL_TO_WEIGHTS.get(model_name)
        if not mapping or key not in mapping:
            raise KeyError(f"No weights '{key}' for model '{model_name}'.")
        return mapping[key]


# ------------------------------------------------------------------------------------------
# checkpoint -> flat tensor dict
# ------------------------------------------------------------------------------------------
_WRAPPERS = ("state_dict", "model", "ema_state_dict", "model_state", "net")


def _walk_module_like(obj, prefix: str = "") -> Dict[str, torch.Tensor]:
    """Flatten an (unpickled, possibly stubbed) nn.Module-shaped object without calling it."""
    out: Dict[str, torch.Tensor] = {}
    for attr in ("_parameters", "_buffers"):
        d = getattr(obj, attr, None)
        if isinstance(d, dict):
            out.update({prefix + k: v for k, v in d.items() if isinstance(v, torch.Tensor)})
    kids = getattr(obj, "_modules", None)
    if isinstance(kids, dict):
        for name, child in kids.items():
            out.update(_walk_module_like(child, f"{prefix}{name}."))
    return out


def _from_object(obj) -> Optional[Dict[str, torch.Tensor]]:
    sd_fn = getattr(obj, "state_dict", None)
    if callable(sd_fn):
        try:
            sd = sd_fn()
            if isinstance(sd, dict) and sd:
                return sd
        except Exception:
            pass
    flat = _walk_module_like(obj)
    return flat or None


def extract_state_dict(obj) -> Dict[str, torch.Tensor]:
    got = _from_object(obj)
    if got:
        return got
    cur = obj
    for _ in range(3):  # unwrap up to three levels of {"model": {...}} style nesting
        if not isinstance(cur, dict):
            break
        nxt = None
        for key in _WRAPPERS:
            v = cur.get(key)
            if v is None:
                continue
            got = _from_object(v)
            if got:
                return got
            if isinstance(v, dict) and v:
                nxt = v
                break
        if nxt is None:
            break
        cur = nxt
    return cur


def strip_common_prefixes(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        while k.startswith(("module.", "model.")):
            k = k.split(".", 1)[1]
        out[k] = v
    return out


def adapt_state_dict_for_lean(loaded) -> Dict[str, torch.Tensor]:
    sd = extract_state_dict(loaded)
    return strip_common_prefixes({k: v for k, v in sd.items() if isinstance(v, torch.Tensor)})


def remap_by_shape(src: Dict[str, torch.Tensor], dst: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Greedy in-order fill: each destination key takes the next source tensor of equal shape."""
    items = list(src.items())
    out, i = {}, 0
    for dk, dv in dst.items():
        while i < len(items) and items[i][1].shape != dv.shape:
            i += 1
        if i >= len(items):
            break
        out[dk] = items[i][1]
        i += 1
    return out


# official layer index -> our attribute path (reference: models/yolov10/keymap.py:6-31)
INDEX_TO_NAME = {
    0: "backbone.cv0", 1: "backbone.cv1", 2: "backbone.c2", 3: "backbone.cv3", 4: "backbone.c4",
    5: "backbone.sc5", 6: "backbone.c6", 7: "backbone.sc7", 8: "backbone.c8", 9: "backbone.sppf9",
    10: "backbone.psa10", 13: "neck.p5_p4_c2f", 16: "neck.p4_p3_c2f", 17: "neck.p3_down",
    19: "neck.p3_p4_c2f", 20: "neck.p4_down", 22: "neck.p4_p5_c2f", 23: "head",
}
_FUSED_REP = re.compile(r"\.cv1\.2\.(conv\.weight|bn\.(?:weight|bias|running_mean|running_var))$")


def remap_official_keys_by_name(src: Dict[str, Any], dst_keys) -> Dict[str, Any]:
    out = {}
    for k, v in src.items():
        m = re.match(r"model\.(\d+)\.(.+)$", k)
        if not m or int(m.group(1)) not in INDEX_TO_NAME:
            continue
        nk = f"{INDEX_TO_NAME[int(m.group(1))]}.{m.group(2)}"
        if nk in dst_keys:
            out[nk] = v
            continue
        fused = _FUSED_REP.search(nk)  # fused RepVGGDW checkpoints carry only '.cv1.2.{conv,bn}.*'
        if fused:
            alt = nk[: fused.start()] + ".cv1.2.conv." + fused.group(1)
            if alt in dst_keys:
                out[alt] = v
    return out


def remap_official_yolov10_to_lean(loaded, dst_model: torch.nn.Module) -> Dict[str, torch.Tensor]:
    raw = extract_state_dict(loaded)
    dst = dst_model.state_dict()
    named = {k: v for k, v in remap_official_keys_by_name(raw, dst).items()
             if isinstance(v, torch.Tensor) and v.shape == dst[k].shape}
    rest = {k: v for k, v in dst.items() if k not in named}
    out = dict(named)
    out.update(remap_by_shape(strip_common_prefixes({k: v for k, v in raw.items() if isinstance(v, torch.Tensor)}), rest))
    for dk in dst:  # fused checkpoint: give the absent 3x3 branch zero weights and an identity BN
        if dk.endswith(".cv1.2.conv1.conv.weight") and dk not in out and dk.replace("conv1.conv.weight", "conv.conv.weight") in out:
            stem = dk[: -len("conv.weight")]
            out[dk] = torch.zeros(dst[dk].shape)
            if stem + "bn.weight" in dst and stem + "bn.weight" not in out:
                out[stem + "bn.weight"] = torch.ones_like(dst[stem + "bn.weight"])
                out[stem + "bn.bias"] = torch.zeros_like(dst[stem + "bn.bias"])
                out[stem + "bn.running_mean"] = torch.zeros_like(dst[stem + "bn.running_mean"])
                out[stem + "bn.running_var"] = torch.ones_like(dst[stem + "bn.running_var"])
    return out


# ---------------------------------------------------------------------------------------------------
# Pre-packed weight cache: the BN-folded, re-parameterised, re-ordered, padded parameter blobs
# (bf16 / fp32 weights + fp32 biases) that the plan consumes, serialised next to the checkpoint and keyed by the
# checkpoint's sha256, so that a later ``get_model(..., weights=...)`` + first forward skips remap-independent work
# (fp64 folding and packing) and only uploads.  Reference counterpart of the key: utils/weights.py:140-207 (the
# sha256-verified cache of the .pt itself).
PACK_FORMAT = 1     # bump when the packing in plan.py / modules.py changes meaning


def file_sha256(path: str) -> str:
    return _sha256(path)


class PackCache:
    """One cache file per (checkpoint sha256, tag); ``tag`` names the variant, precision and class count."""

    def __init__(self, weights_path: str, sha256: str, tag: str):
        self.weights_path, self.sha256, self.tag = os.path.abspath(weights_path), sha256, tag
        self.hit = False

    def _candidates(self):
        base = os.path.basename(self.weights_path)
        name = f"{base}.{self.sha256[:16]}.{self.tag}.lypack"
        yield os.path.join(os.path.dirname(self.weights_path), name)
        yield os.path.join(os.environ.get("LEANYOLO_CACHE_DIR", os.path.join(os.path.expanduser("~"), ".cache", "leanyolo")), name)

    def load(self, signature: str):
        """(w, b) CPU tensors, or None when there is no valid cache for this lowering."""
        for path in self._candidates():
            if not os.path.isfile(path):
                continue
            try:
                d = torch.load(path, map_location="cpu", weights_only=True)
                if (d.get("format") == PACK_FORMAT and d.get("sha256") == self.sha256 and d.get("signature") == signature
                        and isinstance(d.get("w"), torch.Tensor) and isinstance(d.get("b"), torch.Tensor)):
                    self.hit = True
                    return d["w"], d["b"]
            except Exception:
                continue
        return None

    def save(self, signature: str, w: torch.Tensor, b: torch.Tensor) -> Optional[str]:
        payload = {"format": PACK_FORMAT, "sha256": self.sha256, "signature": signature, "tag": self.tag, "w": w.cpu(), "b": b.cpu()}
        for path in self._candidates():
            try:
                os.makedirs(os.path.dirname(path), exist_ok=True)
                with tempfile.NamedTemporaryFile(delete=False, dir=os.path.dirname(path)) as tmp:
                    torch.save(payload, tmp)
                os.replace(tmp.name, path)
                return path
            except OSError:
                continue
        return None
