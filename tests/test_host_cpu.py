"""CPU-side tests: drop-in API contract, weight plumbing, plan lowering, C-ABI surface.

Mirrors the reference's own API tests (leanyolo/tests/test_get_model_api.py,
test_get_model_local_weights.py, test_state_dict_roundtrip.py, test_convert_script.py,
test_weights_safe_unpickle.py) against leanyolo_b200, and checks the lowering with a
torch-CPU interpreter of the op list (tests/plan_interp.py) against the oracle.
"""
import ctypes
import os
import re
import sys
import types
import warnings

import pytest
import torch

import leanyolo_b200
from leanyolo_b200 import get_model, get_model_weights, list_models
from leanyolo_b200 import _native as N
from leanyolo_b200.plan import PlanBuilder
from leanyolo_b200.synth import synth_images, synth_state_dict
from leanyolo_b200.weights import INDEX_TO_NAME, WeightsEntry
from oracle import yolov10_oracle as O

sys.path.insert(0, os.path.dirname(__file__))
from plan_interp import run_plan  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = [f"class{i}" for i in range(80)]
VARIANTS = ["yolov10n", "yolov10s", "yolov10m", "yolov10b", "yolov10l", "yolov10x"]


# ------------------------------------------------------------------ API contract
def test_list_models_and_registry():
    assert tuple(list_models()) == tuple(VARIANTS)
    w = get_model_weights("yolov10s")()
    assert list(w.list("yolov10s")) == ["PRETRAINED_COCO"]
    e = w.get("yolov10s", "PRETRAINED_COCO")
    assert e.filename == "yolov10s.pt" and e.url.endswith("/v1.1/yolov10s.pt") and len(e.sha256) == 64
    with pytest.raises(KeyError):
        w.get("yolov10s", "NOPE")
    with pytest.raises(ValueError):
        get_model_weights("yolov9")


def test_get_model_errors_match_reference():
    with pytest.raises(ValueError, match="Unknown model"):
        get_model("yolov99", weights=None, class_names=NAMES)
    with pytest.raises(ValueError, match="weights must be a filename, 'PRETRAINED_COCO', or None"):
        get_model("yolov10n", weights="DEFAULT", class_names=NAMES)
    with pytest.raises(ValueError):
        get_model("yolov10n", weights=None, class_names=NAMES, input_norm_subtract=[0.0, 1.0])
    with pytest.raises(TypeError):
        get_model("yolov10n", None, NAMES)  # weights / class_names are keyword-only


def test_norm_broadcast_and_buffers():
    m = get_model("yolov10n", weights=None, class_names=["a", "b"], input_norm_subtract=[0.5], input_norm_divide=[2.0])
    assert m.input_subtract.shape == (1, 3, 1, 1) and torch.allclose(m.input_subtract.flatten(), torch.tensor([0.5] * 3))
    assert torch.allclose(m.input_divide.flatten(), torch.tensor([2.0] * 3))
    d = get_model("yolov10n", weights=None, class_names=["a", "b"])
    assert torch.allclose(d.input_divide.flatten(), torch.tensor([255.0] * 3))
    assert d.class_names == ["a", "b"] and d.head.nc == 2 and d.head.reg_max == 16 and d.training


def test_local_weights_roundtrip_and_mismatch(tmp_path):
    m = get_model("yolov10n", weights=None, class_names=["a", "b"])
    sd = synth_state_dict(m.state_dict(), seed=3)
    p = tmp_path / "w.pt"
    torch.save(sd, p)
    m2 = get_model("yolov10n", weights=str(p), class_names=["a", "b"])
    for k, v in m2.state_dict().items():
        assert torch.equal(v, sd[k]), k
    torch.save({"state_dict": sd, "model_name": "yolov10n"}, p)
    m3 = get_model("yolov10n", weights=str(p), class_names=["a", "b"])
    assert torch.equal(m3.state_dict()["head.cv3.0.2.bias"], sd["head.cv3.0.2.bias"])
    with pytest.raises(ValueError, match="compatible with this library version"):
        get_model("yolov10s", weights=str(p), class_names=["a", "b"])
    with pytest.raises(ValueError, match="compatible with this library version"):
        get_model("yolov10n", weights=str(p), class_names=["a", "b", "c"])


def test_pretrained_resolves_from_weights_dir_lean_format(tmp_path, monkeypatch):
    """reference: tests/test_convert_script.py — a lean state_dict named yolov10n.pt in
    LEANYOLO_WEIGHTS_DIR is picked up by 'PRETRAINED_COCO' (shape-order fill)."""
    src = get_model("yolov10n", weights=None, class_names=NAMES)
    sd = synth_state_dict(src.state_dict(), seed=9)
    torch.save(sd, tmp_path / "yolov10n.pt")
    monkeypatch.setenv("LEANYOLO_WEIGHTS_DIR", str(tmp_path))
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        m = get_model("yolov10n", weights="PRETRAINED_COCO", class_names=NAMES)
    assert any("Weights loaded" in str(w.message) for w in rec)
    got = m.state_dict()
    for k in sd:
        if not k.endswith("num_batches_tracked"):
            assert torch.equal(got[k], sd[k]), k


def _official_format(sd):
    """Invert the index map: lean keys -> 'model.<idx>.' keys with fused RepVGGDW naming."""
    inv = sorted(((v, k) for k, v in INDEX_TO_NAME.items()), key=lambda t: -len(t[0]))
    out = {}
    for k, v in sd.items():
        if k in ("input_subtract", "input_divide") or ".cv1.2.conv1." in k:
            continue  # official checkpoints have no norm buffers and ship the fused 7x7 only
        for pre, idx in inv:
            if k.startswith(pre + "."):
                nk = f"model.{idx}." + k[len(pre) + 1:]
                nk = nk.replace(".cv1.2.conv.conv.", ".cv1.2.conv.").replace(".cv1.2.conv.bn.", ".cv1.2.bn.")
                out[nk] = v
                break
    return out


def test_pretrained_official_format_pickle_with_stub_class(tmp_path, monkeypatch):
    """Official-format fake: 'model.N.' keys inside a pickled ultralytics-like object; loads
    without ultralytics installed (reference: tests/test_weights_safe_unpickle.py) and the
    fused RepVGGDW branch is zero-filled so the re-parameterised conv is exact."""
    lean = get_model("yolov10s", weights=None, class_names=NAMES)
    sd = synth_state_dict(lean.state_dict(), seed=4)
    off = _official_format(sd)
    mod = types.ModuleType("ultralytics.nn.tasks")
    pkg = types.ModuleType("ultralytics"); nn_ = types.ModuleType("ultralytics.nn")
    class YOLOv10DetectionModel:  # noqa: N801 - pickled by qualified name
        pass
    YOLOv10DetectionModel.__module__ = "ultralytics.nn.tasks"
    YOLOv10DetectionModel.__qualname__ = "YOLOv10DetectionModel"
    mod.YOLOv10DetectionModel = YOLOv10DetectionModel
    for n, mm in (("ultralytics", pkg), ("ultralytics.nn", nn_), ("ultralytics.nn.tasks", mod)):
        sys.modules[n] = mm
    try:
        obj = YOLOv10DetectionModel()
        obj._parameters, obj._buffers, obj._modules = {}, {}, {}
        holder = YOLOv10DetectionModel()
        holder._parameters, holder._buffers, holder._modules = dict(off), {}, {}
        # nest once, like ckpt['model'] = DetectionModel(model=Sequential(...))
        torch.save({"model": holder, "epoch": -1}, tmp_path / "yolov10s.pt")
    finally:
        for n in ("ultralytics.nn.tasks", "ultralytics.nn", "ultralytics"):
            sys.modules.pop(n, None)
    monkeypatch.setenv("LEANYOLO_WEIGHTS_DIR", str(tmp_path))
    with warnings.catch_warnings(record=True):
        warnings.simplefilter("always")
        m = get_model("yolov10s", weights="PRETRAINED_COCO", class_names=NAMES)
    got = m.state_dict()
    for k, v in sd.items():
        if k.endswith("num_batches_tracked") or k in ("input_subtract", "input_divide"):
            continue
        if ".cv1.2.conv1." in k:
            if k.endswith("conv.weight"):
                assert float(got[k].abs().max()) == 0.0, k
            continue
        assert torch.equal(got[k], v), k


def test_pretrained_missing_soft_fails_with_warning(tmp_path, monkeypatch):
    monkeypatch.setenv("LEANYOLO_WEIGHTS_DIR", str(tmp_path))
    monkeypatch.setenv("LEANYOLO_CACHE_DIR", str(tmp_path / "cache"))
    monkeypatch.setattr(WeightsEntry, "get_state_dict", lambda self, **kw: (_ for _ in ()).throw(FileNotFoundError("offline")))
    with pytest.warns(RuntimeWarning, match="Proceeding with randomly initialized weights"):
        m = get_model("yolov10n", weights="PRETRAINED_COCO", class_names=NAMES)
    assert m is not None


def test_cache_hash_verification(tmp_path, monkeypatch):
    monkeypatch.delenv("LEANYOLO_WEIGHTS_DIR", raising=False)
    f = tmp_path / "tiny.pt"
    torch.save({"a": torch.ones(2)}, f)
    import hashlib
    good = hashlib.sha256(f.read_bytes()).hexdigest()
    e = WeightsEntry(name="tiny", url=None, filename="tiny.pt", sha256=good)
    assert torch.equal(e.get_state_dict(cache_dir=str(tmp_path))["a"], torch.ones(2))
    bad = WeightsEntry(name="tiny", url=None, filename="tiny.pt", sha256="0" * 64)
    with pytest.raises(FileNotFoundError):
        bad.get_state_dict(cache_dir=str(tmp_path))


# ------------------------------------------------------------------ no CPU fallback
def test_cpu_tensors_are_rejected_loudly():
    m = get_model("yolov10n", weights=None, class_names=NAMES).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="CUDA"):
        leanyolo_b200.decode_v10_official_topk([torch.zeros(1, 144, 8, 8)], num_classes=80, strides=(8,))
    with pytest.raises(NotImplementedError):
        get_model("yolov10n", weights=None, class_names=NAMES)(torch.zeros(1, 3, 64, 64))  # train mode


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "leanyolo_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn
                assert "/root/reference" not in src, fn


# ------------------------------------------------------------------ C ABI surface
def test_shared_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "leanyolo_b200.h")).read()
    declared = set(re.findall(r"LY_API\s+[\w\s\*]+?\b(ly_\w+)\s*\(", header))
    assert declared == set(N.PROTOTYPES), declared ^ set(N.PROTOTYPES)
    assert N.LIB_PATH.exists(), "build with: python -m leanyolo_b200.build"
    lib = ctypes.CDLL(str(N.LIB_PATH))
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert N.lib().ly_abi_version() == N.ABI_VERSION
    assert ctypes.sizeof(N.LyView) == 32 and ctypes.sizeof(N.LyOp) == 272 and ctypes.sizeof(N.LyChain) == 672


# ------------------------------------------------------------------ lowering vs oracle
@pytest.mark.parametrize("name", VARIANTS)
def test_lowering_matches_oracle_fp32(name):
    m = get_model(name, weights=None, class_names=NAMES)
    sd = synth_state_dict(m.state_dict(), seed=1, gain=1.25)
    m.load_state_dict(sd)
    x = synth_images(2, 64, 64, seed=2)
    taps = {}
    ref = O.forward(sd, x, taps=taps)
    pb = PlanBuilder(2, 64, 64, "f32")
    m.emit(pb, taps=True)
    outs = run_plan(pb, x)
    for k in taps:
        assert float((outs[(k, 0)] - taps[k]).abs().max() / taps[k].abs().max()) < 1e-4, k
    for br in ("one2many", "one2one"):
        for i in range(3):
            assert float((outs[(br, i)] - ref[br][i]).abs().max() / ref[br][i].abs().max()) < 1e-4, (br, i)


@pytest.mark.parametrize("chain,tail", [("0", "0"), ("0", "1"), ("1", "1")])
def test_lowering_bf16_storage_emulation_within_tolerance(chain, tail, monkeypatch):
    """tail=1 (default): the 3x3 -> 1x1 regression tails lowered to LY_OP_CHAIN (back-to-back GEMM kernel);
    chain=1: also the C2f block (opt-in fused chains)."""
    monkeypatch.setenv("LEANYOLO_FUSE_CHAIN", chain)
    monkeypatch.setenv("LEANYOLO_FUSE_TAIL", tail)
    m = get_model("yolov10s", weights=None, class_names=NAMES)
    sd = synth_state_dict(m.state_dict(), seed=1, gain=1.25)
    m.load_state_dict(sd)
    x = synth_images(1, 128, 128, seed=2)
    ref = O.forward(sd, x)
    pb = PlanBuilder(1, 128, 128, "bf16")
    m.emit(pb)
    n_chain = sum(op.kind == "chain" for op in pb.ops)
    # tails: 2 branches x 3 levels; + the whole C2f block with chain fusion
    assert n_chain == (0 if tail == "0" and chain == "0" else (6 if chain == "0" else 7))
    outs = run_plan(pb, x, quant=lambda t: t.to(torch.bfloat16).float())
    for i in range(3):
        r = ref["one2one"][i]
        assert float((outs[("one2one", i)] - r).abs().max() / r.abs().max()) < 2e-2


def test_lowering_stride2_pair_emulation_within_tolerance(monkeypatch):
    """LEANYOLO_FUSE_S2=1: backbone cv1 (3x3 / s2) -> c2.cv1 (1x1) lowered to one stride-2 chain op."""
    monkeypatch.setenv("LEANYOLO_FUSE_S2", "1")
    m = get_model("yolov10s", weights=None, class_names=NAMES)
    sd = synth_state_dict(m.state_dict(), seed=1, gain=1.25)
    m.load_state_dict(sd)
    x = synth_images(1, 128, 96, seed=2)
    ref = O.forward(sd, x)
    pb = PlanBuilder(1, 128, 96, "bf16")
    m.emit(pb)
    assert sum(op.kind == "chain" and op.extra.get("stride0") == 2 for op in pb.ops) == 1
    outs = run_plan(pb, x, quant=lambda t: t.to(torch.bfloat16).float())
    for i in range(3):
        r = ref["one2one"][i]
        assert float((outs[("one2one", i)] - r).abs().max() / r.abs().max()) < 2e-2


def test_lowering_handles_nonstandard_class_count_and_norm():
    """nc=2 (head N padded to 16), nc=90 on yolov10n (cls width 90 -> padded 96), non-zero subtract."""
    for name, nc in (("yolov10n", 2), ("yolov10n", 90)):
        m = get_model(name, weights=None, class_names=[str(i) for i in range(nc)],
                      input_norm_subtract=[10.0, 20.0, 30.0], input_norm_divide=[58.0, 57.0, 59.0])
        sd = synth_state_dict(m.state_dict(), seed=5, gain=1.25)
        m.load_state_dict(sd)
        x = synth_images(1, 64, 64, seed=6)
        ref = O.forward(sd, x)
        pb = PlanBuilder(1, 64, 64, "f32")
        m.emit(pb)
        outs = run_plan(pb, x)
        for i in range(3):
            r = ref["one2many"][i]
            assert outs[("one2many", i)].shape == r.shape
            assert float((outs[("one2many", i)] - r).abs().max() / r.abs().max()) < 1e-4


def test_plan_is_pure_views_no_copy_ops_and_counts_flops():
    m = get_model("yolov10s", weights=None, class_names=NAMES)
    pb = PlanBuilder(1, 640, 640, "bf16")
    m.emit(pb)
    kinds = {}
    for op in pb.ops:
        kinds[op.kind] = kinds.get(op.kind, 0) + 1
    # 87 dense convs in the reference = stem + 86 GEMM convs, of which the first reg conv of the two
    # head branches is one GEMM per level (-3) and 12 follow a depthwise conv in a fused dw->1x1
    # launch; RepVGGDW pairs merged: 24 dw -> 22, 12 of them inside the fused launches
    # the two upsample+concat+1x1 of the top-down neck are 2 convs each (half-resolution part + skip part), no upsample op
    # round 2: the six 3x3 -> 1x1 regression tails (2 branches x 3 levels) are one back-to-back GEMM launch each (chain ops)
    assert kinds == {"stem": 1, "conv": 61, "chain": 6, "dw": 10, "dwpw": 12, "pool": 1, "attn": 1}
    # opt-in: the backbone's cv1 (3x3 / s2) -> c2.cv1 (1x1) pair as one stride-2 back-to-back launch
    os.environ["LEANYOLO_FUSE_S2"] = "1"
    try:
        pbs = PlanBuilder(1, 640, 640, "bf16")
        m.emit(pbs)
    finally:
        del os.environ["LEANYOLO_FUSE_S2"]
    assert sum(op.kind == "conv" for op in pbs.ops) == 59 and sum(op.kind == "chain" for op in pbs.ops) == 7
    assert abs(pbs.dense_flops() - pb.dense_flops()) < 1e-6 * pb.dense_flops()
    os.environ["LEANYOLO_FUSE_TAIL"] = "0"
    try:
        pbt = PlanBuilder(1, 640, 640, "bf16")
        m.emit(pbt)
    finally:
        del os.environ["LEANYOLO_FUSE_TAIL"]
    assert sum(op.kind == "conv" for op in pbt.ops) == 73 and not any(op.kind == "chain" for op in pbt.ops)
    assert abs(pbt.dense_flops() - pb.dense_flops()) < 1e-6 * pb.dense_flops()
    # opt-in fused chains (LY_OP_CHAIN): the 160x160 C2f block (4 convs) and the 3x3 -> 1x1 tails of the six regression stacks (2 each)
    os.environ["LEANYOLO_FUSE_CHAIN"] = "1"
    try:
        pbc = PlanBuilder(1, 640, 640, "bf16")
        m.emit(pbc)
    finally:
        del os.environ["LEANYOLO_FUSE_CHAIN"]
    kc = {}
    for op in pbc.ops:
        kc[op.kind] = kc.get(op.kind, 0) + 1
    assert kc == {"stem": 1, "conv": 57, "chain": 7, "dw": 10, "dwpw": 12, "pool": 1, "attn": 1}
    assert abs(pbc.dense_flops() - pb.dense_flops()) < 1e-6 * pb.dense_flops()
    # the reference does 24.625 dense GFLOP / image; applying the upsampled half of the
    # two neck 1x1 convs at half resolution removes 0.629 of them
    assert abs(pb.dense_flops() / 1e9 - 23.996) < 0.01
    os.environ["LEANYOLO_FUSE_UPCAT"] = "0"
    try:
        pb2 = PlanBuilder(1, 640, 640, "bf16")
        m.emit(pb2)
    finally:
        del os.environ["LEANYOLO_FUSE_UPCAT"]
    assert abs(pb2.dense_flops() / 1e9 - 24.625) < 0.01 and sum(op.kind == "up" for op in pb2.ops) == 2


def test_export_wrapper_surface_matches_reference():
    """leanyolo_b200.export mirrors leanyolo/models/yolov10/export.py:34-94,201-221 (constructor arguments, validation)."""
    from leanyolo_b200.export import YOLOv10ONNXExport, build_export_wrapper
    m = get_model("yolov10n", weights=None, class_names=NAMES)
    w = build_export_wrapper(m, imgsz=320, max_dets=50, conf=0.1, decode="nms", iou=0.5, pre_topk=200)
    assert isinstance(w, YOLOv10ONNXExport) and w.nms and (w.imgsz, w.max_dets, w.pre_topk) == (320, 50, 200)
    assert (w.num_classes, w.reg_max, w.strides) == (80, 16, (8, 16, 32)) and not w.model.training
    with pytest.raises(ValueError, match="V10Detect"):
        YOLOv10ONNXExport(torch.nn.Linear(2, 2))
    with pytest.raises(RuntimeError, match="CUDA"):        # no CPU fallback
        w(torch.zeros(1, 3, 64, 64))


@pytest.mark.parametrize("name", ["yolov10n", "yolov10s", "yolov10x"])
def test_workspace_reuse_never_overlaps_live_buffers(name):
    """PlanBuilder.assign_offsets: buffers whose lifetimes (first..last op touching them) intersect must not share
    bytes; the reused layout is several times smaller than one-buffer-one-range."""
    m = get_model(name, weights=None, class_names=NAMES)
    pb = PlanBuilder(2, 320, 320, "bf16")
    m.emit(pb, taps=True)
    flat = pb.assign_offsets(reuse=False)
    packed = pb.assign_offsets(reuse=True)
    assert packed * 3 < flat
    first, last = {}, {}
    for i, op in enumerate(pb.ops):
        for v in (op.src, op.dst, op.res, op.extra.get("up")):
            if v is not None:
                first.setdefault(v.buf.id, i)
                last[v.buf.id] = i
    bufs = [b for b in pb.bufs if b.id in first]
    size = {b.id: pb.B * b.H * b.W * b.C * pb.esize for b in bufs}
    for i, x in enumerate(bufs):
        assert x.offset % 1024 == 0 and x.offset + size[x.id] <= packed
        for y in bufs[i + 1:]:
            if first[x.id] <= last[y.id] and first[y.id] <= last[x.id]:
                assert x.offset + size[x.id] <= y.offset or y.offset + size[y.id] <= x.offset, (x.id, y.id)


def test_pack_cache_roundtrip_and_keying(tmp_path):
    """weights.PackCache: blobs are stored next to the checkpoint, keyed by the checkpoint's
    sha256 + the lowering's signature; a different signature or a changed checkpoint is a miss."""
    from leanyolo_b200.weights import PackCache, file_sha256
    m = get_model("yolov10n", weights=None, class_names=NAMES)
    ck = tmp_path / "w.pt"
    torch.save(synth_state_dict(m.state_dict(), seed=3), ck)
    m2 = get_model("yolov10n", weights=str(ck), class_names=NAMES)
    assert m2._weights_source == (str(ck), file_sha256(str(ck)))
    pb = PlanBuilder(1, 64, 64, "bf16")
    m2.emit(pb)
    w, b = pb.finalize_params()
    dry = PlanBuilder(2, 128, 96, "bf16", dry=True)          # another shape, no folding: same packing
    m2.emit(dry)
    assert dry.signature() == pb.signature() and (dry._w_len, dry._b_len) == (w.numel(), b.numel()) and not dry._w
    pc = m2._pack_cache("bf16")
    assert pc is not None and pc.load(pb.signature()) is None
    path = pc.save(pb.signature(), w, b)
    assert path is not None and os.path.dirname(path) == str(tmp_path)
    got = m2._pack_cache("bf16").load(pb.signature())
    assert got is not None and torch.equal(got[0], w) and torch.equal(got[1], b)
    assert pc.load("0" * 64) is None                          # other lowering (fusion switches, package version)
    assert m2._pack_cache("f32").load(pb.signature()) is None  # other precision: other file
    m2.load_state_dict(m.state_dict())                        # parameters replaced: the checkpoint key is forgotten
    assert m2._weights_source is None and m2._pack_cache("bf16") is None
    m3 = get_model("yolov10n", weights=str(ck), class_names=NAMES).half()
    assert m3._pack_cache("bf16") is None                     # values changed by the cast


def test_deepcopy_rebinds_submodules_and_drops_engines():
    import copy
    m = get_model("yolov10n", weights=None, class_names=NAMES)
    m._engines[("x",)] = object()
    c = copy.deepcopy(m)
    assert c._engines == {} and c.backbone._root() is c and c.head._root() is c and m.backbone._root() is m
    assert c.backbone.cv0.conv.weight is not m.backbone.cv0.conv.weight
    assert all(torch.equal(a, b) for a, b in zip(c.state_dict().values(), m.state_dict().values()))
    m._engines.clear()


def test_coco_bbox_map_known_answers():
    """leanyolo_b200.val.coco_bbox_map on hand-computed cases (COCOeval protocol; pycocotools is absent here)."""
    from leanyolo_b200.val import coco_bbox_map
    gt = [{"image_id": 1, "category_id": 3, "bbox": [10, 10, 100, 100]}, {"image_id": 2, "category_id": 3, "bbox": [0, 0, 50, 50]},
          {"image_id": 2, "category_id": 7, "bbox": [20, 20, 40, 40]}]
    perfect = [dict(g, score=0.9 - 0.1 * i) for i, g in enumerate(gt)]
    assert coco_bbox_map(gt, perfect) == {"mAP50-95": 1.0, "mAP50": 1.0, "mAP75": 1.0}
    # a confident false positive ahead of the true positive: precision 1/2 at every recall level
    one = [{"image_id": 1, "category_id": 3, "bbox": [10, 10, 100, 100]}]
    d = [{"image_id": 1, "category_id": 3, "bbox": [300, 300, 20, 20], "score": 0.9},
         {"image_id": 1, "category_id": 3, "bbox": [10, 10, 100, 100], "score": 0.8}]
    assert abs(coco_bbox_map(one, d)["mAP50-95"] - 0.5) < 1e-9
    # IoU 0.64 (100x100 vs 80x80 inside): a hit for thresholds 0.50, 0.55, 0.60 only
    d = [{"image_id": 1, "category_id": 3, "bbox": [10, 10, 80, 80], "score": 0.9}]
    r = coco_bbox_map(one, d)
    assert abs(r["mAP50-95"] - 0.3) < 1e-9 and r["mAP50"] == 1.0 and r["mAP75"] == 0.0
    # a detection inside a crowd region is ignored (neither TP nor FP); a missed gt costs recall
    gtc = one + [{"image_id": 1, "category_id": 3, "bbox": [200, 200, 100, 100], "iscrowd": 1}]
    d = [{"image_id": 1, "category_id": 3, "bbox": [210, 210, 30, 30], "score": 0.95},
         {"image_id": 1, "category_id": 3, "bbox": [10, 10, 100, 100], "score": 0.5}]
    assert coco_bbox_map(gtc, d)["mAP50-95"] == 1.0
    assert coco_bbox_map(one, [])["mAP50-95"] == 0.0


# ---------------------------------------------------------------------------------------------
# C-side shape validation (ly_op_validate, run by ly_plan_create on every op): host-only code, so it is exercised
# here on the REAL op lists of every variant (nothing the lowering emits may be rejected) and on corrupted ops.
def _all_plans():
    from leanyolo_b200.engine import serialise_ops
    cases = [(n, 2, 64, 64, "bf16", {}) for n in list_models()]
    cases += [("yolov10s", 1, 352, 608, "bf16", {}), ("yolov10s", 2, 64, 64, "f32", {}), ("yolov10m", 1, 64, 96, "f32", {}),
              ("yolov10s", 1, 640, 640, "bf16", {"LEANYOLO_FUSE_CHAIN": "1"}), ("yolov10s", 1, 640, 640, "bf16", {"LEANYOLO_FUSE_S2": "1"}),
              ("yolov10s", 1, 640, 640, "bf16", {"LEANYOLO_FUSE_TAIL": "0"})]
    for name, B, H, W, dt, env in cases:
        m = get_model(name, weights=None, class_names=NAMES)
        emits = [("full", lambda pb, m=m: m.emit(pb)), ("taps", lambda pb, m=m: m.emit(pb, True)),
                 ("one2one", lambda pb, m=m: m.emit(pb, False, "one2one"))]
        if not env:
            emits += [(p, lambda pb, m=m, p=p: m._emit_part(pb, p)) for p in ("backbone", "neck", "head")]
        for tag, emit in emits:
            os.environ.update(env)
            try:
                pb = PlanBuilder(B, H, W, dt, dry=True)
                emit(pb)
            finally:
                for k in env:
                    del os.environ[k]
            pb.assign_offsets()
            for in_u8 in (False, True, "lb") if tag == "full" and dt == "bf16" else (False,):
                arr, chains, _, _ = serialise_ops(pb, B, dt, N.IMPL_AUTO, 0x7f0000000000, 0x7e0000000000, 0x7d0000000000, in_u8)
                yield f"{name} {B}x{H}x{W} {dt} {tag} {env} u8={in_u8}", pb, arr, chains


def test_op_validate_accepts_every_op_the_lowering_emits():
    lib = N.lib()
    n = 0
    for tag, pb, arr, chains in _all_plans():
        for i in range(len(pb.ops)):
            rc = lib.ly_op_validate(ctypes.byref(arr[i]))
            assert rc == 0, f"{tag}: op {i} ({pb.ops[i].kind}) rejected: {lib.ly_last_error().decode()}"
            n += 1
    assert n > 3000


def test_op_validate_rejects_bad_shapes_with_a_message():
    from leanyolo_b200.engine import serialise_ops
    lib = N.lib()
    m = get_model("yolov10n", weights=None, class_names=NAMES)
    pb = PlanBuilder(1, 64, 64, "bf16", dry=True)
    m.emit(pb)
    pb.assign_offsets()

    def fresh():
        return serialise_ops(pb, 1, "bf16", N.IMPL_AUTO, 0x7f0000000000, 0x7e0000000000, 0x7d0000000000)

    conv = next(i for i, op in enumerate(pb.ops) if op.kind == "conv")
    attn = next(i for i, op in enumerate(pb.ops) if op.kind == "attn")
    dwpw = next(i for i, op in enumerate(pb.ops) if op.kind == "dwpw")

    def bad(i, mutate, needle):
        arr, chains, _, _ = fresh()
        mutate(arr[i])
        assert lib.ly_op_validate(ctypes.byref(arr[i])) == -1        # LY_E_ARG
        assert needle in lib.ly_last_error().decode()

    bad(conv, lambda o: setattr(o, "kind", 99), "unknown op kind 99")
    bad(conv, lambda o: setattr(o, "dtype", 7), "unknown dtype 7")
    bad(conv, lambda o: setattr(o, "B", 0), "batch 0")
    bad(conv, lambda o: setattr(o.src, "c", o.src.ctot + 16), "src view")
    bad(conv, lambda o: setattr(o.dst, "H", 0), "dst view")
    bad(conv, lambda o: setattr(o, "stride", 3), "stride 3")
    bad(conv, lambda o: setattr(o, "k", 2), "k 2")
    bad(attn, lambda o: setattr(o, "nh", 0), "attention")
    bad(dwpw, lambda o: setattr(o, "pre_k", 5), "dwpw")
    assert lib.ly_op_validate(None) == -1


# ---------------------------------------------------------------------------------------------
# bench.py --impl reference (the reference's algorithm on the host cores): JSON contract of the line the driver parses
def test_bench_reference_arm_contract_and_rank_gating():
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--model", "yolov10n", "--imgsz", "64",
           "--steps", "1", "--warmup", "1", "--gpus", "2"]
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                     # exactly ONE JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/sec (fwd+decode)" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 2 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "yolov10n 64x64" in d["config"]["workload"] and "model" not in d["config"]
    # under torchrun only rank 0 runs the CPU arm; the other ranks exit 0 without work or output
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=60, env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""
