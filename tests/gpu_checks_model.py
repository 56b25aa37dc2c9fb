"""Whole-model GPU parity checks against the CPU oracle / reference goldens."""
from __future__ import annotations

import os

import torch

from leanyolo_b200 import get_model
from leanyolo_b200 import postprocess as PP
from leanyolo_b200.synth import synth_images, synth_state_dict
from oracle import yolov10_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")
NAMES = [f"class{i}" for i in range(80)]
DEV = "cuda"


def _errs(a, b):
    a, b = a.float().cpu(), b.float()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12)), float((a - b).norm() / b.norm().clamp(min=1e-12))


def build(name, seed=1, gain=1.25, precision="bf16", conv_impl=None, nc=80, chain=False):
    os.environ["LEANYOLO_FUSE_CHAIN"] = "1" if chain else "0"     # opt-in fused conv chains (LY_OP_CHAIN)
    if conv_impl:
        os.environ["LEANYOLO_CONV_IMPL"] = conv_impl
    else:
        os.environ.pop("LEANYOLO_CONV_IMPL", None)
    m = get_model(name, weights=None, class_names=[f"class{i}" for i in range(nc)])
    sd = synth_state_dict(m.state_dict(), seed=seed, gain=gain)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    m.precision = precision
    return m, sd


def check_model(name="yolov10s", precision="bf16", hw=64, B=2, conv_impl=None, seed=1, chain=False, nc=80):
    """bf16: raw head outputs and the c3..p5 taps within 2e-2 (max-abs / max|ref| and rel-L2);
    fp32 check mode: within 1e-4 (north_star tolerances).  ``hw``: int or (H, W) -- the reference takes any multiple of
    32 per axis (letterbox(auto=True) gives e.g. 384x640); ``nc``: class count (head widths c3 = max(ch0, min(nc, 100)))."""
    m, sd = build(name, seed=seed, precision=precision, conv_impl=conv_impl, chain=chain, nc=nc)
    H, W = (hw, hw) if isinstance(hw, int) else hw
    x = synth_images(B, H, W, seed=seed + 10)
    taps = {}
    ref = O.forward(sd, x, taps=taps)
    got = m.forward_with_taps(x.to(DEV))
    tol = 1e-4 if precision == "fp32" else 2e-2
    worst = {}
    for k in ("c3", "c4", "c5", "p3", "p4", "p5"):
        worst[k] = _errs(got[k], taps[k])
    for br in ("one2many", "one2one"):
        for i in range(3):
            assert got[br][i].shape == ref[br][i].shape and got[br][i].is_contiguous()
            worst[f"{br}{i}"] = _errs(got[br][i], ref[br][i])
    bad = {k: v for k, v in worst.items() if v[0] >= tol or v[1] >= tol}
    assert not bad, f"{name} {precision}: over tolerance {tol}: " + ", ".join(f"{k}={v[0]:.2e}/{v[1]:.2e}" for k, v in bad.items())
    return {"max_relmax": max(v[0] for v in worst.values()), "max_rel_l2": max(v[1] for v in worst.values())}


def check_model_golden(name="yolov10s"):
    """fp32 check mode against the REFERENCE's committed outputs (tests/golden/forward_*.pt)."""
    g = torch.load(os.path.join(G, f"forward_{name}.pt"))
    m, _ = build(name, seed=g["seed_weights"], gain=g["gain"], precision="fp32")
    x = synth_images(1, g["hw"], g["hw"], seed=g["seed_input"])
    got = m.forward_with_taps(x.to(DEV))
    worst = 0.0
    for k, ref in g["taps"].items():
        worst = max(worst, _errs(got[k], ref)[0])
    for br in ("one2many", "one2one"):
        for i in range(3):
            worst = max(worst, _errs(got[br][i], g[br][i])[0])
    assert worst < 1e-4, f"fp32 check mode vs reference golden: {worst:.2e}"
    return {"max_relmax": worst}


def check_model_golden_640(precision="bf16"):
    """The CUDA path against the REFERENCE's own yolov10s @640x640 outputs (tests/golden/forward640_yolov10s.pt: 2048
    sampled positions + abs-sum checksum per head tensor): bf16 within 2e-2, fp32 check mode within 1e-4."""
    g = torch.load(os.path.join(G, "forward640_yolov10s.pt"))
    m, _ = build("yolov10s", seed=g["seed_weights"], gain=g["gain"], precision=precision)
    x = synth_images(1, g["hw"], g["hw"], seed=g["seed_input"])
    raw = m(x.to(DEV))
    tol = 2e-2 if precision == "bf16" else 1e-4
    worst = 0.0
    for br, ts in (("one2many", raw), ("one2one", m._eval_branches["one2one"])):
        for i in range(3):
            ref = g[f"{br}{i}"]
            t = ts[i].float().cpu()
            assert list(t.shape) == ref["shape"]
            err = float((t.flatten()[ref["idx"]] - ref["val"]).abs().max()) / ref["abs_max"]
            worst = max(worst, err)
            assert err < tol, f"{br}{i}: {err:.2e} vs the reference golden at 640x640"
            csum = abs(float(t.double().abs().sum()) - ref["abs_sum"]) / ref["abs_sum"]
            assert csum < tol, f"{br}{i}: abs-sum checksum off by {csum:.2e}"
    return {"max_relmax": worst}


def check_subbatch_and_graph(name="yolov10s"):
    """A plan built for a sub-batch swept over the batch, and the same forward replayed from a
    CUDA graph, must give bit-identical head tensors to the one-shot run."""
    m, _ = build(name)
    x = synth_images(5, 64, 64, seed=3).to(DEV)
    full = [t.clone() for t in m(x)]
    m.sub_batch = 2
    part = m(x)
    for a, b in zip(full, part):
        # not bit-identical: the pixel brick (hence the order in which taps are accumulated) is
        # chosen per batch size; the difference is fp32 summation order, far below bf16 rounding
        err = float((a - b).abs().max() / a.abs().max())
        assert err < 5e-3, f"sub-batched sweep differs from one-shot forward by {err:.2e}"
    m.sub_batch = None
    eng = m.engine(torch.device(DEV))
    outs = eng.alloc_outputs(5, 64, 64, 5)
    eng.run(x, outs=outs)  # warm
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.run(x, outs=outs)
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(graph):
        eng.run(x, outs=outs)
    for v in outs.values():
        v.zero_()
    graph.replay()
    torch.cuda.synchronize()
    for i in range(3):
        assert torch.equal(outs[("one2many", i)], full[i]), "CUDA-graph replay differs"
    return {}


def check_decode_e2e(name="yolov10s"):
    """model.decode_forward(model(x)) on the GPU == oracle top-k decode of the SAME head tensors
    (tie-free rows index-exact), and has the reference's List[B][1] x [300,6] structure."""
    m, sd = build(name)
    x = synth_images(2, 128, 128, seed=5).to(DEV)
    raw = m(x)
    dets = m.decode_forward(raw)
    assert isinstance(dets, list) and len(dets) == 2 and len(dets[0]) == 1
    A = sum(t.shape[2] * t.shape[3] for t in raw)
    assert dets[0][0].shape == (min(300, A), 6)
    ref = O.decode_topk([t.cpu() for t in m._eval_branches["one2one"]], num_classes=80)
    for i in range(2):
        a, b = dets[i][0].cpu(), ref[i][0]
        assert torch.allclose(a[:, 4], b[:, 4], atol=2e-6)
    fixed = m.detect(x)
    assert fixed.shape == (2, min(300, A), 6)
    assert torch.equal(fixed[0], dets[0][0]) and set(m._eval_branches) == {"one2many", "one2one"}
    # opt-in fused path: the one-to-many branch is not computed; the one2one head tensors are the same values up to the
    # fp32 summation order of the first regression conv (its own N=64 GEMM instead of the merged N=128 one)
    ref_o2o = [t.clone() for t in m._eval_branches["one2one"]]
    fused = m.detect(x, one2one_only=True)
    assert set(m._eval_branches) == {"one2one"} and fused.shape == fixed.shape
    worst = max(float((u - v).abs().max() / v.abs().max()) for u, v in zip(m._eval_branches["one2one"], ref_o2o))
    assert worst < 2e-2, f"one2one-only head tensors differ by {worst:.2e}"
    assert torch.allclose(fused[..., 4], fixed[..., 4], atol=2e-2)
    m(x)
    # uint8 images are consumed by the stem kernel directly and must equal the x.float() path
    x8 = x.to(torch.uint8)
    a = [t.clone() for t in m(x8)]
    b = m(x8.float())
    for u, v in zip(a, b):
        assert torch.equal(u, v), "uint8 input path differs from the float path"
    return {"one2one_only_vs_both_relmax": worst}


def check_submodules(name="yolov10s", precision="bf16", hw=64, B=2, seed=4):
    """The reference's component surface (tests/fidelity/test_fidelity_core.py:46-62):
    model.backbone(x_normalised) -> (c3, c4, c5); model.neck(c3, c4, c5) -> (p3, p4, p5);
    model.head(feats) -> one2many list; head.forward_feat(feats, cv2, cv3) selects the branch.
    Each stage is fed the ORACLE's tensors, so the tolerances are per stage."""
    m, sd = build(name, seed=seed, precision=precision)
    x = synth_images(B, hw, hw, seed=seed + 10)
    taps = {}
    ref = O.forward(sd, x, taps=taps)
    tol = 1e-4 if precision == "fp32" else 2e-2
    worst = {}
    xn = (x / 255.0).to(DEV)
    c3, c4, c5 = m.backbone(xn)
    for k, t in zip(("c3", "c4", "c5"), (c3, c4, c5)):
        assert t.is_contiguous() and t.dtype == torch.float32 and t.shape == taps[k].shape
        worst[k] = _errs(t, taps[k])
    p = m.neck(taps["c3"].to(DEV), taps["c4"].to(DEV), taps["c5"].to(DEV))
    for k, t in zip(("p3", "p4", "p5"), p):
        worst[k] = _errs(t, taps[k])
    feats = [taps[k].to(DEV) for k in ("p3", "p4", "p5")]
    h = m.head(feats)
    o2o = m.head.forward_feat(feats, m.head.one2one_cv2, m.head.one2one_cv3)
    o2m = m.head.forward_feat(feats, m.head.cv2, m.head.cv3)
    for i in range(3):
        worst[f"head{i}"] = _errs(h[i], ref["one2many"][i])
        worst[f"o2o{i}"] = _errs(o2o[i], ref["one2one"][i])
        assert torch.equal(o2m[i], h[i])
    try:
        m.head.forward_feat(feats, m.head.cv2, m.head.one2one_cv3)
        raise AssertionError("mixed branch stacks must be rejected")
    except ValueError:
        pass
    bad = {k: v for k, v in worst.items() if v[0] >= tol or v[1] >= tol}
    assert not bad, f"{name} {precision} sub-modules over tolerance {tol}: " + ", ".join(f"{k}={v[0]:.2e}/{v[1]:.2e}" for k, v in bad.items())
    return {"max_relmax": max(v[0] for v in worst.values())}


def check_config_nms(name="yolov10m", hw=640, B=2, conf=0.001, iou=0.7, max_det=300, seed=7):
    """BASELINE config 3 at full resolution: model(x) -> decode_v10_predictions (NMS stress: with random
    weights nearly every anchor passes conf 0.001).  The GPU NMS decode of the GPU head tensors must equal the
    oracle's greedy NMS of the SAME tensors (the bit-exact keep-set test on identical boxes is nms_exact_*)."""
    m, sd = build(name, seed=seed)
    x = synth_images(B, hw, hw, seed=seed + 1).to(DEV)
    raw = m(x)
    got = PP.decode_v10_predictions(raw, num_classes=80, strides=(8, 16, 32), conf_thresh=conf, iou_thresh=iou, max_det=max_det)
    ref = O.decode_nms([t.cpu() for t in raw], num_classes=80, strides=(8, 16, 32), conf_thresh=conf, iou_thresh=iou, max_det=max_det)
    from gpu_checks_decode import _canon, flip_count, nms_on_oracle_candidates
    # exact half: NMS of the oracle-decoded candidates of the GPU's head tensors, identical inputs on both sides
    kept_exact = nms_on_oracle_candidates([t.cpu() for t in raw], conf, iou, max_det)
    n_tot = flips = 0
    for i in range(B):
        a, b = _canon(got[i][0].cpu()), _canon(ref[i][0])
        assert a.shape == b.shape, f"image {i}: kept {tuple(a.shape)} vs oracle {tuple(b.shape)}"
        flips += flip_count(b, a)
        n_tot += a.shape[0]
    # end to end each side decodes its own boxes (ulp-level differences): flips are reported and bounded at 1 %
    assert flips <= max(2, n_tot // 100), f"{flips} of {n_tot} rows differ end to end"
    return {"kept": n_tot, "flips_end_to_end": flips, "kept_exact_on_identical_candidates": kept_exact}


def check_config_large(name="yolov10l", hw=1280, B=1, seed=8):
    """BASELINE config 5 (1280x1280: 160/80/40 maps, 33600 anchors) and config 4's variant (x): bf16 head
    tensors within 2e-2 of the fp32 oracle, then the top-k decode against the oracle on the same tensors."""
    m, sd = build(name, seed=seed)
    x = synth_images(B, hw, hw, seed=seed + 1)
    ref = O.forward(sd, x)
    raw = m(x.to(DEV))
    worst = 0.0
    for br, ts in (("one2many", raw), ("one2one", m._eval_branches["one2one"])):
        for i in range(3):
            e = _errs(ts[i], ref[br][i])
            worst = max(worst, e[0], e[1])
    assert worst < 2e-2, f"{name}@{hw}: {worst:.2e}"
    dets = m.decode_forward(raw)
    refd = O.decode_topk([t.cpu() for t in m._eval_branches["one2one"]], num_classes=80)
    for i in range(B):
        assert dets[i][0].shape == (300, 6)
        assert torch.allclose(dets[i][0][:, 4].cpu(), refd[i][0][:, 4], atol=2e-6)
    return {"max_rel": worst}


def check_fullsize_properties(name="yolov10s", B=256, hw=640, seed=11):
    """BASELINE config 2 at its FULL size (batch 256 @640^2), where the CPU oracle cannot follow: size-independent
    properties of the path.
      * batch consistency: images run alone give the same head tensors as inside the batch of 256 (only the tile
        mapping, hence the fp32 summation order, may differ: rel-L2 < 5e-3);
      * top-k decode: scores descending, finite boxes, class ids integral in [0, nc), exactly max_det rows, and the
        scores equal the 300 largest of the oracle's per-image decode for a sample of images (exact top-k selection
        out of 8400 x 80 candidates);
      * NMS decode (conf 0.25 / iou 0.45): per image count <= max_det, scores descending and > conf, and NO kept
        pair with IoU > thr (the defining property of greedy class-agnostic NMS), rows past the count are zero."""
    m, sd = build(name, seed=seed)
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randint(0, 256, (B, 3, hw, hw), dtype=torch.uint8, device=DEV, generator=g)
    raw = [t.clone() for t in m(x)]
    o2o = [t.clone() for t in m._eval_branches["one2one"]]
    det = m.detect(x)
    assert det.shape == (B, 300, 6) and bool(torch.isfinite(det).all())
    sc = det[:, :, 4]
    assert bool((sc[:, :-1] >= sc[:, 1:]).all()), "top-k scores are not descending"
    cls = det[:, :, 5]
    assert bool(((cls == cls.round()) & (cls >= 0) & (cls < 80)).all())
    # batch consistency + exact top-k selection on a sample
    for i in (0, 97, B - 1):
        alone = m(x[i:i + 1])
        for lvl in range(3):
            e = _errs(alone[lvl][0], raw[lvl][i].cpu())
            assert e[1] < 5e-3, f"image {i} level {lvl}: alone vs in-batch rel-L2 {e[1]:.2e}"
        ref = O.decode_topk([t[i:i + 1].cpu() for t in o2o], num_classes=80)[0][0]
        assert torch.allclose(det[i, :, 4].cpu(), ref[:, 4], atol=2e-6), f"image {i}: top-k scores differ from the oracle"
    # NMS decode of the whole batch
    out, count, _ = PP.nms_raw(raw, num_classes=80, strides=(8, 16, 32), conf_thresh=0.25, iou_thresh=0.45, max_det=300)
    cnt = count.cpu()
    assert int(cnt.max()) <= 300 and int(cnt.min()) >= 0
    out_c = out.cpu()
    worst = 0.0
    for i in range(B):
        n = int(cnt[i])
        d = out_c[i]
        assert bool((d[n:] == 0).all()), "rows past the count must be zero"
        if n == 0:
            continue
        s = d[:n, 4]
        assert bool((s > 0.25).all()) and bool((s[:-1] >= s[1:]).all())
        if n > 1:
            iou = O.box_iou(d[:n, :4], d[:n, :4])
            iou.fill_diagonal_(0)
            worst = max(worst, float(iou.max()))
    assert worst <= 0.45 + 1e-6, f"two kept boxes overlap with IoU {worst:.4f} > 0.45"
    return {"max_kept_iou": worst, "mean_kept": float(cnt.float().mean())}


def check_pack_cache(name="yolov10s"):
    """the second ``get_model(weights=file).to('cuda')`` + first forward takes the packed blobs
    from the cache next to the checkpoint (no fp64 fold) and gives bit-identical head tensors; timed."""
    import tempfile
    import time
    with tempfile.TemporaryDirectory() as d:
        ck = os.path.join(d, f"{name}.pt")
        m0 = get_model(name, weights=None, class_names=NAMES)
        torch.save(synth_state_dict(m0.state_dict(), seed=7, gain=1.25), ck)
        x = synth_images(2, 128, 128, seed=8).to(DEV)
        outs, secs, hits = [], [], []
        for _ in range(2):
            m = get_model(name, weights=ck, class_names=NAMES).to(DEV).eval()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            y = [t.clone() for t in m(x)]
            torch.cuda.synchronize()
            secs.append(time.perf_counter() - t0)
            eng = m.engine(torch.device(DEV, torch.cuda.current_device()))
            hits.append(bool(eng.pack_cache is not None and eng.pack_cache.hit))
            outs.append(y)
            packs = eng.pack_seconds
        assert hits == [False, True], hits
        assert any(f.endswith(".lypack") for f in os.listdir(d)), os.listdir(d)
        for a, b in zip(*outs):
            assert torch.equal(a, b), "cached pack gives different head tensors"
        # a second shape on the warm engine lowers dry (no folding): must match a cold engine bit for bit
        x2 = synth_images(1, 64, 96, seed=9).to(DEV)
        y_warm = [t.clone() for t in m(x2)]
        os.environ["LEANYOLO_PACK_CACHE"] = "0"
        try:
            mc = get_model(name, weights=ck, class_names=NAMES).to(DEV).eval()
            y_cold = mc(x2)
        finally:
            os.environ.pop("LEANYOLO_PACK_CACHE")
        for a, b in zip(y_warm, y_cold):
            assert torch.equal(a, b)
    return {"first_forward_s_cold": round(secs[0], 3), "first_forward_s_cached": round(secs[1], 3), "pack_seconds_cached": round(packs, 4)}
