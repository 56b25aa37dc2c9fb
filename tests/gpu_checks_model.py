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


def build(name, seed=1, gain=1.25, precision="bf16", conv_impl=None, nc=80):
    if conv_impl:
        os.environ["LEANYOLO_CONV_IMPL"] = conv_impl
    else:
        os.environ.pop("LEANYOLO_CONV_IMPL", None)
    m = get_model(name, weights=None, class_names=NAMES[:nc])
    sd = synth_state_dict(m.state_dict(), seed=seed, gain=gain)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    m.precision = precision
    return m, sd


def check_model(name="yolov10s", precision="bf16", hw=64, B=2, conv_impl=None, seed=1):
    """bf16: raw head outputs and the c3..p5 taps within 2e-2 (max-abs / max|ref| and rel-L2);
    fp32 check mode: within 1e-4 (north_star tolerances)."""
    m, sd = build(name, seed=seed, precision=precision, conv_impl=conv_impl)
    x = synth_images(B, hw, hw, seed=seed + 10)
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
    # uint8 images are consumed by the stem kernel directly and must equal the x.float() path
    x8 = x.to(torch.uint8)
    a = [t.clone() for t in m(x8)]
    b = m(x8.float())
    for u, v in zip(a, b):
        assert torch.equal(u, v), "uint8 input path differs from the float path"
    return {}
