"""Throughput of the GPU letterbox / unletterbox kernels (CUDA events, L2 flushed between runs).
    python tools/bench_preprocess.py [B] [H] [W] [S]"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leanyolo_b200 import preprocess as P  # noqa: E402

B, H, W, S = (int(v) for v in (sys.argv[1:5] + ["256", "720", "1280", "640"][len(sys.argv) - 1:]))
dev = "cuda"
imgs = [torch.randint(0, 256, (H + (i % 3), W - (i % 5), 3), dtype=torch.uint8, device=dev) for i in range(B)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for it in range(6):
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    batch, meta = P.letterbox_batch(imgs, S)
    e1.record()
    torch.cuda.synchronize()
    if it:
        ts.append(e0.elapsed_time(e1))
ms = statistics.median(ts)
byts = sum(i.numel() for i in imgs) + batch.numel()
print(f"letterbox_batch {B} x {H}x{W} -> {S}: {ms:.3f} ms (incl. descriptor upload), {B / ms * 1e3:.0f} img/s, {byts / ms / 1e6:.0f} GB/s (source + batch bytes)")
geo = [P.letterbox_params(i.shape[0], i.shape[1], S) for i in imgs]
dd = P._descs(imgs, geo, batch.device)
ts = []
for it in range(6):
    flush.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    P._run(dd, B, batch, True, (114, 114, 114))
    e1.record()
    torch.cuda.synchronize()
    if it:
        ts.append(e0.elapsed_time(e1))
ms = statistics.median(ts)
print(f"  kernel only: {ms:.3f} ms, {byts / ms / 1e6:.0f} GB/s")
dets = torch.rand(B, 300, 6, device=dev) * S
ts = []
for it in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    P.unletterbox_dets(dets, meta)
    e1.record()
    torch.cuda.synchronize()
    if it:
        ts.append(e0.elapsed_time(e1))
print(f"unletterbox_dets [{B},300,6]: {statistics.median(ts) * 1e3:.1f} us")
