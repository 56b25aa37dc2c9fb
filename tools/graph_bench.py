"""Eager vs CUDA-graph replay of one headline step (forward of both branches into static outputs + top-k decode)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from leanyolo_b200 import get_model, postprocess as PP  # noqa: E402
from leanyolo_b200.synth import synth_state_dict  # noqa: E402
from leanyolo_b200.variants import STRIDES  # noqa: E402

dev = torch.device("cuda", 0)
names = [f"class{i}" for i in range(80)]
m = get_model("yolov10s", weights=None, class_names=names)
m.load_state_dict(synth_state_dict(m.state_dict(), seed=0, gain=1.0), strict=True)
m = m.to(dev).eval()
B = 256
x = torch.randint(0, 256, (B, 3, 640, 640), dtype=torch.uint8, device=dev).float()
eng = m.engine(dev)
outs = eng.alloc_outputs(B, 640, 640, B)


def step():
    eng.run(x, outs=outs)
    return PP.topk_raw([outs[("one2one", i)] for i in range(3)], num_classes=80, strides=STRIDES, max_det=300)[0]


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("eager ms", timeit(step))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g):
    det = step()
print("graph ms", timeit(g.replay))
os.environ["LY_PDL"] = "0"
