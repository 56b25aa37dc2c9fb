import os, sys
os.environ["LEANYOLO_FUSE_S2"] = "1"
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import gpu_checks_model as M
for c in [dict(name="yolov10s", precision="bf16", hw=640, B=2), dict(name="yolov10s", precision="bf16", hw=(352, 608), B=3)]:
    print("PASS", c, M.check_model(**c), flush=True)
