import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import gpu_checks_model as M
cases = [dict(name="yolov10s", precision="bf16", hw=(352, 608), B=3), dict(name="yolov10m", precision="bf16", hw=(320, 320), B=5),
         dict(name="yolov10n", precision="bf16", hw=(736, 1280), B=1), dict(name="yolov10b", precision="bf16", hw=(224, 416), B=2),
         dict(name="yolov10s", precision="bf16", hw=(640, 640), B=7, nc=20), dict(name="yolov10l", precision="bf16", hw=(96, 160), B=9)]
for c in cases:
    try:
        print("PASS", c, M.check_model(**c), flush=True)
    except Exception as e:
        print("FAIL", c, type(e).__name__, str(e)[:300], flush=True)
