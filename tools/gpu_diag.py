"""Bring-up diagnostic: run every GPU parity check in crash-isolated subprocesses.

    python tools/gpu_diag.py [--filter SUBSTR] [--out gpurun_out/diag.json]

A CUDA fault (trap, illegal address) is sticky for its process, so checks run in a
worker subprocess; when the worker dies the parent records the failure and continues
with the remaining checks in a fresh worker.  Assertion failures do not kill the worker.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def registry():
    import gpu_checks as G
    import gpu_checks_decode as D
    import gpu_checks_model as M
    import gpu_checks_preprocess as PR
    import gpu_checks_chain as CH

    R = []

    def add(_n, fn, **kw):
        R.append((_n, fn, kw))

    # CUDA-core conv (check mode + bring-up)
    add("simt_f32_3x3", G.check_conv, dtype="f32", impl="simt", cin=32, cout=48, k=3)
    add("simt_f32_1x1_s2_res", G.check_conv, dtype="f32", impl="simt", cin=64, cout=80, k=1, stride=2, res=True)
    add("simt_bf16_3x3_views", G.check_conv, dtype="bf16", impl="simt", cin=32, cout=32, k=3, src_off=16, src_extra=16, dst_off=32, dst_extra=16)
    add("simt_f32_nchw", G.check_conv, dtype="f32", impl="simt", cin=64, cout=80, k=1, act=False, nchw=True, nchw_c=77)
    # tensor-core conv: start with the simplest shape, then widen
    add("tc_1x1_64_64", G.check_conv, cin=64, cout=64, k=1, H=16, W=16, B=2)
    add("tc_1x1_64_64_noact", G.check_conv, cin=64, cout=64, k=1, act=False)
    add("tc_1x1_128_256", G.check_conv, cin=128, cout=256, k=1)
    add("tc_1x1_k32", G.check_conv, cin=32, cout=64, k=1)
    add("tc_1x1_k16", G.check_conv, cin=16, cout=32, k=1)
    add("tc_1x1_k48_n48", G.check_conv, cin=48, cout=48, k=1)
    add("tc_1x1_512_512_ntiles", G.check_conv, cin=512, cout=512, k=1, H=8, W=8)
    add("tc_1x1_576_576", G.check_conv, cin=576, cout=576, k=1, H=8, W=8)
    add("tc_1x1_tail_M", G.check_conv, cin=64, cout=64, k=1, H=5, W=7, B=3)
    add("tc_3x3_64_64", G.check_conv, cin=64, cout=64, k=3, H=16, W=16)
    add("tc_3x3_32_32", G.check_conv, cin=32, cout=32, k=3, H=32, W=32)
    add("tc_3x3_16_16", G.check_conv, cin=16, cout=16, k=3, H=32, W=32)
    add("tc_3x3_128_128_nonres", G.check_conv, cin=128, cout=128, k=3, H=16, W=16)
    add("tc_3x3_s2", G.check_conv, cin=64, cout=128, k=3, stride=2, H=32, W=32)
    add("tc_3x3_s2_k32", G.check_conv, cin=32, cout=64, k=3, stride=2, H=32, W=32)
    add("tc_3x3_s2_80_b3", G.check_conv, cin=64, cout=128, k=3, stride=2, H=80, W=80, B=3)
    add("tc_3x3_s2_k128_odd", G.check_conv, cin=128, cout=64, k=3, stride=2, H=24, W=40, B=2, res=True)
    add("tc_3x3_20x20_b3", G.check_conv, cin=64, cout=64, k=3, H=20, W=20, B=3)
    add("tc_3x3_40x40_b3", G.check_conv, cin=64, cout=64, k=3, H=40, W=40, B=3)
    add("tc_3x3_odd_10x6", G.check_conv, cin=64, cout=64, k=3, H=10, W=6, B=5)
    add("tc_3x3_res", G.check_conv, cin=64, cout=64, k=3, res=True)
    add("tc_3x3_views", G.check_conv, cin=64, cout=64, k=3, src_off=64, src_extra=64, dst_off=128, dst_extra=64)
    add("tc_1x1_inplace_res", G.check_conv, cin=128, cout=64, k=1, act=False, inplace_res=True, dst_off=64, dst_extra=0)
    add("tc_1x1_nchw_80", G.check_conv, cin=128, cout=80, k=1, act=False, nchw=True)
    add("tc_1x1_nchw_pad77", G.check_conv, cin=64, cout=80, k=1, act=False, nchw=True, nchw_c=77)
    add("tc_1x1_up_128", G.check_conv, cin=128, cout=128, k=1, H=16, W=24, B=3, up=True)
    add("tc_1x1_up_256_noact", G.check_conv, cin=64, cout=256, k=1, H=20, W=20, act=False, up=True)
    add("tc_1x1_up_views", G.check_conv, cin=128, cout=64, k=1, H=8, W=8, up=True, src_off=256, dst_off=0, dst_extra=64)
    # round 2: TMA-store epilogue (resident weights, flat / brick mappings): clipping at image edges and at the tail of the pixel
    # range, channel-slice destinations (neighbouring channels must stay untouched), batch-spanning bricks; tile-parallel
    # epilogue in pair mode with a directly loaded shortcut
    add("tc_3x3_s2_k32_odd_views", G.check_conv, cin=32, cout=64, k=3, stride=2, H=26, W=38, B=3, dst_off=64, dst_extra=32)
    add("tc_1x1_views_tail", G.check_conv, cin=64, cout=64, k=1, H=5, W=7, B=3, src_off=32, src_extra=32, dst_off=64, dst_extra=32)
    add("tc_1x1_96_64", G.check_conv, cin=96, cout=64, k=1, H=12, W=20, B=3)
    add("tc_1x1_res_128", G.check_conv, cin=64, cout=128, k=1, H=9, W=11, B=4, res=True)
    add("tc_3x3_pair_res_views", G.check_conv, cin=128, cout=128, k=3, H=12, W=20, B=2, res=True, dst_off=128, dst_extra=64)
    add("tc_3x3_tiny_maps_b40", G.check_conv, cin=64, cout=64, k=3, H=4, W=4, B=40)
    add("tc_3x3_s2_tiny_b33", G.check_conv, cin=32, cout=32, k=3, stride=2, H=6, W=10, B=33)
    add("tc_3x3_big", G.check_conv, cin=64, cout=64, k=3, H=160, W=160, B=4)
    add("tc_1x1_big", G.check_conv, cin=256, cout=128, k=1, H=80, W=80, B=8)
    # fused conv chains (LY_OP_CHAIN): single stages first (bisecting), then the real blocks
    add("chain_single_1x1_64", CH.check_chain, kind="single", k=1, cin=64, cout=64, H=16, W=16)
    add("chain_single_3x3_64", CH.check_chain, kind="single", k=3, cin=64, cout=64, H=16, W=16)
    add("chain_single_3x3_32", CH.check_chain, kind="single", k=3, cin=32, cout=32, H=24, W=20)
    add("chain_single_3x3_16_n48", CH.check_chain, kind="single", k=3, cin=16, cout=48, H=12, W=28, B=3)
    add("chain_tail_64_nchw", CH.check_chain, kind="tail", c=64, cout=64, H=20, W=20, nchw=True)
    add("chain_tail_64_80", CH.check_chain, kind="tail", c=64, cout=64, H=80, W=80, B=3, nchw=True, nchw_c=61)
    # tails with a public NCHW output take the back-to-back GEMM kernel (conv_b2b.cu); odd sizes, channel padding, batch tails
    add("chain_tail_64_odd_b5", CH.check_chain, kind="tail", c=64, cout=64, H=13, W=27, B=5, nchw=True)
    add("chain_tail_32_48_nchw", CH.check_chain, kind="tail", c=32, cout=48, H=24, W=40, B=2, nchw=True, nchw_c=40)
    add("chain_tail_64_160", CH.check_chain, kind="tail", c=64, cout=64, H=160, W=160, B=2, nchw=True)
    add("chain_tail_64_act_last", CH.check_chain, kind="tail", c=64, cout=32, H=40, W=40, B=3, nchw=True, act_last=True)
    # stride-2 first stage (backbone cv1 -> c2.cv1): four parity planes, NHWC slice destination, odd plane sizes, two TMA pieces per row
    add("chain_s2_32_64_64", CH.check_chain, kind="tail", c=32, cmid=64, cout=64, H=32, W=32, B=2, stride0=2, act_last=True)
    add("chain_s2_views_odd", CH.check_chain, kind="tail", c=32, cmid=64, cout=64, H=26, W=38, B=3, stride0=2, act_last=True,
        src_off=32, src_extra=32, dst_off=0, dst_extra=32)
    add("chain_s2_wide_320", CH.check_chain, kind="tail", c=32, cmid=64, cout=64, H=16, W=320, B=2, stride0=2, act_last=True, dst_off=64, dst_extra=0)
    add("chain_s2_64_64_32_nchw", CH.check_chain, kind="tail", c=64, cmid=64, cout=32, H=24, W=40, B=2, stride0=2, nchw=True)
    add("chain_tail_32_views", CH.check_chain, kind="tail", c=32, cout=48, H=24, W=40, act_last=True, src_off=32, src_extra=16, dst_off=16, dst_extra=32)
    add("chain_two_in", CH.check_chain, kind="two_in", H=20, W=36)
    add("chain_c2f_32", CH.check_chain, kind="c2f", c=32, H=40, W=40)
    add("chain_c2f_32_160", CH.check_chain, kind="c2f", c=32, H=160, W=160, B=3)
    add("chain_c2f_32_noshortcut_odd", CH.check_chain, kind="c2f", c=32, shortcut=False, H=36, W=52, B=3, src_off=64, dst_off=64, dst_extra=64)
    add("chain_c2f_16", CH.check_chain, kind="c2f", c=16, H=48, W=48)
    # bandwidth kernels
    for dt in ("bf16", "f32"):
        add(f"dw3_{dt}", G.check_dw, dtype=dt, k=3)
        add(f"dw3_s2_res_{dt}", G.check_dw, dtype=dt, k=3, stride=2, res=True, act=False)
        add(f"dw7_{dt}", G.check_dw, dtype=dt, k=7, H=20, W=20)
        add(f"pool_{dt}", G.check_pool, dtype=dt)
        add(f"up_{dt}", G.check_up, dtype=dt)
        add(f"attn_32_64_{dt}", G.check_attn, dtype=dt)
        add(f"attn_36_72_{dt}", G.check_attn, dtype=dt, kd=36, hd=72, H=8, W=8)
        add(f"stem_{dt}", G.check_stem, dtype=dt)
        add(f"stem_norm_{dt}", G.check_stem, dtype=dt, cout=48, sub=(10.0, 20.0, 30.0), div=(58.0, 57.0, 59.0))
        add(f"export_import_{dt}", G.check_export_import, dtype=dt)
    # TMA-staged depthwise: every tile configuration / channel-block width
    add("dwtma_3_c64_80", G.check_dw, k=3, c=64, H=80, W=80)
    add("dwtma_3_c128_20_res", G.check_dw, k=3, c=128, H=20, W=20, res=True)
    add("dwtma_3_c256_40", G.check_dw, k=3, c=256, H=40, W=40, B=3)
    add("dwtma_3_c96", G.check_dw, k=3, c=96, H=24, W=24)
    add("dwtma_3_c48", G.check_dw, k=3, c=48, H=24, W=24)
    add("dwtma_3_odd", G.check_dw, k=3, c=64, H=13, W=27, B=3)
    add("dwtma_3s2_c256_80", G.check_dw, k=3, stride=2, c=256, H=80, W=80, act=False)
    add("dwtma_3s2_c64_40", G.check_dw, k=3, stride=2, c=64, H=40, W=40, act=False)
    add("dwtma_3s2_c96", G.check_dw, k=3, stride=2, c=96, H=24, W=24)
    add("dwtma_3s2_c48_odd", G.check_dw, k=3, stride=2, c=48, H=14, W=22)
    add("dwtma_7_c512_20", G.check_dw, k=7, c=512, H=20, W=20)
    add("dwtma_7_c64_40", G.check_dw, k=7, c=64, H=40, W=40)
    add("dwtma_7_c96", G.check_dw, k=7, c=96, H=16, W=16)
    add("dwtma_7_c48", G.check_dw, k=7, c=48, H=16, W=16)
    add("stem_odd_sizes", G.check_stem, H=96, W=160, cout=16, B=3)
    # fused depthwise -> 1x1 (tensor-core): every map size of the head, k-block counts 1..8, views, NCHW
    add("dwpw_128_128_20", G.check_dwpw, c=128, cout=128, H=20, W=20)
    add("dwpw_128_128_80", G.check_dwpw, c=128, cout=128, H=80, W=80, B=3)
    add("dwpw_256_128_40", G.check_dwpw, c=256, cout=128, H=40, W=40, B=2)
    add("dwpw_512_128_20", G.check_dwpw, c=512, cout=128, H=20, W=20, B=3)
    add("dwpw_64_80_odd", G.check_dwpw, c=64, cout=80, H=13, W=27, B=3)
    add("dwpw_64_256_noact", G.check_dwpw, c=64, cout=256, H=16, W=16, dw_act=False, act=False)
    add("dwpw_192_192_views", G.check_dwpw, c=192, cout=192, H=24, W=24, src_off=64, src_extra=64, dst_off=64, dst_extra=32)
    add("dwpw_128_80_nchw", G.check_dwpw, c=128, cout=80, H=20, W=20, act=False, nchw=True, nchw_c=77)
    add("dwpw_big", G.check_dwpw, c=128, cout=128, H=80, W=80, B=16)
    # decode tail
    add("topk_golden", D.check_topk_golden)
    add("topk_vs_oracle_b4", D.check_topk_vs_oracle, B=4, seed=3)
    add("topk_small_levels", D.check_topk_vs_oracle, B=2, seed=4, hw=[(6, 8)], nc=5, reg_max=8, max_det=10, strides=(8,))
    add("topk_regmax1", D.check_topk_vs_oracle, B=1, seed=5, hw=[(4, 4), (2, 2), (1, 1)], nc=3, reg_max=1)
    add("nms_exact_3000", D.check_nms_exact, n=3000)
    add("nms_exact_classwise", D.check_nms_exact, n=2000, classwise=True)
    add("nms_exact_9000_global_sort", D.check_nms_exact, n=20000, thr=0.6)
    add("nms_golden", D.check_nms_golden)
    add("decode_nms_default", D.check_decode_nms, conf=0.25, iou=0.45, cls_mean=-3.0, seed=21)
    add("decode_nms_stress", D.check_decode_nms, conf=0.001, iou=0.7, cls_mean=-2.0, seed=22)
    add("decode_nms_empty", D.check_decode_nms, conf=0.25, iou=0.45, cls_mean=-12.0, seed=23)
    add("decode_nms_direct", D.check_decode_nms_direct)
    # export-style fixed-shape outputs + the class-wise pre-top-k NMS (export.py:126-198), pinned to the reference's wrapper
    add("export_golden", D.check_export_golden)
    add("export_nms_vs_oracle_img0", D.check_export_vs_oracle)
    add("export_nms_big_offset", D.check_export_vs_oracle, B=2, seed=52, img0=5000, iou=0.45, conf=0.05)
    add("export_topk_vs_oracle", D.check_export_vs_oracle, nms=False, seed=53, conf=0.3, max_dets=200, cls_mean=-5.0)
    add("export_nms_small_pyramid", D.check_export_vs_oracle, B=2, seed=54, hw=[(6, 8), (3, 4)], nc=5, strides=(8, 16), imgsz=64,
        pre_topk=1000, max_dets=300, conf=0.01, cls_mean=-1.0)
    add("export_topk_small_pyramid", D.check_export_vs_oracle, B=2, seed=55, nms=False, hw=[(6, 8), (3, 4)], nc=5, strides=(8, 16),
        imgsz=64, max_dets=300, conf=0.2, cls_mean=-1.0)
    # whole model
    for name in ("yolov10n", "yolov10s"):
        add(f"model_{name}_f32", M.check_model, name=name, precision="fp32", hw=64, B=2)
        add(f"model_{name}_bf16_simt", M.check_model, name=name, precision="bf16", hw=64, B=2, conv_impl="simt")
        add(f"model_{name}_bf16", M.check_model, name=name, precision="bf16", hw=64, B=2)
    for name in ("yolov10m", "yolov10b", "yolov10l", "yolov10x"):
        add(f"model_{name}_bf16", M.check_model, name=name, precision="bf16", hw=64, B=2)
        add(f"model_{name}_f32", M.check_model, name=name, precision="fp32", hw=64, B=1)
    add("model_s_bf16_320", M.check_model, name="yolov10s", precision="bf16", hw=320, B=2)
    add("model_s_bf16_chain_64", M.check_model, name="yolov10s", precision="bf16", hw=64, B=2, chain=True)
    add("model_s_bf16_chain_320", M.check_model, name="yolov10s", precision="bf16", hw=320, B=2, chain=True)
    add("model_n_bf16_chain_160", M.check_model, name="yolov10n", precision="bf16", hw=160, B=3, chain=True)
    # non-square inputs (H != W, W < H, odd H/32 or W/32) and class counts that change the head widths / NCHW padding
    add("model_s_bf16_384x640", M.check_model, name="yolov10s", precision="bf16", hw=(384, 640), B=2)
    add("model_s_bf16_224x96", M.check_model, name="yolov10s", precision="bf16", hw=(224, 96), B=3)
    add("model_n_f32_96x160", M.check_model, name="yolov10n", precision="fp32", hw=(96, 160), B=2)
    add("model_m_bf16_160x96", M.check_model, name="yolov10m", precision="bf16", hw=(160, 96), B=2)
    add("model_n_bf16_nc1", M.check_model, name="yolov10n", precision="bf16", hw=(64, 96), B=2, nc=1)
    add("model_n_bf16_nc90", M.check_model, name="yolov10n", precision="bf16", hw=(96, 64), B=2, nc=90)
    add("model_s_bf16_nc7", M.check_model, name="yolov10s", precision="bf16", hw=64, B=2, nc=7)
    add("model_n_f32_nc90", M.check_model, name="yolov10n", precision="fp32", hw=64, B=1, nc=90)
    add("model_s_f32_nc7", M.check_model, name="yolov10s", precision="fp32", hw=(64, 96), B=1, nc=7)
    add("model_s_352x608_b3", M.check_model, name="yolov10s", precision="bf16", hw=(352, 608), B=3)
    add("model_s_640_b7_nc20", M.check_model, name="yolov10s", precision="bf16", hw=(640, 640), B=7, nc=20)
    add("model_s_golden", M.check_model_golden, name="yolov10s")
    add("model_s_golden_640_bf16", M.check_model_golden_640, precision="bf16")
    add("model_s_golden_640_f32", M.check_model_golden_640, precision="fp32")
    add("model_s_subbatch_graph", M.check_subbatch_and_graph, name="yolov10s")
    add("model_s_decode_e2e", M.check_decode_e2e, name="yolov10s")
    add("model_s_pack_cache", M.check_pack_cache, name="yolov10s")
    add("model_x_pack_cache", M.check_pack_cache, name="yolov10x")
    # the reference's component surface: backbone / neck / head / forward_feat callables
    add("submodules_s_bf16", M.check_submodules, name="yolov10s", precision="bf16")
    add("submodules_s_f32", M.check_submodules, name="yolov10s", precision="fp32")
    add("submodules_m_bf16", M.check_submodules, name="yolov10m", precision="bf16", B=1)
    # BASELINE.json configs 3-5 at their real resolutions (small batches: the CPU oracle has to follow)
    add("config3_m_640_nms_stress", M.check_config_nms, name="yolov10m", hw=640, B=2)
    add("config4_x_640", M.check_config_large, name="yolov10x", hw=640, B=2)
    add("config5_l_1280", M.check_config_large, name="yolov10l", hw=1280, B=1)
    add("model_n_640_vs_oracle", M.check_config_large, name="yolov10n", hw=640, B=2)
    add("model_s_640_vs_oracle", M.check_config_large, name="yolov10s", hw=640, B=2)
    add("model_b_640_vs_oracle", M.check_config_large, name="yolov10b", hw=640, B=1)
    add("config2_s_640_b256_properties", M.check_fullsize_properties, name="yolov10s", B=256, hw=640)
    # pre / post-processing around the path: bit-exact with the reference's cv2 letterbox
    add("letterbox_golden", PR.check_letterbox_golden)
    add("letterbox_batch_640", PR.check_letterbox_batch)
    add("unletterbox_batch", PR.check_unletterbox_batch)
    add("detect_images_e2e", PR.check_detect_images)
    add("val_loop_coco", PR.check_val_loop)
    add("letterbox_fused_stem", PR.check_fused_letterbox)
    return R


def worker(names):
    import torch  # noqa: F401
    reg = {n: (f, kw) for n, f, kw in registry()}
    for n in names:
        f, kw = reg[n]
        t = time.time()
        try:
            res = f(**kw)
            print("RESULT " + json.dumps({"name": n, "ok": True, "res": res, "s": round(time.time() - t, 2)}), flush=True)
        except AssertionError as e:
            print("RESULT " + json.dumps({"name": n, "ok": False, "err": "ASSERT: " + str(e)[:300], "s": round(time.time() - t, 2)}), flush=True)
        except Exception as e:  # CUDA errors are sticky: stop this worker
            print("RESULT " + json.dumps({"name": n, "ok": False, "err": f"{type(e).__name__}: {str(e)[:400]}", "fatal": True}), flush=True)
            return


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", default=None)
    ap.add_argument("--filter", default="")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "diag.json"))
    ap.add_argument("--timeout", type=int, default=300)
    a = ap.parse_args()
    if a.worker is not None:
        worker(a.worker.split(","))
        return
    names = [n for n, _, _ in registry() if a.filter in n]
    results = []
    todo = list(names)
    while todo:
        try:
            p = subprocess.run([sys.executable, __file__, "--worker", ",".join(todo)], capture_output=True, text=True, timeout=a.timeout)
            out, err = p.stdout, p.stderr
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            err = "TIMEOUT"
        done = [json.loads(l[7:]) for l in out.splitlines() if l.startswith("RESULT ")]
        results += done
        finished = {d["name"] for d in done}
        rest = [n for n in todo if n not in finished]
        if done and done[-1].get("fatal"):
            done[-1]["stderr"] = err[-600:]
        elif rest:  # worker died without reporting (trap / segfault / timeout) on rest[0]
            results.append({"name": rest[0], "ok": False, "err": "worker died: " + err[-600:]})
            rest = rest[1:]
        todo = rest
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(results, open(a.out, "w"), indent=1)
    bad = [r for r in results if not r["ok"]]
    for r in results:
        print(("PASS " if r["ok"] else "FAIL ") + r["name"] + "  " + (json.dumps(r.get("res")) if r["ok"] else r["err"].replace("\n", " | ")[:500]))
    print(f"{len(results) - len(bad)}/{len(results)} checks passed")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
