"""ctypes binding of libleanyolo_b200.so (the C ABI in include/leanyolo_b200.h).

There is no CPU fallback: if the shared library is missing the import of any
compute path raises, loudly, with the build command.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# LEANYOLO_B200_LIB: load an experimental build (LY_BUILD_DIR=... python -m leanyolo_b200.build) instead of the in-tree one
LIB_PATH = Path(os.environ.get("LEANYOLO_B200_LIB") or Path(__file__).resolve().parent / "_lib" / "libleanyolo_b200.so")

ABI_VERSION = 5
LY_BF16, LY_F32 = 0, 1
OP_STEM, OP_CONV, OP_DW, OP_POOL, OP_UP, OP_ATTN, OP_EXPORT, OP_IMPORT, OP_DWPW, OP_CHAIN = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10
CHAIN_MAX_STAGES, CHAIN_MAX_BLOCKS, CHAIN_MAX_REGIONS = 6, 4, 8
IMPL_AUTO, IMPL_SIMT, STEM_IN_U8, STEM_IN_LB = 0, 1, 2, 3


class LyView(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32), ("ctot", C.c_int32),
                ("c0", C.c_int32), ("c", C.c_int32)]


class LyChainBlk(C.Structure):      # mirrors `ly_chain_blk`
    _fields_ = [("region", C.c_int32), ("c0", C.c_int32), ("c", C.c_int32)]


class LyChainStage(C.Structure):    # mirrors `ly_chain_stage`
    _fields_ = [("k", C.c_int32), ("act", C.c_int32), ("cout", C.c_int32), ("n_src", C.c_int32),
                ("src", LyChainBlk * CHAIN_MAX_BLOCKS), ("dst", LyChainBlk), ("res", LyChainBlk),
                ("w", C.c_void_p), ("bias", C.c_void_p)]


class LyChain(C.Structure):         # mirrors `ly_chain`
    _fields_ = [("n_regions", C.c_int32), ("n_in", C.c_int32), ("region_c", C.c_int32 * CHAIN_MAX_REGIONS),
                ("n_stages", C.c_int32), ("reserved", C.c_int32), ("st", LyChainStage * CHAIN_MAX_STAGES)]


class LyOp(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("dtype", C.c_int32), ("B", C.c_int32), ("k", C.c_int32), ("stride", C.c_int32),
        ("act", C.c_int32), ("impl", C.c_int32), ("nh", C.c_int32), ("kdp", C.c_int32), ("hd", C.c_int32),
        ("scale", C.c_float), ("sub", C.c_float * 3), ("div", C.c_float * 3),
        ("src", LyView), ("dst", LyView), ("res", LyView),
        ("w", C.c_void_p), ("bias", C.c_void_p), ("nchw", C.c_void_p),
        ("nchw_ctot", C.c_int32), ("nchw_c0", C.c_int32), ("nchw_c", C.c_int32), ("ext_slot", C.c_int32),
        ("pre_w", C.c_void_p), ("pre_bias", C.c_void_p), ("pre_k", C.c_int32), ("pre_act", C.c_int32),
        ("up", LyView),
        ("chain", C.POINTER(LyChain)),
    ]


class LyLevels(C.Structure):
    _fields_ = [
        ("preds", C.c_void_p * 4), ("H", C.c_int32 * 4), ("W", C.c_int32 * 4), ("stride", C.c_int32 * 4),
        ("n_levels", C.c_int32), ("B", C.c_int32), ("nc", C.c_int32), ("reg_max", C.c_int32),
        ("direct", C.c_int32), ("clamp_h", C.c_int32), ("clamp_w", C.c_int32),
    ]


class LyLbDesc(C.Structure):      # mirrors `ly_lb_desc` (40 bytes)
    _fields_ = [("src", C.c_void_p), ("src_pitch", C.c_int64), ("src_h", C.c_int32), ("src_w", C.c_int32),
                ("new_h", C.c_int32), ("new_w", C.c_int32), ("top", C.c_int32), ("left", C.c_int32)]


# name -> (restype, argtypes); must list every symbol declared in include/leanyolo_b200.h
PROTOTYPES = {
    "ly_abi_version": (C.c_int32, []),
    "ly_last_error": (C.c_char_p, []),
    "ly_device_check": (C.c_int32, [C.POINTER(C.c_int32)]),
    "ly_launch_count": (C.c_int64, []),
    "ly_launch": (C.c_int32, [C.POINTER(LyOp), C.c_void_p]),
    "ly_op_validate": (C.c_int32, [C.POINTER(LyOp)]),
    "ly_plan_create": (C.c_int32, [C.POINTER(LyOp), C.c_int32, C.POINTER(C.c_void_p)]),
    "ly_plan_run": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p]),
    "ly_plan_profile": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_void_p,
                                    C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "ly_plan_num_launches": (C.c_int32, [C.c_void_p]),
    "ly_plan_destroy": (None, [C.c_void_p]),
    "ly_decode_scratch_bytes": (C.c_int64, [C.POINTER(LyLevels), C.c_int32]),
    "ly_decode_topk": (C.c_int32, [C.POINTER(LyLevels), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int64, C.c_void_p]),
    "ly_decode_topk_lb": (C.c_int32, [C.POINTER(LyLevels), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int64, C.c_void_p]),
    "ly_decode_nms": (C.c_int32, [C.POINTER(LyLevels), C.c_float, C.c_float, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ly_decode_export_scratch_bytes": (C.c_int64, [C.POINTER(LyLevels), C.c_int32, C.c_int32]),
    "ly_decode_export": (C.c_int32, [C.POINTER(LyLevels), C.c_float, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ly_letterbox_u8": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "ly_unletterbox": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "ly_nms_scratch_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "ly_nms": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float,
                           C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
}

_lib = None


class NativeError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the shared library once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -m leanyolo_b200.build` "
                "(leanyolo_b200 has no CPU or PyTorch fallback path)")
        l = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        if l.ly_abi_version() != ABI_VERSION:
            raise NativeError("libleanyolo_b200.so ABI version mismatch: rebuild")
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ly_last_error().decode(errors="replace")
        raise NativeError(f"{what or 'leanyolo_b200'} failed (code {rc}): {msg}")
