/* leanyolo_b200 — C ABI of the B200-native YOLOv10 inference hot path.
 *
 * Drop-in boundary for jremillard/leanyolo's `get_model()` -> `model(x)` ->
 * `model.decode_forward()` path.  The reference is pure Python/PyTorch and has no
 * FFI of its own, so each entry point below names the reference Python call it
 * replaces (paths relative to the reference root).  Plain pointers and sizes
 * only: no torch types.  All pointers are DEVICE pointers unless named `h_*`.
 * Every function is asynchronous on `stream` (a cudaStream_t passed as void*),
 * performs no allocation, and returns 0 on success or a negative LY_E_* code;
 * `ly_last_error()` gives the message.  INTEGRATION.md shows the ctypes binding.
 *
 * Layouts: activations NHWC `[B,H,W,Ctot]`, element = bf16 (LY_BF16) or fp32
 * (LY_F32, the 1e-4 check mode); dense weights `[Cout_pad][kh][kw][Cin_pad]`,
 * depthwise weights `[kh*kw][C_pad]`, biases fp32.  Channel counts of views are
 * multiples of 16 (zero-padded weights make padded channels exact zeros).
 */
#ifndef LEANYOLO_B200_H
#define LEANYOLO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LY_ABI_VERSION 5

#if defined(LY_BUILD) && defined(__GNUC__)
#define LY_API __attribute__((visibility("default")))
#else
#define LY_API
#endif

enum { LY_BF16 = 0, LY_F32 = 1 };

enum {
  LY_OK = 0,
  LY_E_ARG = -1,      /* bad argument (shape, alignment, unsupported size) */
  LY_E_CUDA = -2,     /* a CUDA runtime / driver call failed                */
  LY_E_ARCH = -3,     /* device is not sm_100 (no CPU / other-arch fallback) */
};

/* op kinds of ly_op.kind */
enum {
  LY_OP_STEM = 1,     /* backbone.cv0: NCHW fp32 image -> NHWC, normalise + 3x3 s2 + SiLU   backbone.py:68, yolov10s.py:107-112 */
  LY_OP_CONV = 2,     /* dense Conv+BN+SiLU k in {1,3}, s in {1,2} (implicit GEMM)         layers.py:51-88 */
  LY_OP_DW = 3,       /* depthwise Conv+BN(+SiLU) k in {3,7}, s in {1,2}; RepVGGDW merged  layers.py:274-294,455 */
  LY_OP_POOL = 4,     /* SPPF: three chained 5x5 max-pools written into the concat buffer  layers.py:210-217 */
  LY_OP_UP = 5,       /* nearest x2 upsample into a concat slice                           layers.py:240, neck.py:116-120 */
  LY_OP_ATTN = 6,     /* PSA attention core softmax(q^T k * scale) applied to v            layers.py:369-378 */
  LY_OP_EXPORT = 7,   /* NHWC storage -> public NCHW fp32 (taps / sub-module outputs)      */
  LY_OP_IMPORT = 8,   /* public NCHW fp32 -> NHWC storage (sub-module inputs)              */
  LY_OP_DWPW = 9,     /* fused depthwise 3x3 (+BN+SiLU) -> 1x1 Conv+BN(+SiLU): the depthwise result is
                         produced straight into the GEMM's shared-memory A tile      head.py:95-107, layers.py:256-264 */
  LY_OP_CHAIN = 10,   /* a chain of dense Conv+BN(+SiLU) stages (k in {1,3}, stride 1) run per spatial tile with every
                         intermediate tensor held in shared memory: a whole C2f block (cv1 -> Bottleneck -> cv2,
                         layers.py:91-173) or the 3x3 -> 1x1 tail of a box-regression stack (head.py:86-92) in ONE
                         launch; only the block's input and output touch HBM.  Described by ly_op.chain.            */
};

/* conv implementation selector (ly_op.impl) */
enum {
  LY_IMPL_AUTO = 0,   /* bf16: tcgen05/TMEM/TMA implicit GEMM; f32: CUDA-core check kernel */
  LY_IMPL_SIMT = 1,   /* force the CUDA-core tiled kernel (bring-up / bisecting only)      */
  LY_STEM_IN_U8 = 2,  /* STEM only: the external image tensor is uint8 NCHW instead of fp32 */
  LY_STEM_IN_LB = 3,  /* STEM only (bf16): the external "image" is a DEVICE array of ly_lb_desc[B]; the loader samples the
                         letterboxed pixels (utils/letterbox.py:9-91: cv2 bilinear resize + border) straight from the
                         source images, so the letterboxed batch never exists in HBM.  Border colour = (nh, kdp, hd). */
};

/* A channel slice [c0, c0+c) of an NHWC buffer whose pixel pitch is `ctot` elements. */
typedef struct ly_view {
  void* ptr;          /* base of the whole buffer (NULL = absent) */
  int32_t H, W;
  int32_t ctot;
  int32_t c0, c;
} ly_view;

/* ---- LY_OP_CHAIN ---------------------------------------------------------------------------------
 * The tile's tensors live in shared-memory REGIONS: [pixels of the tile + halo][<= 64 channels] bf16.
 * Regions 0 .. n_in-1 are the 64-channel blocks of op.src (loaded by TMA, zero padding = OOB fill); the
 * others are written by the stages.  A stage is one implicit GEMM over the tile: its K dimension is
 * the list of source blocks (all the same width), its output goes to a region (zeroed outside the
 * image, which is the next conv's zero padding) or, for the last stage, to op.dst / op.nchw.          */
#define LY_CHAIN_MAX_STAGES 6
#define LY_CHAIN_MAX_BLOCKS 4
#define LY_CHAIN_MAX_REGIONS 8

typedef struct ly_chain_blk {
  int32_t region;     /* < 0: absent */
  int32_t c0, c;      /* channels [c0, c0 + c) of the region, multiples of 16 */
} ly_chain_blk;

typedef struct ly_chain_stage {
  int32_t k;          /* 1 or 3 (stride 1, zero padding k/2) */
  int32_t act;        /* 1 = SiLU after bias */
  int32_t cout;       /* multiple of 16, <= 256 */
  int32_t n_src;      /* source blocks in weight-column order, all of width 16, 32 or 64 */
  ly_chain_blk src[LY_CHAIN_MAX_BLOCKS];
  ly_chain_blk dst;   /* region < 0: the last stage, written to op.dst (NHWC slice) or op.nchw */
  ly_chain_blk res;   /* added after the activation (Bottleneck shortcut); region < 0: none */
  const void* w;      /* bf16 [cout][k*k][sum of the source widths] */
  const float* bias;  /* [cout] fp32 */
} ly_chain_stage;

typedef struct ly_chain {
  int32_t n_regions, n_in;
  int32_t region_c[LY_CHAIN_MAX_REGIONS];   /* channels per region: 16, 32 or 64 */
  int32_t n_stages, reserved;               /* reserved == 2: stage 0 is a 3x3 with STRIDE 2 (two-stage chains of the back-to-back
                                               kernel, conv_b2b.cu); 0: every stage has stride 1 */
  ly_chain_stage st[LY_CHAIN_MAX_STAGES];
} ly_chain;

/* One kernel launch.  Unused fields are zero. */
typedef struct ly_op {
  int32_t kind;       /* LY_OP_*  */
  int32_t dtype;      /* LY_BF16 | LY_F32: element type of src/dst/res and dense/dw weights */
  int32_t B;          /* images in this launch */
  int32_t k, stride;  /* filter size, stride */
  int32_t act;        /* 1 = SiLU after bias, 0 = none */
  int32_t impl;       /* LY_IMPL_* (CONV only) */
  int32_t nh, kdp, hd;/* ATTN: heads, padded key dim, head (value) dim */
  float scale;        /* ATTN: key_dim^-0.5 */
  float sub[3], div[3]; /* STEM: x' = (x - sub) / div per input channel */
  ly_view src, dst, res;  /* res: added after the activation (Bottleneck/CIB/PSA shortcuts) */
  const void* w;      /* weights (dtype; STEM: fp32 [Cout_pad][27]) */
  const float* bias;  /* [Cout_pad] fp32 */
  float* nchw;        /* optional public NCHW fp32 tensor [B, nchw_ctot, Ho, Wo] (CONV/EXPORT dst, STEM/IMPORT src) */
  int32_t nchw_ctot, nchw_c0, nchw_c;
  int32_t ext_slot;   /* plan only: >=0 -> `nchw` is taken from ly_plan_run's ext[] table */
  /* DWPW only: the depthwise stage that feeds the 1x1 (k, w, bias, act above describe the 1x1) */
  const void* pre_w;      /* depthwise weights [pre_k*pre_k][C_pad] (dtype) */
  const float* pre_bias;  /* [C_pad] fp32 */
  int32_t pre_k, pre_act; /* depthwise filter size (3), 1 = SiLU after the depthwise bias */
  /* CONV (bf16, tensor-core path) only: half-resolution NHWC tensor [B, Ho/2, Wo/2, up.c] added to the
   * accumulator BEFORE the activation at (h/2, w/2).  A 1x1 conv commutes with nearest x2 upsampling, so
   * conv1x1(cat[up(a), b]) = act(W_b*b + up(W_a*a) + bias): the neck's upsample + concat never
   * materialises (neck.py:116-121).  up.ptr == NULL: absent. */
  ly_view up;
  const ly_chain* chain;  /* LY_OP_CHAIN only (HOST pointer, copied by ly_launch / ly_plan_create) */
} ly_op;

/* ---- library ---------------------------------------------------------- */
LY_API int32_t ly_abi_version(void);
LY_API const char* ly_last_error(void);
/* LY_OK iff the current device is compute capability 10.x; fills sm count.  */
LY_API int32_t ly_device_check(int32_t* sm_count);
/* number of kernels launched by this process through this library (bench `gpu_launches`) */
LY_API int64_t ly_launch_count(void);

/* ---- single ops ------------------------------------------------------- */
/* Launch one op (any kind).  Replaces one `Conv.forward` / `nn.MaxPool2d` /
 * `F.interpolate` / attention matmul+softmax call site of layers.py.          */
LY_API int32_t ly_launch(const ly_op* op, void* stream);

/* ---- whole-forward plan ---------------------------------------------- */
/* Replaces `YOLOv10x.forward` (yolov10s.py:105-122): a pre-resolved op list
 * (tensor maps encoded once) executed back to back on one stream.  `ext[]`
 * carries the per-call external tensors (input image, NCHW outputs); `img0`
 * offsets them by whole images so one plan built for a sub-batch can sweep a
 * larger batch.                                                              */
typedef struct ly_plan ly_plan;
/* Shape validation of one op description on the HOST (no device needed; ly_plan_create runs it on every op first): known
 * kind and dtype, B >= 1, every present view a channel slice [c0, c0+c) of its buffer, kind-specific required fields.
 * Returns LY_E_ARG with ly_last_error() set.  The reference validates the same things implicitly when nn.Conv2d /
 * torch.cat raise on mismatched shapes (layers.py:51-88, 157-173). */
LY_API int32_t ly_op_validate(const ly_op* op);
LY_API int32_t ly_plan_create(const ly_op* ops, int32_t n_ops, ly_plan** out);
LY_API int32_t ly_plan_run(ly_plan* plan, float* const* ext, int32_t n_ext, int32_t img0, void* stream);
/* Same as ly_plan_run, with a CUDA event between consecutive launches: fills
 * h_ms[n_ops] (HOST array) with each op's device time and h_is_tc[n_ops] (may be
 * NULL) with 1 for tcgen05 convs.  Synchronises the stream.  Measurement only. */
LY_API int32_t ly_plan_profile(ly_plan* plan, float* const* ext, int32_t n_ext, int32_t img0, void* stream,
                               float* h_ms, int32_t* h_is_tc);
LY_API int32_t ly_plan_num_launches(const ly_plan* plan);
LY_API void ly_plan_destroy(ly_plan* plan);

/* ---- decode tail ------------------------------------------------------ */
/* `preds[l]` = level-l head tensor, NCHW fp32 `[B, 4*reg_max+nc, Hl, Wl]`
 * (exactly what `model(x)` returns / caches in `_eval_branches`).            */
typedef struct ly_levels {
  const float* preds[4];
  int32_t H[4], W[4], stride[4];
  int32_t n_levels;
  int32_t B, nc, reg_max;
  /* direct != 0: legacy `[B, 4+nc, H, W]` direct-offset layout of
   * decode_v10_predictions (postprocess.py:70-101), boxes clamped to
   * [0,clamp_w] x [0,clamp_h] when clamp_w > 0 (its `img_size`). NMS path only. */
  int32_t direct, clamp_h, clamp_w;
} ly_levels;

/* Bytes of scratch the decode entry points need for `lv` (same for both). */
LY_API int64_t ly_decode_scratch_bytes(const ly_levels* lv, int32_t max_det);

/* Replaces `decode_v10_official_topk` (postprocess.py:166-261): DFL expectation,
 * anchor decode, sigmoid, two-stage top-k with the canonical tie rule (score
 * desc, index asc).  out: `[B, k, 6]` fp32 rows [x1,y1,x2,y2,score,cls],
 * k = min(max_det, A).  Optional out_anchor/out_cls `[B,k]` int32 (may be NULL). */
LY_API int32_t ly_decode_topk(const ly_levels* lv, int32_t max_det, float* out, int32_t* out_anchor,
                       int32_t* out_cls, void* scratch, int64_t scratch_bytes, void* stream);

/* Same, with `unletterbox_coords` (utils/box_ops.py:96-124) fused into the decode epilogue: the boxes come out in each
 * SOURCE image's own pixel coordinates.  lb_meta `[B,6]` fp32 = (gain_w, gain_h, pad_left, pad_top, orig_h, orig_w), as
 * for ly_unletterbox (the caller loop tools/infer.py:110-138, tools/val.py:176-178).                             */
LY_API int32_t ly_decode_topk_lb(const ly_levels* lv, int32_t max_det, const float* lb_meta, float* out, int32_t* out_anchor,
                          int32_t* out_cls, void* scratch, int64_t scratch_bytes, void* stream);

/* Replaces `decode_v10_predictions` DFL branch (postprocess.py:103-161) +
 * `nms` (box_ops.py:49-78): per-anchor max-class candidate, `score > conf`,
 * greedy IoU NMS (suppress iff IoU > iou_thresh), first max_det survivors.
 * classwise=0: class-agnostic (what the reference code does); 1: only equal
 * labels suppress each other (export.py:145-198).  out `[B, max_det, 6]` (rows
 * past the count are zero), out_count `[B]`, out_anchor `[B,max_det]` (-1 pad). */
LY_API int32_t ly_decode_nms(const ly_levels* lv, float conf_thresh, float iou_thresh, int32_t max_det,
                      int32_t classwise, float* out, int32_t* out_count, int32_t* out_anchor,
                      void* scratch, int64_t scratch_bytes, void* stream);

/* Replaces the decode of `YOLOv10ONNXExport.forward` (models/yolov10/export.py:97-198): fixed-shape detections
 * out `[B, kmax, 6]` + num_dets `[B]` (int32).
 *   nms == 0 (export.py:126-144): top-k anchors by best class score, scores below `conf` masked, argmax class, boxes
 *            clamped to [0,img_w] x [0,img_h]; kmax = min(max_det, A); rows past num_dets hold the lowest-index masked
 *            anchors (the reference leaves the tie order of torch.topk there).
 *   nms != 0 (export.py:145-198): the `pre_topk` best (anchor, class) pairs, boxes offset per (image, class) group in fp32
 *            exactly like the reference (`img0` = global index of image 0 of this batch enters that arithmetic), ONE
 *            torchvision-style greedy NMS (unclamped areas, no epsilon, fp32 overlap > double threshold), first
 *            kmax = min(max_det, pre_topk, A*nc) survivors, rows below `conf` zeroed.
 * max_det, pre_topk <= 1024.                                                                                     */
LY_API int64_t ly_decode_export_scratch_bytes(const ly_levels* lv, int32_t max_det, int32_t pre_topk);
LY_API int32_t ly_decode_export(const ly_levels* lv, float conf, int32_t max_det, int32_t nms, double iou_thresh, int32_t pre_topk,
                                int32_t img_h, int32_t img_w, int32_t img0, float* out, int32_t* num_dets, void* scratch,
                                int64_t scratch_bytes, void* stream);

/* Replaces `leanyolo.utils.box_ops.nms` on explicit inputs: boxes `[B,N,4]`
 * xyxy fp32, scores `[B,N]`, labels `[B,N]` int32 (may be NULL when
 * classwise=0), n_valid `[B]` int32 (may be NULL = N).  keep `[B,max_keep]`
 * indices into the input order, score-descending; keep_count `[B]`.          */
LY_API int64_t ly_nms_scratch_bytes(int32_t B, int32_t N);
LY_API int32_t ly_nms(const float* boxes, const float* scores, const int32_t* labels, const int32_t* n_valid,
               int32_t B, int32_t N, float iou_thresh, int32_t max_keep, int32_t classwise,
               int32_t* keep, int32_t* keep_count, void* scratch, int64_t scratch_bytes, void* stream);

/* ---- pre / post-processing around the path -------- */
/* One source image of a letterboxed batch.  The host computes the geometry exactly as
 * `letterbox` does (utils/letterbox.py:44-80: new_w/new_h = round(orig * r), pads split by round()). */
typedef struct ly_lb_desc {
  const uint8_t* src;      /* device pointer, uint8 HWC RGB                                 */
  int64_t src_pitch;       /* bytes per source row                                          */
  int32_t src_h, src_w;
  int32_t new_h, new_w;    /* size after the resize (== src: plain copy)                    */
  int32_t top, left;       /* border added above / left; the rest of the slot is border too */
} ly_lb_desc;

/* Replaces `letterbox` (utils/letterbox.py:9-91: cv2.resize INTER_LINEAR + cv2.copyMakeBorder) and the
 * HWC->CHW transpose of tools/infer.py:112-114 for a whole batch in one launch.  `descs`: DEVICE array [B].
 * dst: uint8 `[B,3,dst_h,dst_w]` (chw != 0, what the stem kernel consumes) or `[B,dst_h,dst_w,3]`.
 * Bit-exact with cv2's 8-bit fixed-point bilinear.  fill: 3 bytes RGB (NULL = 114,114,114).          */
LY_API int32_t ly_letterbox_u8(const ly_lb_desc* descs, int32_t B, uint8_t* dst, int32_t dst_h, int32_t dst_w,
                               int32_t chw, const uint8_t* fill, void* stream);

/* Replaces `unletterbox_coords` (utils/box_ops.py:96-124) for a batch of detections, in place:
 * dets `[B,K,row]` fp32 with xyxy in the first four columns; meta `[B,6]` fp32 =
 * (gain_w, gain_h, pad_left, pad_top, orig_h, orig_w).                                               */
LY_API int32_t ly_unletterbox(float* dets, int32_t B, int32_t K, int32_t row, const float* meta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LEANYOLO_B200_H */
