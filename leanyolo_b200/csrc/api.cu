// extern "C" boundary of libleanyolo_b200.so: error state, op dispatch, whole-forward plan.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "common.cuh"

namespace ly {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
thread_local int g_reverse = 0;
static int reverse_enabled() {
  static const int v = getenv("LY_REVERSE") ? atoi(getenv("LY_REVERSE")) : 0;   // measured: no gain at batch 256 (15.11 vs 15.05 ms), kept as a switch
  return v;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {   // of the CURRENT device (cached per device)
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int n = cache[dev & 63].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("LY_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

static int32_t check_arch() {   // of the CURRENT device (cached per device: 0 unknown, 1 ok, 2 wrong architecture)
  static std::atomic<int> cache[64];
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device available (leanyolo_b200 has no CPU fallback)");
    return LY_E_CUDA;
  }
  int st = cache[dev & 63].load(std::memory_order_relaxed);
  if (st == 0) {
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
      set_error("no CUDA device available (leanyolo_b200 has no CPU fallback)");
      return LY_E_CUDA;
    }
    st = (major == 10) ? 1 : 2;
    cache[dev & 63].store(st, std::memory_order_relaxed);
  }
  if (st != 1) {
    set_error("device %d is not compute capability 10.x (this library is built for sm_100a only)", dev);
    return LY_E_ARCH;
  }
  return LY_OK;
}

static bool use_tc(const ly_op& op) { return op.dtype == LY_BF16 && op.impl != LY_IMPL_SIMT; }

static int32_t dispatch(const ly_op& op, cudaStream_t s) {
  switch (op.kind) {
    case LY_OP_STEM: return launch_stem(op, s);
    case LY_OP_CONV: {
      if (!use_tc(op)) return launch_conv_simt(op, s);
      ConvTcState* st = nullptr;
      int32_t rc = conv_tc_prepare(op, &st);
      if (rc != LY_OK) return rc;
      rc = conv_tc_launch(st, nullptr, s);
      conv_tc_free(st);
      return rc;
    }
    case LY_OP_DWPW: {
      DwPwState* st = nullptr;
      int32_t rc = dwpw_prepare(op, &st);
      if (rc != LY_OK) return rc;
      rc = dwpw_launch(st, nullptr, s);
      dwpw_free(st);
      return rc;
    }
    case LY_OP_CHAIN: {
      ChainState* st = nullptr;
      int32_t rc = chain_tc_prepare(op, &st);
      if (rc != LY_OK) return rc;
      rc = chain_tc_launch(st, nullptr, s);
      chain_tc_free(st);
      return rc;
    }
    case LY_OP_DW: return launch_dw(op, s);
    case LY_OP_POOL: return launch_pool(op, s);
    case LY_OP_UP: return launch_up(op, s);
    case LY_OP_ATTN: return launch_attn(op, s);
    case LY_OP_EXPORT: return launch_export(op, s);
    case LY_OP_IMPORT: return launch_import(op, s);
    default: set_error("unknown op kind %d", op.kind); return LY_E_ARG;
  }
}

}  // namespace ly

using namespace ly;

static_assert(sizeof(ly_chain_stage) == 104 && sizeof(ly_chain) == 672, "ly_chain layout is part of the C ABI (mirrored by ctypes in _native.py)");
static_assert(sizeof(ly_view) == 32 && sizeof(ly_op) == 272, "ly_op layout is part of the C ABI (mirrored by ctypes in _native.py)");

struct ly_plan {
  std::vector<ly_op> ops;
  std::vector<ConvTcState*> tc;   // per op (nullptr when not a tensor-core conv)
  std::vector<DwPwState*> fz;     // per op (nullptr when not a fused dw->1x1)
  std::vector<ChainState*> chn;   // per op (nullptr when not a fused chain)
  ~ly_plan() {
    for (auto* t : chn)
      if (t) chain_tc_free(t);
    for (auto* t : tc)
      if (t) conv_tc_free(t);
    for (auto* t : fz)
      if (t) dwpw_free(t);
  }
};

extern "C" {

int32_t ly_abi_version(void) { return LY_ABI_VERSION; }
const char* ly_last_error(void) { return g_err; }
int64_t ly_launch_count(void) { return g_launches.load(); }

int32_t ly_device_check(int32_t* sms) {
  int32_t rc = check_arch();
  if (rc != LY_OK) return rc;
  if (sms) *sms = sm_count();
  return LY_OK;
}

int32_t ly_launch(const ly_op* op, void* stream) {
  LY_CHECK_ARG(op != nullptr, "ly_launch: null op");
  int32_t rc = check_arch();
  if (rc != LY_OK) return rc;
  return dispatch(*op, (cudaStream_t)stream);
}

static int32_t validate_view(const char* what, const ly_view& v) {
  if (!v.ptr) return LY_OK;
  LY_CHECK_ARG(v.H > 0 && v.W > 0 && v.c > 0 && v.ctot > 0 && v.c0 >= 0 && (long long)v.c0 + v.c <= v.ctot,
               "%s view [H %d, W %d, channels %d..%d of %d] is not a channel slice of an NHWC buffer", what, v.H, v.W, v.c0,
               v.c0 + v.c, v.ctot);
  return LY_OK;
}

/* Host-side shape validation of one op description (no device needed): what every kernel family assumes before it
 * looks at an op.  The per-kernel limits (alignment, channel multiples, shared-memory fit) stay with the kernels. */
int32_t ly_op_validate(const ly_op* op) {
  LY_CHECK_ARG(op != nullptr, "null op");
  LY_CHECK_ARG(op->kind >= LY_OP_STEM && op->kind <= LY_OP_CHAIN, "unknown op kind %d", op->kind);
  LY_CHECK_ARG(op->dtype == LY_BF16 || op->dtype == LY_F32, "unknown dtype %d", op->dtype);
  LY_CHECK_ARG(op->B >= 1, "batch %d", op->B);
  int32_t rc;
  if ((rc = validate_view("src", op->src)) != LY_OK) return rc;
  if ((rc = validate_view("dst", op->dst)) != LY_OK) return rc;
  if ((rc = validate_view("res", op->res)) != LY_OK) return rc;
  if ((rc = validate_view("up", op->up)) != LY_OK) return rc;
  if (op->nchw_ctot > 0)
    LY_CHECK_ARG(op->nchw_c0 >= 0 && op->nchw_c >= 0 && (long long)op->nchw_c0 + op->nchw_c <= op->nchw_ctot,
                 "NCHW channels %d..%d of %d", op->nchw_c0, op->nchw_c0 + op->nchw_c, op->nchw_ctot);
  switch (op->kind) {
    case LY_OP_CONV:
    case LY_OP_DW:
      LY_CHECK_ARG(op->k >= 1 && (op->k & 1) && (op->stride == 1 || op->stride == 2), "k %d stride %d", op->k, op->stride);
      LY_CHECK_ARG(op->src.ptr != nullptr, "conv: no source");
      LY_CHECK_ARG(op->dst.ptr != nullptr || op->nchw != nullptr || op->ext_slot >= 0, "conv: no destination");
      if (op->res.ptr && op->dst.ptr)
        LY_CHECK_ARG(op->res.H == op->dst.H && op->res.W == op->dst.W, "shortcut %dx%d vs output %dx%d", op->res.H, op->res.W, op->dst.H, op->dst.W);
      break;
    case LY_OP_DWPW:
      LY_CHECK_ARG(op->pre_k == 3 && op->pre_w && op->pre_bias && op->src.ptr, "dwpw: depthwise stage missing or not 3x3");
      break;
    case LY_OP_ATTN:
      LY_CHECK_ARG(op->nh >= 1 && op->kdp >= 1 && op->hd >= 1 && op->src.ptr && op->dst.ptr, "attention: heads %d key dim %d head dim %d", op->nh, op->kdp, op->hd);
      break;
    case LY_OP_CHAIN:
      LY_CHECK_ARG(op->chain != nullptr, "chain: no description");
      LY_CHECK_ARG(op->chain->n_stages >= 1 && op->chain->n_stages <= LY_CHAIN_MAX_STAGES && op->chain->n_regions >= 1 &&
                   op->chain->n_regions <= LY_CHAIN_MAX_REGIONS, "chain: %d stages, %d regions", op->chain->n_stages, op->chain->n_regions);
      break;
    case LY_OP_POOL:
    case LY_OP_UP:
      LY_CHECK_ARG(op->src.ptr && op->dst.ptr, "pool / upsample: null view");
      break;
    default: break;
  }
  return LY_OK;
}

int32_t ly_plan_create(const ly_op* ops, int32_t n_ops, ly_plan** out) {
  LY_CHECK_ARG(ops && n_ops > 0 && out, "ly_plan_create: bad arguments");
  int32_t rc = check_arch();
  if (rc != LY_OK) return rc;
  for (int i = 0; i < n_ops; ++i) {
    if (ly_op_validate(&ops[i]) != LY_OK) {
      char msg[400];
      snprintf(msg, sizeof(msg), "%.380s", g_err);
      set_error("plan op %d: %s", i, msg);
      return LY_E_ARG;
    }
  }
  ly_plan* pl = new ly_plan();
  pl->ops.assign(ops, ops + n_ops);
  pl->tc.assign(n_ops, nullptr);
  pl->fz.assign(n_ops, nullptr);
  pl->chn.assign(n_ops, nullptr);
  for (int i = 0; i < n_ops; ++i) {
    const ly_op& op = pl->ops[i];
    g_reverse = reverse_enabled() ? (i & 1) : 0;
    if ((op.kind == LY_OP_CONV && use_tc(op)) || op.kind == LY_OP_DWPW || op.kind == LY_OP_CHAIN) {
      ly_op tmp = op;
      if (tmp.ext_slot >= 0 && !tmp.nchw) tmp.nchw = reinterpret_cast<float*>(16);  // placeholder: real pointer comes at run time
      rc = op.kind == LY_OP_CHAIN ? chain_tc_prepare(tmp, &pl->chn[i])
                                  : (op.kind == LY_OP_DWPW ? dwpw_prepare(tmp, &pl->fz[i]) : conv_tc_prepare(tmp, &pl->tc[i]));
      pl->ops[i].chain = nullptr;   // the prepared state owns a copy; the caller's description may go away
      if (rc != LY_OK) {
        char msg[400];
        snprintf(msg, sizeof(msg), "%s", g_err);
        set_error("plan op %d: %s", i, msg);
        delete pl;
        return rc;
      }
    }
  }
  g_reverse = 0;
  *out = pl;
  return LY_OK;
}

static int32_t plan_run_impl(ly_plan* pl, float* const* ext, int32_t n_ext, int32_t img0, cudaStream_t s,
                             cudaEvent_t* ev) {
  for (size_t i = 0; i < pl->ops.size(); ++i) {
    if (ev) cudaEventRecord(ev[i], s);
    const ly_op& op = pl->ops[i];
    float* nchw = op.nchw;
    if (op.ext_slot >= 0) {
      LY_CHECK_ARG(ext && op.ext_slot < n_ext && ext[op.ext_slot], "ly_plan_run: op %d needs ext[%d]", (int)i, op.ext_slot);
      // offset by img0 whole images of the external NCHW tensor
      long long per_img;
      if (op.kind == LY_OP_STEM && op.impl == LY_STEM_IN_LB) per_img = (long long)sizeof(ly_lb_desc);
      else if (op.kind == LY_OP_STEM) per_img = 3LL * (2 * op.dst.H) * (2 * op.dst.W);
      else if (op.kind == LY_OP_IMPORT) per_img = (long long)op.nchw_ctot * op.dst.H * op.dst.W;
      else if (op.kind == LY_OP_EXPORT) per_img = (long long)op.nchw_ctot * op.src.H * op.src.W;
      else per_img = (long long)op.nchw_ctot * (op.src.H / op.stride) * (op.src.W / op.stride);
      const long long esz = (op.kind == LY_OP_STEM && (op.impl == LY_STEM_IN_U8 || op.impl == LY_STEM_IN_LB)) ? 1 : 4;
      nchw = reinterpret_cast<float*>(reinterpret_cast<char*>(ext[op.ext_slot]) + (long long)img0 * per_img * esz);
    }
    int32_t rc;
    g_reverse = reverse_enabled() ? (int)(i & 1) : 0;
    if (pl->tc[i]) {
      rc = conv_tc_launch(pl->tc[i], op.ext_slot >= 0 ? nchw : nullptr, s);
    } else if (pl->fz[i]) {
      rc = dwpw_launch(pl->fz[i], op.ext_slot >= 0 ? nchw : nullptr, s);
    } else if (pl->chn[i]) {
      rc = chain_tc_launch(pl->chn[i], op.ext_slot >= 0 ? nchw : nullptr, s);
    } else {
      ly_op tmp = op;
      tmp.nchw = nchw;
      rc = dispatch(tmp, s);
    }
    if (rc != LY_OK) {
      char msg[400];
      snprintf(msg, sizeof(msg), "%s", g_err);
      set_error("plan op %d (kind %d): %s", (int)i, op.kind, msg);
      return rc;
    }
  }
  g_reverse = 0;
  if (ev) cudaEventRecord(ev[pl->ops.size()], s);
  return LY_OK;
}

int32_t ly_letterbox_u8(const ly_lb_desc* descs, int32_t B, uint8_t* dst, int32_t dst_h, int32_t dst_w, int32_t chw,
                        const uint8_t* fill, void* stream) {
  int32_t rc = check_arch();
  if (rc != LY_OK) return rc;
  return launch_letterbox(descs, B, dst, dst_h, dst_w, chw, fill, (cudaStream_t)stream);
}

int32_t ly_unletterbox(float* dets, int32_t B, int32_t K, int32_t row, const float* meta, void* stream) {
  int32_t rc = check_arch();
  if (rc != LY_OK) return rc;
  return launch_unletterbox(dets, B, K, row, meta, (cudaStream_t)stream);
}

int32_t ly_plan_run(ly_plan* pl, float* const* ext, int32_t n_ext, int32_t img0, void* stream) {
  LY_CHECK_ARG(pl != nullptr, "ly_plan_run: null plan");
  return plan_run_impl(pl, ext, n_ext, img0, (cudaStream_t)stream, nullptr);
}

int32_t ly_plan_profile(ly_plan* pl, float* const* ext, int32_t n_ext, int32_t img0, void* stream, float* h_ms,
                        int32_t* h_is_tc) {
  LY_CHECK_ARG(pl != nullptr && h_ms != nullptr, "ly_plan_profile: null argument");
  const size_t n = pl->ops.size();
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) LY_CUDA(cudaEventCreate(&e));
  int32_t rc = plan_run_impl(pl, ext, n_ext, img0, (cudaStream_t)stream, ev.data());
  if (rc == LY_OK) {
    cudaError_t e = cudaEventSynchronize(ev[n]);
    if (e != cudaSuccess) { set_error("ly_plan_profile: %s", cudaGetErrorString(e)); rc = LY_E_CUDA; }
  }
  if (rc == LY_OK)
    for (size_t i = 0; i < n; ++i) {
      cudaEventElapsedTime(&h_ms[i], ev[i], ev[i + 1]);
      if (h_is_tc) h_is_tc[i] = (pl->tc[i] || pl->fz[i] || pl->chn[i]) ? 1 : 0;
    }
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

int32_t ly_plan_num_launches(const ly_plan* pl) { return pl ? (int32_t)pl->ops.size() : 0; }

void ly_plan_destroy(ly_plan* pl) { delete pl; }

}  // extern "C"
