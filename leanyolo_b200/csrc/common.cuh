// Shared helpers for the leanyolo_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include "../../include/leanyolo_b200.h"

namespace ly {

// ---- error plumbing ---------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define LY_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ly::set_error(__VA_ARGS__);               \
      return LY_E_ARG;                          \
    }                                           \
  } while (0)

#define LY_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      ly::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return LY_E_CUDA;                                                                 \
    }                                                                                   \
  } while (0)

inline int32_t post_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return LY_E_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return LY_OK;
}

// ---- per-op launchers (defined in the .cu files) ----------------------------
int32_t launch_stem(const ly_op& op, cudaStream_t s);
int32_t launch_conv_simt(const ly_op& op, cudaStream_t s);
int32_t launch_dw(const ly_op& op, cudaStream_t s);
int32_t launch_dw_tma(const ly_op& op, cudaStream_t s);   // bf16, TMA-staged (dw_tma.cu)
bool dw_tma_supported(const ly_op& op);
int32_t launch_pool(const ly_op& op, cudaStream_t s);
int32_t launch_up(const ly_op& op, cudaStream_t s);
int32_t launch_attn(const ly_op& op, cudaStream_t s);
int32_t launch_export(const ly_op& op, cudaStream_t s);
int32_t launch_import(const ly_op& op, cudaStream_t s);
int32_t launch_letterbox(const ly_lb_desc* descs, int32_t B, uint8_t* dst, int32_t dst_h, int32_t dst_w, int32_t chw,
                         const uint8_t* fill, cudaStream_t s);
int32_t launch_unletterbox(float* dets, int32_t B, int32_t K, int32_t row, const float* meta, cudaStream_t s);

// tcgen05 implicit-GEMM conv: tensor maps are encoded once (prepare) and reused.
struct ConvTcState;
int32_t conv_tc_prepare(const ly_op& op, ConvTcState** out);
int32_t conv_tc_launch(const ConvTcState* st, float* nchw_override, cudaStream_t s);
void conv_tc_free(ConvTcState* st);
bool conv_tc_supported(const ly_op& op);

// fused depthwise 3x3 -> 1x1 (dwpw_tc.cu)
struct DwPwState;
int32_t dwpw_prepare(const ly_op& op, DwPwState** out);
int32_t dwpw_launch(const DwPwState* st, float* nchw_override, cudaStream_t s);
void dwpw_free(DwPwState* st);
bool dwpw_supported(const ly_op& op);
// ... with the depthwise stage on the tensor cores too (dwpw_mma.cu; C <= 128, Cout <= 128): chosen by dwpw_prepare
struct DwPwMmaState;
int32_t dwpw_mma_prepare(const ly_op& op, DwPwMmaState** out);
int32_t dwpw_mma_launch(const DwPwMmaState* st, float* nchw_override, cudaStream_t s);
void dwpw_mma_free(DwPwMmaState* st);
bool dwpw_mma_supported(const ly_op& op);

// fused chain of dense conv stages with shared-memory intermediates (chain_tc.cu)
struct ChainState;
int32_t chain_tc_prepare(const ly_op& op, ChainState** out);
int32_t chain_tc_launch(const ChainState* st, float* nchw_override, cudaStream_t s);
void chain_tc_free(ChainState* st);
bool chain_tc_supported(const ly_op& op);
// 3x3 -> 1x1 tail of a regression stack as a back-to-back GEMM (conv_b2b.cu); chain_tc_prepare prefers it when it applies
struct B2bState;
bool conv_b2b_supported(const ly_op& op);
int32_t conv_b2b_prepare(const ly_op& op, B2bState** out);
int32_t conv_b2b_launch(const B2bState* st, float* nchw_override, cudaStream_t s);
void conv_b2b_free(B2bState* st);

int sm_count();
// true exactly once per (flag, current device): per-device one-time setup such as cudaFuncSetAttribute (the
// Python API allows models on several GPUs of one process; a process-wide flag would skip the second device)
inline bool first_on_device(std::atomic<unsigned long long>& devs) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  const unsigned long long bit = 1ull << (dev & 63);
  return (devs.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}

// Tile traversal direction of the op being prepared / launched (set by the plan executor: ops
// alternate).  A layer that walks its tiles in the opposite order to its producer starts on the
// part of the producer's output that is still in the 126 MB L2 (the tensors are 100-800 MB, so a
// same-direction walk always starts on lines that were evicted long ago).
extern thread_local int g_reverse;

// Programmatic dependent launch: every kernel of the forward is launched with the
// programmatic-stream-serialization attribute, so its CTAs are scheduled (and run their
// prologue: barrier init, TMEM allocation, tensor-map prefetch) while the previous kernel
// drains, and block in pdl_wait() until that kernel has completed and flushed.  Every
// kernel launched this way MUST execute pdl_wait() before touching activations and before
// it exits (that is what makes completion transitive along the stream).  LY_PDL=0 disables.
bool pdl_enabled();

// ---- device helpers ---------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... Exp, typename... Act>
inline cudaError_t launch_k(void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Act&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at;
  memset(&at, 0, sizeof(at));
  at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = &at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Exp>(args)...);
}

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// SiLU(x) = x*sigmoid(x) = h + h*tanh(h), h = x/2: one MUFU op (tanh.approx, rel. err 2^-11),
// used by the bf16 hot-path epilogues (the result is rounded to bf16 = 2^-9 anyway)
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
// check-mode SiLU: full-precision expf and IEEE division (1e-4 parity vs the fp32 oracle)
__device__ __forceinline__ float silu_precise(float x) { return x / (1.0f + expf(-x)); }

template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int kVec = 4;  // elements per 16-byte vector
  __device__ static __forceinline__ float to_f(float v) { return v; }
  __device__ static __forceinline__ float from_f(float v) { return v; }
  __device__ static __forceinline__ float act(float v) { return silu_precise(v); }
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int kVec = 8;
  __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  __device__ static __forceinline__ float act(float v) { return silu_f(v); }
};

// 16-byte vector of T <-> float[kVec]
template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float* out);
template <>
__device__ __forceinline__ void load_vec<float>(const float* p, float* out) {
  float4 v = *reinterpret_cast<const float4*>(p);
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <>
__device__ __forceinline__ void load_vec<__nv_bfloat16>(const __nv_bfloat16* p, float* out) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    out[2 * i] = f.x;
    out[2 * i + 1] = f.y;
  }
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const float* in);
template <>
__device__ __forceinline__ void store_vec<float>(float* p, const float* in) {
  *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
}
template <>
__device__ __forceinline__ void store_vec<__nv_bfloat16>(__nv_bfloat16* p, const float* in) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}

// 16 bf16 channels (32 bytes, 32-byte aligned) in ONE 256-bit store (sm_100 STG.256).  The conv epilogues
// write 32 bytes per thread at a pixel-pitch stride, i.e. every lane is its own L1 request: the L1/LSU
// request rate, not DRAM, bounded the thin layers (ncu: l1tex 68 % busy at 58 % DRAM), and two 16-byte
// stores are two requests.
__device__ __forceinline__ void store_bf16x16(__nv_bfloat16* p, const float* in) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
               "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}

#endif  // __CUDACC__

}  // namespace ly
