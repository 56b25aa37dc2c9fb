// PSA attention core (leanyolo/models/yolov10/layers.py:369-378), fused flash-style:
//   out[n, h*hd + d] = sum_m softmax_m( (q_n . k_m) * scale ) * v_m[d]
// The N x N score matrix never reaches HBM (the reference materialises 4 x 400^2 fp32 per
// image).  Small-N (400 tokens @640^2, 1600 @1280^2), so this is a latency/bandwidth
// kernel: one thread per query keeps q, the running max/sum and the output row in
// registers; K/V blocks are staged in shared memory and read as warp-wide broadcasts.
//
// Channel layout of the qkv buffer (re-ordered at weight-pack time, modules.py):
//   [ q: nh x kdp | k: nh x kdp | v: nh x hd ]   (kdp = key_dim padded to 8, pad = zeros)
#include "common.cuh"

namespace ly {

namespace {

constexpr int QB = 128;  // queries (= threads) per CTA
constexpr int KB = 32;   // keys per shared-memory block

template <typename T, int KDP, int HD>
__global__ void __launch_bounds__(QB)
attn_kernel(const T* __restrict__ qkv, int N, int sCtot, int sC0, int nh, T* __restrict__ out, int dCtot, int dC0,
            float scale) {
  __shared__ __align__(16) float Ks[KB][KDP];
  __shared__ __align__(16) float Vs[KB][HD];
  constexpr int V = Elem<T>::kVec;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n = blockIdx.x * QB + threadIdx.x;
  const bool active = n < N;
  const T* base = qkv + (long long)b * N * sCtot + sC0;
  const int qoff = h * KDP, koff = nh * KDP + h * KDP, voff = 2 * nh * KDP + h * HD;

  float q[KDP];
  if (active) {
#pragma unroll
    for (int d = 0; d < KDP; d += V) {
      float t[V];
      load_vec<T>(base + (long long)n * sCtot + qoff + d, t);
#pragma unroll
      for (int j = 0; j < V; ++j) q[d + j] = t[j] * scale;
    }
  } else {
#pragma unroll
    for (int d = 0; d < KDP; ++d) q[d] = 0.f;
  }
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  float mx = -INFINITY, l = 0.f;

  for (int m0 = 0; m0 < N; m0 += KB) {
    __syncthreads();
    // stage K/V block (zero rows past N are masked below)
    for (int i = threadIdx.x; i < KB * (KDP / V); i += QB) {
      const int r = i / (KDP / V), d = (i - r * (KDP / V)) * V;
      float t[V];
      if (m0 + r < N) load_vec<T>(base + (long long)(m0 + r) * sCtot + koff + d, t);
      else {
#pragma unroll
        for (int j = 0; j < V; ++j) t[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < V; ++j) Ks[r][d + j] = t[j];
    }
    for (int i = threadIdx.x; i < KB * (HD / V); i += QB) {
      const int r = i / (HD / V), d = (i - r * (HD / V)) * V;
      float t[V];
      if (m0 + r < N) load_vec<T>(base + (long long)(m0 + r) * sCtot + voff + d, t);
      else {
#pragma unroll
        for (int j = 0; j < V; ++j) t[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < V; ++j) Vs[r][d + j] = t[j];
    }
    __syncthreads();
    const int nk = min(KB, N - m0);
    float s[KB];
    float bm = -INFINITY;
#pragma unroll
    for (int r = 0; r < KB; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < KDP; d += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(&Ks[r][d]);
        acc = fmaf(q[d], kv.x, acc);
        acc = fmaf(q[d + 1], kv.y, acc);
        acc = fmaf(q[d + 2], kv.z, acc);
        acc = fmaf(q[d + 3], kv.w, acc);
      }
      s[r] = r < nk ? acc : -INFINITY;
      bm = fmaxf(bm, s[r]);
    }
    const float nm = fmaxf(mx, bm);
    const float corr = sizeof(T) == 4 ? expf(mx - nm) : __expf(mx - nm);  // mx=-inf first time -> 0
    l *= corr;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] *= corr;
    mx = nm;
#pragma unroll
    for (int r = 0; r < KB; ++r) {
      const float p = sizeof(T) == 4 ? expf(s[r] - mx) : __expf(s[r] - mx);  // masked keys: exp(-inf) = 0
      l += p;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&Vs[r][d]);
        o[d] = fmaf(p, vv.x, o[d]);
        o[d + 1] = fmaf(p, vv.y, o[d + 1]);
        o[d + 2] = fmaf(p, vv.z, o[d + 2]);
        o[d + 3] = fmaf(p, vv.w, o[d + 3]);
      }
    }
  }
  if (!active) return;
  const float inv = 1.0f / l;
  T* op = out + ((long long)b * N + n) * dCtot + dC0 + h * HD;
#pragma unroll
  for (int d = 0; d < HD; d += V) {
    float t[V];
#pragma unroll
    for (int j = 0; j < V; ++j) t[j] = o[d + j] * inv;
    store_vec<T>(op + d, t);
  }
}

template <typename T, int KDP, int HD>
int32_t run(const ly_op& op, cudaStream_t s) {
  const int N = op.src.H * op.src.W;
  dim3 grid((N + QB - 1) / QB, op.nh, op.B);
  attn_kernel<T, KDP, HD><<<grid, QB, 0, s>>>((const T*)op.src.ptr, N, op.src.ctot, op.src.c0, op.nh, (T*)op.dst.ptr,
                                              op.dst.ctot, op.dst.c0, op.scale);
  return post_launch("psa_attention");
}

}  // namespace

int32_t launch_attn(const ly_op& op, cudaStream_t s) {
  LY_CHECK_ARG(op.src.ptr && op.dst.ptr, "attention: null pointer");
  LY_CHECK_ARG(op.src.c >= 2 * op.nh * op.kdp + op.nh * op.hd && op.dst.c >= op.nh * op.hd, "attention: channel mismatch");
  LY_CHECK_ARG(op.src.c0 % 8 == 0 && op.dst.c0 % 8 == 0 && op.src.ctot % 8 == 0 && op.dst.ctot % 8 == 0,
               "attention: views must be 16-byte aligned");
  const bool f32 = op.dtype == LY_F32;
  if (op.kdp == 32 && op.hd == 64)
    return f32 ? run<float, 32, 64>(op, s) : run<__nv_bfloat16, 32, 64>(op, s);
  if (op.kdp == 40 && op.hd == 72)
    return f32 ? run<float, 40, 72>(op, s) : run<__nv_bfloat16, 40, 72>(op, s);
  if (op.kdp == 16 && op.hd == 32)
    return f32 ? run<float, 16, 32>(op, s) : run<__nv_bfloat16, 16, 32>(op, s);
  LY_CHECK_ARG(false, "attention: unsupported (key_dim_pad=%d, head_dim=%d); built: (32,64) (40,72) (16,32)", op.kdp, op.hd);
}

}  // namespace ly
