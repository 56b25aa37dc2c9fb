// Bandwidth-bound kernels of the YOLOv10 forward: stem conv, depthwise convs (3x3 / 7x7,
// stride 1 / 2), SPPF max-pool pyramid, nearest x2 upsample into a concat slice and the
// NHWC<->NCHW boundary converters.  All NHWC with 16-byte channel vectors; templated on
// the storage type (bf16 hot path, fp32 check mode).
#include "common.cuh"

namespace ly {

namespace {

// ---------------------------------------------------------------------------------------
// Stem: backbone.cv0 = 3x3 stride-2 conv on the user's NCHW fp32 image (backbone.py:68),
// with x' = (x - sub) / div applied while loading (yolov10s.py:107-112; zero padding is
// applied to the normalised image).  K = 27 -> arithmetic intensity ~20 FLOP/B: memory /
// CUDA-core bound, not a tensor-core shape.
// Tile: 4 output rows x 32 output cols per CTA (128 threads, one output pixel each).
// ---------------------------------------------------------------------------------------
constexpr int ST_TW = 64, ST_TH = 4;       // output tile per CTA; each thread owns 2 adjacent pixels
constexpr int ST_IW = 2 * ST_TW + 1, ST_IH = 2 * ST_TH + 1;
constexpr int ST_THREADS = ST_TW / 2 * ST_TH;

template <typename T, typename TIn>
__global__ void __launch_bounds__(ST_THREADS)
stem_kernel(const TIn* __restrict__ x, int H, int W, T* __restrict__ dst, int dCtot, int dC0, int Cpad,
            const float* __restrict__ w, const float* __restrict__ bias,
            float s0, float s1, float s2, float d0, float d1, float d2) {
  extern __shared__ __align__(16) float smem[];
  float* ws = smem;                           // [27][Cpad]  (transposed from [Cpad][27])
  float* bs = ws + 27 * Cpad;                 // [Cpad]
  float* tile = bs + Cpad;                    // [3][ST_IH][ST_IW]
  const int Ho = H / 2, Wo = W / 2;
  const int b = blockIdx.z;
  const int ho0 = blockIdx.y * ST_TH, wo0 = blockIdx.x * ST_TW;
  const int tid = threadIdx.x;
  const float sub[3] = {s0, s1, s2}, div[3] = {d0, d1, d2};

  for (int i = tid; i < 27 * Cpad; i += blockDim.x) {
    int co = i / 27, t = i - co * 27;
    ws[t * Cpad + co] = w[i];
  }
  for (int i = tid; i < Cpad; i += blockDim.x) bs[i] = bias[i];
  const int hi0 = 2 * ho0 - 1, wi0 = 2 * wo0 - 1;
  {
    // 27 (channel, row) lines of ST_IW floats; a warp takes a line, a lane 4-5 columns of it.
    // All loads of a thread are issued before any is consumed (latency-bound otherwise).
    constexpr int NWARP = ST_THREADS / 32, LPW = (3 * ST_IH + NWARP - 1) / NWARP, CPL = (ST_IW + 31) / 32;
    const int warp = tid >> 5, lane = tid & 31;
    float v[LPW][CPL];
#pragma unroll
    for (int l = 0; l < LPW; ++l) {
      const int line = warp + l * NWARP;
      const int c = line / ST_IH, iy = line - c * ST_IH;
      const int hi = hi0 + iy;
      const bool rok = line < 3 * ST_IH && hi >= 0 && hi < H;
      const TIn* rp = x + (((long long)b * 3 + (rok ? c : 0)) * H + (rok ? hi : 0)) * W;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int ix = lane + 32 * k, wi = wi0 + ix;
        v[l][k] = (rok && ix < ST_IW && wi >= 0 && wi < W) ? (float)__ldg(rp + wi) : 0.f;
      }
    }
#pragma unroll
    for (int l = 0; l < LPW; ++l) {
      const int line = warp + l * NWARP;
      if (line >= 3 * ST_IH) continue;
      const int c = line / ST_IH, iy = line - c * ST_IH;
      const int hi = hi0 + iy;
      const bool rok = hi >= 0 && hi < H;
      const float sc = c == 0 ? sub[0] : (c == 1 ? sub[1] : sub[2]);
      const float dc = c == 0 ? div[0] : (c == 1 ? div[1] : div[2]);
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int ix = lane + 32 * k, wi = wi0 + ix;
        if (ix < ST_IW) tile[line * ST_IW + ix] = (rok && wi >= 0 && wi < W) ? (v[l][k] - sc) / dc : 0.f;
      }
    }
  }
  __syncthreads();

  const int ty = tid / (ST_TW / 2), tx = (tid - ty * (ST_TW / 2)) * 2;
  const int ho = ho0 + ty, wo = wo0 + tx;
  if (ho >= Ho || wo >= Wo) return;
  float in[2][27];  // [pixel][ky][kx][ci] to match the packed weight order
#pragma unroll
  for (int px = 0; px < 2; ++px)
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          in[px][(ky * 3 + kx) * 3 + c] = tile[(c * ST_IH + 2 * ty + ky) * ST_IW + 2 * (tx + px) + kx];

  T* out = dst + (((long long)b * Ho + ho) * Wo + wo) * dCtot + dC0;
  const bool second = wo + 1 < Wo;
  constexpr int V = Elem<T>::kVec;
  for (int co = 0; co < Cpad; co += V) {
    float a0[V], a1[V];
#pragma unroll
    for (int j = 0; j < V; ++j) a0[j] = a1[j] = bs[co + j];
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      float wv[V];
#pragma unroll
      for (int j = 0; j < V; j += 4) {
        const float4 q = *reinterpret_cast<const float4*>(ws + t * Cpad + co + j);
        wv[j] = q.x; wv[j + 1] = q.y; wv[j + 2] = q.z; wv[j + 3] = q.w;
      }
#pragma unroll
      for (int j = 0; j < V; ++j) {
        a0[j] = fmaf(in[0][t], wv[j], a0[j]);
        a1[j] = fmaf(in[1][t], wv[j], a1[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) { a0[j] = Elem<T>::act(a0[j]); a1[j] = Elem<T>::act(a1[j]); }
    store_vec<T>(out + co, a0);
    if (second) store_vec<T>(out + dCtot + co, a1);
  }
}

// ---------------------------------------------------------------------------------------
// Depthwise k x k conv (+bias, optional SiLU, optional residual added after the activation).
// One thread = a TH x TW patch of output pixels x one 16-byte channel vector: every input
// vector is loaded once per patch and reused by all the outputs that see it (k=3,s=1: 24
// loads for 8 outputs instead of 72), weights are fetched once per tap per patch.
// Consecutive threads walk the channel vectors, so every access is a coalesced 16-byte
// vector; accumulation in fp32.
// ---------------------------------------------------------------------------------------
template <typename T, int K, int S, int TH, int TW>
__global__ void __launch_bounds__(128)
dw_kernel(const T* __restrict__ src, int sH, int sW, int sCtot, int sC0,
          T* __restrict__ dst, int dCtot, int dC0, const T* res, int rCtot, int rC0,
          const T* __restrict__ w, const float* __restrict__ bias,
          int C, int Ho, int Wo, int act, int px, int py, long long total) {
  constexpr int V = Elem<T>::kVec;
  constexpr int IH = (TH - 1) * S + K, IW = (TW - 1) * S + K;
  const int cv = C / V;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * V;
    long long p = idx / cv;
    const int ox0 = (int)(p % px) * TW;
    p /= px;
    const int oy0 = (int)(p % py) * TH;
    const int b = (int)(p / py);
    float acc[TH][TW][V];
    {
      float bv[V];
#pragma unroll
      for (int j = 0; j < V; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(bias + c + j);
        bv[j] = t.x; bv[j + 1] = t.y; bv[j + 2] = t.z; bv[j + 3] = t.w;
      }
#pragma unroll
      for (int y = 0; y < TH; ++y)
#pragma unroll
        for (int x = 0; x < TW; ++x)
#pragma unroll
          for (int j = 0; j < V; ++j) acc[y][x][j] = bv[j];
    }
    const int hi0 = oy0 * S - K / 2, wi0 = ox0 * S - K / 2;
    const T* sb = src + (long long)b * sH * sW * sCtot + sC0 + c;
#pragma unroll
    for (int iy = 0; iy < IH; ++iy) {
      const int hi = hi0 + iy;
      if (hi < 0 || hi >= sH) continue;
      float in[IW][V];
#pragma unroll
      for (int ix = 0; ix < IW; ++ix) {
        const int wi = wi0 + ix;
        if (wi >= 0 && wi < sW) {
          load_vec<T>(sb + ((long long)hi * sW + wi) * sCtot, in[ix]);
        } else {
#pragma unroll
          for (int j = 0; j < V; ++j) in[ix][j] = 0.f;
        }
      }
#pragma unroll
      for (int y = 0; y < TH; ++y) {
        const int ky = iy - y * S;
        if (ky < 0 || ky >= K) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          float wv[V];
          load_vec<T>(w + (ky * K + kx) * C + c, wv);
#pragma unroll
          for (int x = 0; x < TW; ++x)
#pragma unroll
            for (int j = 0; j < V; ++j) acc[y][x][j] = fmaf(in[x * S + kx][j], wv[j], acc[y][x][j]);
        }
      }
    }
#pragma unroll
    for (int y = 0; y < TH; ++y) {
      const int ho = oy0 + y;
      if (ho >= Ho) continue;
#pragma unroll
      for (int x = 0; x < TW; ++x) {
        const int wo = ox0 + x;
        if (wo >= Wo) continue;
        const long long opix = ((long long)b * Ho + ho) * Wo + wo;
        if (act) {
#pragma unroll
          for (int j = 0; j < V; ++j) acc[y][x][j] = Elem<T>::act(acc[y][x][j]);
        }
        if (res) {
          float rv[V];
          load_vec<T>(res + opix * rCtot + rC0 + c, rv);
#pragma unroll
          for (int j = 0; j < V; ++j) acc[y][x][j] += rv[j];
        }
        store_vec<T>(dst + opix * dCtot + dC0 + c, acc[y][x]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// SPPF pyramid: y1 = pool5(x), y2 = pool5(y1), y3 = pool5(y2) with -inf padding
// (layers.py:210-217).  One CTA = one image x one 16-byte channel vector: the H x W plane
// lives in shared memory and each of the three chained pools is a separable row-max /
// column-max pass over it.  Reads channels [0,c) of the concat buffer and writes [c,2c),
// [2c,3c), [3c,4c) of the same buffer.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
pool_kernel(T* buf, int H, int W, int Ctot, int C0, int C) {
  constexpr int V = Elem<T>::kVec;
  extern __shared__ float psm[];            // [2][H*W][V]
  const int HW = H * W;
  float* cur = psm;
  float* tmp = psm + (size_t)HW * V;
  const int c = blockIdx.x * V;
  const int b = blockIdx.y;
  T* base = buf + (long long)b * HW * Ctot + C0 + c;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    float v[V];
    load_vec<T>(base + (long long)p * Ctot, v);
#pragma unroll
    for (int j = 0; j < V; ++j) cur[p * V + j] = v[j];
  }
  __syncthreads();
  for (int stage = 1; stage <= 3; ++stage) {
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {   // row max -> tmp
      const int y = p / W, x = p - y * W;
      float m[V];
#pragma unroll
      for (int j = 0; j < V; ++j) m[j] = -INFINITY;
      for (int dx = -2; dx <= 2; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
#pragma unroll
        for (int j = 0; j < V; ++j) m[j] = fmaxf(m[j], cur[(y * W + xx) * V + j]);
      }
#pragma unroll
      for (int j = 0; j < V; ++j) tmp[p * V + j] = m[j];
    }
    __syncthreads();
    for (int p = threadIdx.x; p < HW; p += blockDim.x) {   // column max -> cur (and out)
      const int y = p / W, x = p - y * W;
      float m[V];
#pragma unroll
      for (int j = 0; j < V; ++j) m[j] = -INFINITY;
      for (int dy = -2; dy <= 2; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int j = 0; j < V; ++j) m[j] = fmaxf(m[j], tmp[(yy * W + x) * V + j]);
      }
      store_vec<T>(base + (long long)p * Ctot + stage * C, m);
      // cur is only read by the row pass, which finished before the barrier above
#pragma unroll
      for (int j = 0; j < V; ++j) cur[p * V + j] = m[j];
    }
    __syncthreads();
  }
}

// nearest x2: dst(b, y, x, :) = src(b, y/2, x/2, :)   (layers.py:240)
template <typename T>
__global__ void __launch_bounds__(256)
up_kernel(const T* __restrict__ src, int sH, int sW, int sCtot, int sC0, T* __restrict__ dst, int dCtot, int dC0,
          int C, long long total) {
  constexpr int V = Elem<T>::kVec;
  const int cv = C / V;
  const int dH = 2 * sH, dW = 2 * sW;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * V;
    long long p = idx / cv;
    const int x = (int)(p % dW);
    p /= dW;
    const int y = (int)(p % dH);
    const int b = (int)(p / dH);
    const uint4 v = *reinterpret_cast<const uint4*>(src + (((long long)b * sH + (y >> 1)) * sW + (x >> 1)) * sCtot + sC0 + c);
    *reinterpret_cast<uint4*>(dst + (((long long)b * dH + y) * dW + x) * dCtot + dC0 + c) = v;
  }
}

// NHWC storage -> NCHW fp32 (pixel index fastest => coalesced stores)
template <typename T>
__global__ void __launch_bounds__(256)
export_kernel(const T* __restrict__ src, int HW, int sCtot, int sC0, float* __restrict__ out, int nCtot, int nC0,
              int nC, long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx % HW);
    long long r = idx / HW;
    const int c = (int)(r % nC);
    const long long b = r / nC;
    out[(b * nCtot + nC0 + c) * HW + p] = Elem<T>::to_f(src[(b * HW + p) * sCtot + sC0 + c]);
  }
}

// NCHW fp32 -> NHWC storage (channel fastest => coalesced stores); channels >= nC are zeroed
template <typename T>
__global__ void __launch_bounds__(256)
import_kernel(const float* __restrict__ in, int HW, int nCtot, int nC0, int nC, T* __restrict__ dst, int dCtot,
              int dC0, int C, long long total) {
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long long r = idx / C;
    const int p = (int)(r % HW);
    const long long b = r / HW;
    const float v = c < nC ? in[(b * nCtot + nC0 + c) * HW + p] : 0.f;
    dst[(b * HW + p) * dCtot + dC0 + c] = Elem<T>::from_f(v);
  }
}

inline unsigned grid_for(long long total, int block = 256) {
  long long g = (total + block - 1) / block;
  const long long cap = (long long)sm_count() * 32;
  return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

template <typename T>
bool aligned16(const ly_view& v) {
  constexpr int V = 16 / sizeof(T);
  return (reinterpret_cast<uintptr_t>(v.ptr) % 16 == 0) && v.ctot % V == 0 && v.c0 % V == 0 && v.c % V == 0;
}

}  // namespace

// ------------------------------------------------------------------------------- launchers
int32_t launch_stem(const ly_op& op, cudaStream_t s) {
  LY_CHECK_ARG(op.nchw && op.dst.ptr && op.w && op.bias, "stem: null pointer");
  const int H = 2 * op.dst.H, W = 2 * op.dst.W;
  const int Cpad = op.dst.c;
  LY_CHECK_ARG(Cpad % 8 == 0 && Cpad <= 256, "stem: Cout_pad must be a multiple of 8 and <= 256");
  LY_CHECK_ARG(op.dst.W % 2 == 0, "stem: output width must be even");
  dim3 grid((op.dst.W + ST_TW - 1) / ST_TW, (op.dst.H + ST_TH - 1) / ST_TH, op.B);
  size_t smem = (3 * ST_IH * ST_IW + 28 * Cpad) * sizeof(float);
#define LY_STEM(T, TIN)                                                                                              \
  stem_kernel<T, TIN><<<grid, ST_THREADS, smem, s>>>((const TIN*)op.nchw, H, W, (T*)op.dst.ptr, op.dst.ctot, op.dst.c0, \
                                                     Cpad, (const float*)op.w, op.bias, op.sub[0], op.sub[1], op.sub[2], \
                                                     op.div[0], op.div[1], op.div[2])
  const bool u8 = op.impl == LY_STEM_IN_U8;
  if (op.dtype == LY_F32) {
    LY_CHECK_ARG(aligned16<float>(op.dst), "stem: dst not 16-byte aligned");
    if (u8) LY_STEM(float, uint8_t); else LY_STEM(float, float);
  } else {
    LY_CHECK_ARG(aligned16<__nv_bfloat16>(op.dst), "stem: dst not 16-byte aligned");
    if (u8) LY_STEM(__nv_bfloat16, uint8_t); else LY_STEM(__nv_bfloat16, float);
  }
#undef LY_STEM
  return post_launch("stem");
}

template <typename T, int K, int S, int TH, int TW>
static int32_t run_dw_cfg(const ly_op& op, cudaStream_t s) {
  constexpr int V = 16 / sizeof(T);
  const int Ho = op.dst.H, Wo = op.dst.W;
  const int px = (Wo + TW - 1) / TW, py = (Ho + TH - 1) / TH;
  const long long total = (long long)op.B * py * px * (op.src.c / V);
  dw_kernel<T, K, S, TH, TW><<<grid_for(total, 128), 128, 0, s>>>(
      (const T*)op.src.ptr, op.src.H, op.src.W, op.src.ctot, op.src.c0, (T*)op.dst.ptr, op.dst.ctot, op.dst.c0,
      (const T*)op.res.ptr, op.res.ctot, op.res.c0, (const T*)op.w, op.bias, op.src.c, Ho, Wo, op.act, px, py, total);
  return post_launch("dwconv");
}

template <typename T>
static int32_t run_dw(const ly_op& op, cudaStream_t s) {
  LY_CHECK_ARG(aligned16<T>(op.src) && aligned16<T>(op.dst) && (!op.res.ptr || aligned16<T>(op.res)),
               "dwconv: views must be 16-byte aligned");
  const int Ho = (op.src.H + op.stride - 1) / op.stride, Wo = (op.src.W + op.stride - 1) / op.stride;
  LY_CHECK_ARG(op.dst.H == Ho && op.dst.W == Wo && op.dst.c == op.src.c, "dwconv: dst shape mismatch");
  if (op.k == 3 && op.stride == 1) return run_dw_cfg<T, 3, 1, 2, 4>(op, s);
  if (op.k == 3 && op.stride == 2) return run_dw_cfg<T, 3, 2, 1, 4>(op, s);
  if (op.k == 7 && op.stride == 1) return run_dw_cfg<T, 7, 1, 2, 2>(op, s);
  return run_dw_cfg<T, 7, 2, 1, 2>(op, s);
}

int32_t launch_dw(const ly_op& op, cudaStream_t s) {
  LY_CHECK_ARG(op.k == 3 || op.k == 7, "dwconv: k must be 3 or 7");
  LY_CHECK_ARG(op.stride == 1 || op.stride == 2, "dwconv: stride must be 1 or 2");
  LY_CHECK_ARG(op.src.ptr && op.dst.ptr && op.w && op.bias, "dwconv: null pointer");
  if (dw_tma_supported(op) && op.impl != LY_IMPL_SIMT) return launch_dw_tma(op, s);
  return op.dtype == LY_F32 ? run_dw<float>(op, s) : run_dw<__nv_bfloat16>(op, s);
}

template <typename T>
static int32_t run_pool(const ly_op& op, cudaStream_t s) {
  constexpr int V = 16 / sizeof(T);
  LY_CHECK_ARG(aligned16<T>(op.src), "pool: view must be 16-byte aligned");
  const size_t smem = (size_t)2 * op.src.H * op.src.W * V * sizeof(float);
  LY_CHECK_ARG(smem <= 200 * 1024, "pool: feature map %dx%d too large for the shared-memory plane", op.src.H, op.src.W);
  static bool attr_set = false;
  if (!attr_set) {
    LY_CUDA(cudaFuncSetAttribute(pool_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid(op.src.c / V, op.B);
  pool_kernel<T><<<grid, 256, smem, s>>>((T*)op.src.ptr, op.src.H, op.src.W, op.src.ctot, op.src.c0, op.src.c);
  return post_launch("sppf_pool");
}

int32_t launch_pool(const ly_op& op, cudaStream_t s) {
  LY_CHECK_ARG(op.src.ptr && op.src.ptr == op.dst.ptr, "pool: src and dst must be slices of the same buffer");
  LY_CHECK_ARG(op.dst.c0 == op.src.c0 + op.src.c && op.dst.c == 3 * op.src.c, "pool: dst must be the 3c channels after src");
  return op.dtype == LY_F32 ? run_pool<float>(op, s) : run_pool<__nv_bfloat16>(op, s);
}

template <typename T>
static int32_t run_up(const ly_op& op, cudaStream_t s) {
  constexpr int V = 16 / sizeof(T);
  LY_CHECK_ARG(aligned16<T>(op.src) && aligned16<T>(op.dst), "upsample: views must be 16-byte aligned");
  const long long total = (long long)op.B * op.dst.H * op.dst.W * (op.src.c / V);
  up_kernel<T><<<grid_for(total), 256, 0, s>>>((const T*)op.src.ptr, op.src.H, op.src.W, op.src.ctot, op.src.c0,
                                               (T*)op.dst.ptr, op.dst.ctot, op.dst.c0, op.src.c, total);
  return post_launch("upsample2x");
}

int32_t launch_up(const ly_op& op, cudaStream_t s) {
  LY_CHECK_ARG(op.src.ptr && op.dst.ptr, "upsample: null pointer");
  LY_CHECK_ARG(op.dst.H == 2 * op.src.H && op.dst.W == 2 * op.src.W && op.dst.c == op.src.c, "upsample: shape mismatch");
  return op.dtype == LY_F32 ? run_up<float>(op, s) : run_up<__nv_bfloat16>(op, s);
}

int32_t launch_export(const ly_op& op, cudaStream_t s) {
  LY_CHECK_ARG(op.src.ptr && op.nchw, "export: null pointer");
  const int HW = op.src.H * op.src.W;
  const long long total = (long long)op.B * op.nchw_c * HW;
  if (op.dtype == LY_F32)
    export_kernel<float><<<grid_for(total), 256, 0, s>>>((const float*)op.src.ptr, HW, op.src.ctot, op.src.c0, op.nchw,
                                                         op.nchw_ctot, op.nchw_c0, op.nchw_c, total);
  else
    export_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, s>>>((const __nv_bfloat16*)op.src.ptr, HW, op.src.ctot,
                                                                 op.src.c0, op.nchw, op.nchw_ctot, op.nchw_c0, op.nchw_c, total);
  return post_launch("export_nchw");
}

int32_t launch_import(const ly_op& op, cudaStream_t s) {
  LY_CHECK_ARG(op.dst.ptr && op.nchw, "import: null pointer");
  const int HW = op.dst.H * op.dst.W;
  const long long total = (long long)op.B * op.dst.c * HW;
  if (op.dtype == LY_F32)
    import_kernel<float><<<grid_for(total), 256, 0, s>>>(op.nchw, HW, op.nchw_ctot, op.nchw_c0, op.nchw_c,
                                                         (float*)op.dst.ptr, op.dst.ctot, op.dst.c0, op.dst.c, total);
  else
    import_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, s>>>(op.nchw, HW, op.nchw_ctot, op.nchw_c0, op.nchw_c,
                                                                 (__nv_bfloat16*)op.dst.ptr, op.dst.ctot, op.dst.c0, op.dst.c, total);
  return post_launch("import_nchw");
}

}  // namespace ly
