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
constexpr int ST_TW = 32, ST_TH = 4;
constexpr int ST_IW = 2 * ST_TW + 1, ST_IH = 2 * ST_TH + 1;

template <typename T>
__global__ void __launch_bounds__(ST_TW* ST_TH)
stem_kernel(const float* __restrict__ x, int H, int W, T* __restrict__ dst, int dCtot, int dC0, int Cpad,
            const float* __restrict__ w, const float* __restrict__ bias,
            float s0, float s1, float s2, float d0, float d1, float d2) {
  extern __shared__ float smem[];
  float* tile = smem;                         // [3][ST_IH][ST_IW]
  float* ws = smem + 3 * ST_IH * ST_IW;       // [27][Cpad]  (transposed from [Cpad][27])
  float* bs = ws + 27 * Cpad;                 // [Cpad]
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
  for (int i = tid; i < 3 * ST_IH * ST_IW; i += blockDim.x) {
    int c = i / (ST_IH * ST_IW);
    int r = i - c * (ST_IH * ST_IW);
    int iy = r / ST_IW, ix = r - iy * ST_IW;
    int hi = hi0 + iy, wi = wi0 + ix;
    float v = 0.f;
    if (hi >= 0 && hi < H && wi >= 0 && wi < W)
      v = (x[(((long long)b * 3 + c) * H + hi) * W + wi] - sub[c]) / div[c];
    tile[i] = v;
  }
  __syncthreads();

  const int ty = tid / ST_TW, tx = tid - ty * ST_TW;
  const int ho = ho0 + ty, wo = wo0 + tx;
  if (ho >= Ho || wo >= Wo) return;
  float in[27];  // [ky][kx][ci] to match the packed weight order
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
      for (int c = 0; c < 3; ++c) in[(ky * 3 + kx) * 3 + c] = tile[(c * ST_IH + 2 * ty + ky) * ST_IW + 2 * tx + kx];

  T* out = dst + (((long long)b * Ho + ho) * Wo + wo) * dCtot + dC0;
  constexpr int V = Elem<T>::kVec;
  for (int co = 0; co < Cpad; co += V) {
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = bs[co + j];
#pragma unroll
    for (int t = 0; t < 27; ++t) {
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = fmaf(in[t], ws[t * Cpad + co + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = Elem<T>::act(acc[j]);
    store_vec<T>(out + co, acc);
  }
}

// ---------------------------------------------------------------------------------------
// Depthwise k x k conv (+bias, optional SiLU, optional residual added after the activation).
// One thread = one output pixel x one 16-byte channel vector; consecutive threads walk the
// channel vectors of a pixel, so every global access is a coalesced 16-byte vector.
// ---------------------------------------------------------------------------------------
template <typename T, int K>
__global__ void __launch_bounds__(256)
dw_kernel(const T* __restrict__ src, int sH, int sW, int sCtot, int sC0,
          T* __restrict__ dst, int dCtot, int dC0, const T* res, int rCtot, int rC0,
          const T* __restrict__ w, const float* __restrict__ bias,
          int C, int Ho, int Wo, int stride, int act, long long total) {
  constexpr int V = Elem<T>::kVec;
  const int cv = C / V;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * V;
    long long p = idx / cv;
    const int wo = (int)(p % Wo);
    p /= Wo;
    const int ho = (int)(p % Ho);
    const int b = (int)(p / Ho);
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = bias[c + j];
    const int hi0 = ho * stride - K / 2, wi0 = wo * stride - K / 2;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int hi = hi0 + ky;
      if (hi < 0 || hi >= sH) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int wi = wi0 + kx;
        if (wi < 0 || wi >= sW) continue;
        float xv[V], wv[V];
        load_vec<T>(src + (((long long)b * sH + hi) * sW + wi) * sCtot + sC0 + c, xv);
        load_vec<T>(w + (ky * K + kx) * C + c, wv);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = fmaf(xv[j], wv[j], acc[j]);
      }
    }
    const long long opix = ((long long)b * Ho + ho) * Wo + wo;
    if (act) {
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = Elem<T>::act(acc[j]);
    }
    if (res) {
      float rv[V];
      load_vec<T>(res + opix * rCtot + rC0 + c, rv);
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] += rv[j];
    }
    store_vec<T>(dst + opix * dCtot + dC0 + c, acc);
  }
}

// ---------------------------------------------------------------------------------------
// SPPF pyramid: y1 = pool5(x), y2 = pool5(y1), y3 = pool5(y2) with -inf padding
// (layers.py:210-217) == max over clipped 5x5 / 9x9 / 13x13 windows of x.  Reads channels
// [0,c) of the concat buffer and writes [c,2c), [2c,3c), [3c,4c) of the same buffer.
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
pool_kernel(T* buf, int H, int W, int Ctot, int C0, int C, long long total) {
  constexpr int V = Elem<T>::kVec;
  const int cv = C / V;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % cv) * V;
    long long p = idx / cv;
    const int x = (int)(p % W);
    p /= W;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    float m5[V], m9[V], m13[V];
#pragma unroll
    for (int j = 0; j < V; ++j) m5[j] = m9[j] = m13[j] = -INFINITY;
    for (int dy = -6; dy <= 6; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= H) continue;
      const int ady = dy < 0 ? -dy : dy;
      for (int dx = -6; dx <= 6; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= W) continue;
        const int adx = dx < 0 ? -dx : dx;
        const int r = ady > adx ? ady : adx;
        float v[V];
        load_vec<T>(buf + (((long long)b * H + yy) * W + xx) * Ctot + C0 + c, v);
#pragma unroll
        for (int j = 0; j < V; ++j) {
          m13[j] = fmaxf(m13[j], v[j]);
          if (r <= 4) m9[j] = fmaxf(m9[j], v[j]);
          if (r <= 2) m5[j] = fmaxf(m5[j], v[j]);
        }
      }
    }
    T* o = buf + (((long long)b * H + y) * W + x) * Ctot + C0 + c;
    store_vec<T>(o + C, m5);
    store_vec<T>(o + 2 * C, m9);
    store_vec<T>(o + 3 * C, m13);
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
  dim3 grid((op.dst.W + ST_TW - 1) / ST_TW, (op.dst.H + ST_TH - 1) / ST_TH, op.B);
  size_t smem = (3 * ST_IH * ST_IW + 28 * Cpad) * sizeof(float);
  if (op.dtype == LY_F32) {
    LY_CHECK_ARG(aligned16<float>(op.dst), "stem: dst not 16-byte aligned");
    stem_kernel<float><<<grid, ST_TW * ST_TH, smem, s>>>(op.nchw, H, W, (float*)op.dst.ptr, op.dst.ctot, op.dst.c0, Cpad,
                                                         (const float*)op.w, op.bias, op.sub[0], op.sub[1], op.sub[2],
                                                         op.div[0], op.div[1], op.div[2]);
  } else {
    LY_CHECK_ARG(aligned16<__nv_bfloat16>(op.dst), "stem: dst not 16-byte aligned");
    stem_kernel<__nv_bfloat16><<<grid, ST_TW * ST_TH, smem, s>>>(op.nchw, H, W, (__nv_bfloat16*)op.dst.ptr, op.dst.ctot,
                                                                 op.dst.c0, Cpad, (const float*)op.w, op.bias, op.sub[0],
                                                                 op.sub[1], op.sub[2], op.div[0], op.div[1], op.div[2]);
  }
  return post_launch("stem");
}

template <typename T>
static int32_t run_dw(const ly_op& op, cudaStream_t s) {
  constexpr int V = 16 / sizeof(T);
  LY_CHECK_ARG(aligned16<T>(op.src) && aligned16<T>(op.dst) && (!op.res.ptr || aligned16<T>(op.res)),
               "dwconv: views must be 16-byte aligned");
  const int Ho = (op.src.H + op.stride - 1) / op.stride, Wo = (op.src.W + op.stride - 1) / op.stride;
  LY_CHECK_ARG(op.dst.H == Ho && op.dst.W == Wo && op.dst.c == op.src.c, "dwconv: dst shape mismatch");
  const long long total = (long long)op.B * Ho * Wo * (op.src.c / V);
  const unsigned g = grid_for(total);
#define LY_DW(K)                                                                                              \
  dw_kernel<T, K><<<g, 256, 0, s>>>((const T*)op.src.ptr, op.src.H, op.src.W, op.src.ctot, op.src.c0,         \
                                    (T*)op.dst.ptr, op.dst.ctot, op.dst.c0, (const T*)op.res.ptr, op.res.ctot, \
                                    op.res.c0, (const T*)op.w, op.bias, op.src.c, Ho, Wo, op.stride, op.act, total)
  if (op.k == 3) LY_DW(3); else LY_DW(7);
#undef LY_DW
  return post_launch("dwconv");
}

int32_t launch_dw(const ly_op& op, cudaStream_t s) {
  LY_CHECK_ARG(op.k == 3 || op.k == 7, "dwconv: k must be 3 or 7");
  LY_CHECK_ARG(op.stride == 1 || op.stride == 2, "dwconv: stride must be 1 or 2");
  LY_CHECK_ARG(op.src.ptr && op.dst.ptr && op.w && op.bias, "dwconv: null pointer");
  return op.dtype == LY_F32 ? run_dw<float>(op, s) : run_dw<__nv_bfloat16>(op, s);
}

template <typename T>
static int32_t run_pool(const ly_op& op, cudaStream_t s) {
  constexpr int V = 16 / sizeof(T);
  LY_CHECK_ARG(aligned16<T>(op.src), "pool: view must be 16-byte aligned");
  const long long total = (long long)op.B * op.src.H * op.src.W * (op.src.c / V);
  pool_kernel<T><<<grid_for(total), 256, 0, s>>>((T*)op.src.ptr, op.src.H, op.src.W, op.src.ctot, op.src.c0, op.src.c, total);
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
