// CUDA-core tiled implicit-GEMM convolution.
//
// Role: (1) the fp32 "check mode" twin of the tensor-core conv (1e-4 parity against
// the fp32 oracle rules out bf16/TF32 MMA, SURVEY §7 hard part 5b); (2) a bring-up /
// bisecting implementation for bf16 storage (LY_IMPL_SIMT).  It is NOT the measured
// hot path: bf16 LY_IMPL_AUTO goes to conv_tc.cu (tcgen05/TMEM/TMA).
//
// Computes dst[b,ho,wo,co] = act(sum_{r,s,ci} src[b,ho*S+r-P,wo*S+s-P,ci] * w[co,r,s,ci] + bias[co]) (+ res)
// Reference semantics: leanyolo/models/yolov10/layers.py:51-88 with BN folded.
#include "common.cuh"

namespace ly {

namespace {

constexpr int TM = 64;   // output pixels per CTA
constexpr int TN = 64;   // output channels per CTA
constexpr int TK = 16;   // input channels per smem stage
constexpr int NT = 256;  // threads

template <typename T>
__global__ void __launch_bounds__(NT)
conv_simt_kernel(const T* __restrict__ src, int sH, int sW, int sCtot, int sC0, int Cin,
                 T* dst, int dCtot, int dC0,
                 const T* res, int rCtot, int rC0,
                 const T* __restrict__ w, const float* __restrict__ bias,
                 float* nchw, int nCtot, int nC0, int nC,
                 int B, int Ho, int Wo, int Cout, int k, int stride, int act) {
  pdl_trigger();
  pdl_wait();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const long long M = (long long)B * Ho * Wo;
  const int pad = k / 2;

  // loader mapping: 64 rows x 16 channels, 4 channels per thread
  const int lrow = tid >> 2;
  const int lch = (tid & 3) * 4;
  long long lm = (long long)m0 + lrow;
  const bool lvalid = lm < M;
  int lb = 0, lho = 0, lwo = 0;
  if (lvalid) {
    lb = (int)(lm / (Ho * Wo));
    int r = (int)(lm - (long long)lb * Ho * Wo);
    lho = r / Wo;
    lwo = r - lho * Wo;
  }
  const int ln = n0 + lrow;  // output channel this thread loads weights for
  const bool lnvalid = ln < Cout;

  // compute mapping: 16x16 threads, 4 pixels x 4 channels each
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int Ktaps = k * k;
  for (int tap = 0; tap < Ktaps; ++tap) {
    const int r = tap / k, s = tap - r * k;
    const int hi = lho * stride + r - pad, wi = lwo * stride + s - pad;
    const bool pvalid = lvalid && hi >= 0 && hi < sH && wi >= 0 && wi < sW;
    const T* sp = src + (((long long)lb * sH + hi) * sW + wi) * sCtot + sC0;
    const T* wp = w + ((long long)ln * Ktaps + tap) * Cin;
    for (int c0 = 0; c0 < Cin; c0 += TK) {
      float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
      if (pvalid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) av[j] = Elem<T>::to_f(sp[c0 + lch + j]);
      }
      if (lnvalid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bv[j] = Elem<T>::to_f(wp[c0 + lch + j]);
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        As[lch + j][lrow] = av[j];
        Bs[lch + j][lrow] = bv[j];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < TK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = (long long)m0 + ty * 4 + i;
    if (m >= M) continue;
    const int b = (int)(m / (Ho * Wo));
    const int rem = (int)(m - (long long)b * Ho * Wo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= Cout) continue;
      float v = acc[i][j] + bias[n];
      if (act) v = Elem<T>::act(v);
      if (res) v += Elem<T>::to_f(res[m * rCtot + rC0 + n]);
      if (dst) dst[m * dCtot + dC0 + n] = Elem<T>::from_f(v);
      if (nchw && n < nC) nchw[((long long)b * nCtot + nC0 + n) * ((long long)Ho * Wo) + rem] = v;
    }
  }
}

template <typename T>
int32_t run(const ly_op& op, cudaStream_t st) {
  const int Ho = op.src.H / op.stride, Wo = op.src.W / op.stride;
  const int Cout = op.dst.ptr ? op.dst.c : (op.nchw_c + 15) / 16 * 16;
  const long long M = (long long)op.B * Ho * Wo;
  dim3 grid((unsigned)((M + TM - 1) / TM), (unsigned)((Cout + TN - 1) / TN));
  launch_k(conv_simt_kernel<T>, grid, dim3(NT), 0, st,
      (const T*)op.src.ptr, op.src.H, op.src.W, op.src.ctot, op.src.c0, op.src.c,
      (T*)op.dst.ptr, op.dst.ctot, op.dst.c0,
      (const T*)op.res.ptr, op.res.ctot, op.res.c0,
      (const T*)op.w, op.bias, op.nchw, op.nchw_ctot, op.nchw_c0, op.nchw_c,
      op.B, Ho, Wo, Cout, op.k, op.stride, op.act);
  return post_launch("conv_simt");
}

}  // namespace

int32_t launch_conv_simt(const ly_op& op, cudaStream_t s) {
  LY_CHECK_ARG(!op.up.ptr, "conv: the upsampled pre-activation addend is only implemented on the bf16 tensor-core path");
  LY_CHECK_ARG(op.k == 1 || op.k == 3, "conv: k must be 1 or 3 (got %d)", op.k);
  LY_CHECK_ARG(op.stride == 1 || op.stride == 2, "conv: stride must be 1 or 2");
  LY_CHECK_ARG(op.src.ptr && op.w && op.bias, "conv: null src/w/bias");
  LY_CHECK_ARG(op.dst.ptr || op.nchw, "conv: no destination");
  LY_CHECK_ARG(op.src.c % 16 == 0, "conv: src channel slice must be a multiple of 16");
  LY_CHECK_ARG(op.src.H % op.stride == 0 && op.src.W % op.stride == 0, "conv: H,W must divide by the stride");
  return op.dtype == LY_F32 ? run<float>(op, s) : run<__nv_bfloat16>(op, s);
}

}  // namespace ly
