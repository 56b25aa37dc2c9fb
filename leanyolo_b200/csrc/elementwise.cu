// Bandwidth-bound kernels of the YOLOv10 forward: stem conv, depthwise convs (3x3 / 7x7,
// stride 1 / 2), SPPF max-pool pyramid, nearest x2 upsample into a concat slice and the
// NHWC<->NCHW boundary converters.  All NHWC with 16-byte channel vectors; templated on
// the storage type (bf16 hot path, fp32 check mode).
#include <type_traits>
#include <stdlib.h>
#include "common.cuh"
#include "preprocess.cuh"
#include <type_traits>

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
  pdl_trigger();
  pdl_wait();
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
// bf16 hot-path stem: the same op as an implicit GEMM on the warp-level tensor-core path.
// D[pixels, Cout] = A[pixels, K = 32] * W[Cout, 32]^T with K = 27 taps padded to 32; the
// CUDA-core version above needs 864 FMAs per output pixel and ran at 1/5 of the layer's
// HBM roofline, this one needs 12 shared-memory loads + 2*Cout/8 mma.sync per 16 pixels.
// (K = 32 with M x N = pixels x 32 is far below what a tcgen05/TMEM pipeline can amortise;
// the layer is bandwidth-bound: 12 B in (3 B as uint8) and 64 B out per 4 input pixels.)
//
// * The normalised image tile ((x - sub) / div, zero outside the image) is staged in shared
//   memory as bf16 [3][2*TH+1][IWP].  uint8 pixels with sub = 0, div = 255 could be kept
//   exact, but one code path serves both input types.
// * K slots are ordered so that the two taps (kx = 0, 1) of a (channel, ky) line are one
//   aligned 32-bit shared-memory load: slots 0..17 = 9 (c, ky) pairs, 18..26 = the kx = 2
//   singles, 27..31 = zero weights.  IWP = 144 makes the fragment loads bank-conflict free.
// * A warp owns one output row of the tile and walks it in 16-pixel m-tiles; the weight
//   fragments live in registers.  The mma columns are a PERMUTATION of the output channels
//   (column j of n-tile nt = channel 2*NT*(j/2) + 2*nt + j%2), so the accumulators a thread
//   holds for a pixel are 2*NT consecutive channels and go to HBM as 16-byte vectors straight
//   from registers (ncu on the first version, which staged the tile through shared memory:
//   25 warp-instructions per output pixel, issue slots 72 % busy at 41 % of the DRAM peak).
//   Weights and bias are pre-halved (SiLU(x) = h + h*tanh(h), h = x/2) and the normalisation is
//   a multiply by the reciprocal.
// ---------------------------------------------------------------------------------------
constexpr int SM_TW = 64, SM_TH = 8;                 // output tile per CTA
constexpr int SM_IH = 2 * SM_TH + 1;                 // input rows per channel
constexpr int SM_IWP = 144;                          // input row pitch (elements): (IWP / 2) mod 32 == 8
constexpr int SM_THREADS = SM_TH * 32;

__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Tile column t holds input column 2*wo0 - 4 + t: the 4-element vectors of the loader land on aligned 8-byte slots (one
// st.shared per vector; with column 0 = input column 2*wo0 - 1 every vector took three narrow stores and the kernel was
// bound by its shared-memory wavefronts, ncu: L1 81 % at 52 % DRAM).  Output pixel px, tap kx reads tile column
// 2*px + 3 + kx, so the (kx = 1, kx = 2) pair of a tap row sits at an EVEN column: one aligned 32-bit load.
// K slots: k < 18: (c, ky) = k >> 1 with kx = 1 + (k & 1); k = 18..26: the kx = 0 singles; 27..31: zero padding.
__device__ __forceinline__ int stem_slot_kx(int k) { return k < 18 ? 1 + (k & 1) : 0; }
// element offset of K slot k inside the staged tile, relative to the pixel's (2*py, 2*px) corner
__device__ __forceinline__ int stem_slot_off(int k) {
  if (k >= 27) return 4;
  const int j = k < 18 ? (k >> 1) : k - 18, kx = stem_slot_kx(k);
  return ((j / 3) * SM_IH + (j % 3)) * SM_IWP + kx + 3;
}
// index of K slot k inside the packed [27] weight row ([ky][kx][c]); -1 = zero padding
__device__ __forceinline__ int stem_slot_w(int k) {
  if (k >= 27) return -1;
  const int j = k < 18 ? (k >> 1) : k - 18, kx = stem_slot_kx(k);
  return ((j % 3) * 3 + kx) * 3 + (j / 3);
}

// one raw staging vector (16 bytes of fp32 or 4 bytes of uint8 pixels) global -> shared; `ok` false = zero fill, no read
template <int BYTES>
__device__ __forceinline__ void stem_cp_async(uint32_t dst, const void* src, bool ok) {
  const int sz = ok ? BYTES : 0;
  if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

template <typename TIn, int NT, int LDR = 0>
__global__ void __launch_bounds__(SM_THREADS, 4)
stem_mma_kernel(const TIn* __restrict__ x, int H, int W, __nv_bfloat16* __restrict__ dst, int dCtot, int dC0,
                const float* __restrict__ w, const float* __restrict__ bias,
                float s0, float s1, float s2, float d0, float d1, float d2, int fr, int fg, int fb, int B) {
  constexpr int CP = NT * 8;                           // padded output channels
  __shared__ __align__(16) __nv_bfloat16 tile[3 * SM_IH * SM_IWP];
  __shared__ __align__(16) __nv_bfloat16 wsm[CP * 32];
  pdl_trigger();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float r0 = 1.0f / d0, r1 = 1.0f / d1, r2 = 1.0f / d2;
  // Persistent CTAs (4 per SM) walk the 64 x 8 tiles: the weight staging and the register fragments below are built once
  // per CTA instead of once per tile (51 200 tiles at batch 256: ~10 % of the instructions of this issue-bound kernel).
  const int ntx = (Wo + SM_TW - 1) / SM_TW, nty = (Ho + SM_TH - 1) / SM_TH, total_tiles = ntx * nty * B;

  for (int i = tid; i < CP * 32; i += SM_THREADS) {
    const int col = i >> 5, wi = stem_slot_w(i & 31);
    const int nt = col >> 3, j = col & 7;
    const int co = 2 * NT * (j >> 1) + 2 * nt + (j & 1);      // channel behind mma column (nt, j)
    wsm[i] = __float2bfloat16_rn(wi >= 0 ? 0.5f * w[co * 27 + wi] : 0.f);
  }
  __syncthreads();

  const unsigned short* tl = reinterpret_cast<const unsigned short*>(tile);

  // LDR == 2: the raw pixels of tile n+1 travel global -> shared (cp.async) while tile n is multiplied; a thread converts
  // exactly the vectors it requested itself, so the raw buffer needs no barrier of its own.
  constexpr int P_VPL = SM_TW * 2 / 4 + 1, P_NL = 3 * SM_IH, P_NV = P_NL * P_VPL, P_PER = (P_NV + SM_THREADS - 1) / SM_THREADS;
  constexpr int P_DL = SM_THREADS / P_VPL, P_DV = SM_THREADS - P_DL * P_VPL;
  constexpr int P_VB = sizeof(TIn) == 1 ? 4 : 16;
  __shared__ __align__(16) uint8_t raw[LDR == 2 ? P_PER * SM_THREADS * P_VB : 16];
  const int p_line0 = tid / P_VPL, p_vi0 = tid - p_line0 * P_VPL;
  auto issue_raw = [&](int tq) {
    if constexpr (LDR == 2 && !std::is_same<TIn, ly_lb_desc>::value) {
      const int bq = tq / (ntx * nty), rq = tq - bq * (ntx * nty);
      const int hi0 = 2 * ((rq / ntx) * SM_TH) - 1, wi0 = 2 * ((rq % ntx) * SM_TW) - 4;
      const TIn* xb = x + (long long)bq * 3 * H * W;
      const bool border = !(hi0 >= 0 && hi0 + SM_IH <= H && wi0 >= 0 && wi0 + 4 * P_VPL <= W);
      const uint32_t rdst = (uint32_t)__cvta_generic_to_shared(raw) + (uint32_t)tid * P_VB;
      int line = p_line0, vi = p_vi0;
#pragma unroll
      for (int it = 0; it < P_PER; ++it) {
        const int c = (line >= SM_IH ? 1 : 0) + (line >= 2 * SM_IH ? 1 : 0);
        const int hi = hi0 + line - c * SM_IH, wi = wi0 + 4 * vi;
        bool ok = it < P_PER - 1 || line < P_NL;
        if (border) ok = ok && (unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W;
        stem_cp_async<P_VB>(rdst + (uint32_t)(it * SM_THREADS * P_VB), ok ? xb + (long long)((c * H + hi) * W + wi) : xb, ok);
        vi += P_DV; line += P_DL;
        if (vi >= P_VPL) { vi -= P_VPL; ++line; }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (LDR == 2 && (int)blockIdx.x < total_tiles) issue_raw(blockIdx.x);

  for (int tix = blockIdx.x; tix < total_tiles; tix += gridDim.x) {
  const int b = tix / (ntx * nty), trem = tix - b * (ntx * nty);
  const int ho0 = (trem / ntx) * SM_TH, wo0 = (trem % ntx) * SM_TW;
  if constexpr (std::is_same<TIn, ly_lb_desc>::value) {
    // Fused letterbox (utils/letterbox.py:9-91): `x` is the per-image descriptor array; every element of the staged tile
    // is sampled straight from the SOURCE image (cv2's fixed-point bilinear / 2x area / copy, border colour outside the
    // resized picture), so the letterboxed uint8 batch never exists in HBM.  The axis coefficient tables of the tile
    // (IEEE double arithmetic, OpenCV's order) are built once per CTA, not per pixel.
    constexpr int NCOL = 2 * SM_TW + 1;
    __shared__ AxisCoef cx[NCOL], cy[SM_IH];
    const ly_lb_desc d = x[b];
    const int mode = lb_mode(d);
    const int hi0 = 2 * ho0 - 1, wi0 = 2 * wo0 - 1;
    if (mode == 2) {
      for (int i = tid; i < NCOL + SM_IH; i += SM_THREADS) {
        if (i < NCOL) {
          const int rx = wi0 + i - d.left;
          if (rx >= 0 && rx < d.new_w) cx[i] = axis_coef(rx, d.new_w, d.src_w, true);
        } else {
          const int ry = hi0 + (i - NCOL) - d.top;
          if (ry >= 0 && ry < d.new_h) cy[i - NCOL] = axis_coef(ry, d.new_h, d.src_h, false);
        }
      }
      __syncthreads();
    }
    for (int i = tid; i < SM_IH * NCOL; i += SM_THREADS) {
      const int r = i / NCOL, col = i - r * NCOL;
      const int hi = hi0 + r, wi = wi0 + col;
      float f0 = 0.f, f1 = 0.f, f2 = 0.f;            // zero padding of the conv is applied AFTER the normalisation
      if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
        int v[3] = {fr, fg, fb};
        const int rx = wi - d.left, ry = hi - d.top;
        if (rx >= 0 && rx < d.new_w && ry >= 0 && ry < d.new_h) {
          if (mode == 0) {
            const uint8_t* p = d.src + ry * d.src_pitch + 3 * rx;
            v[0] = p[0]; v[1] = p[1]; v[2] = p[2];
          } else if (mode == 1) {
            const uint8_t* p0 = d.src + (2 * ry) * d.src_pitch + 6 * rx;
            const uint8_t* p1 = p0 + d.src_pitch;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = (p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2;
          } else {
            lb_bilinear(d, cx[col], cy[r], v);
          }
        }
        f0 = ((float)v[0] - s0) * r0; f1 = ((float)v[1] - s1) * r1; f2 = ((float)v[2] - s2) * r2;
      }
      tile[(0 * SM_IH + r) * SM_IWP + col + 3] = __float2bfloat16_rn(f0);     // tile column = input column - (2*wo0 - 4)
      tile[(1 * SM_IH + r) * SM_IWP + col + 3] = __float2bfloat16_rn(f1);
      tile[(2 * SM_IH + r) * SM_IWP + col + 3] = __float2bfloat16_rn(f2);
    }
  } else
  if constexpr (LDR == 2) {
    using Vec = typename std::conditional<sizeof(TIn) == 1, uchar4, float4>::type;
    const int hi0 = 2 * ho0 - 1, wi0 = 2 * wo0 - 4;
    const bool border = !(hi0 >= 0 && hi0 + SM_IH <= H && wi0 >= 0 && wi0 + 4 * P_VPL <= W);
    asm volatile("cp.async.wait_group 0;" ::: "memory");      // this thread's vectors of THIS tile have landed
    const Vec* rsrc = reinterpret_cast<const Vec*>(raw) + tid;
    int line = p_line0, vi = p_vi0;
#pragma unroll
    for (int it = 0; it < P_PER; ++it) {
      if (it < P_PER - 1 || line < P_NL) {
        const int c = (line >= SM_IH ? 1 : 0) + (line >= 2 * SM_IH ? 1 : 0);
        const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2), rc = c == 0 ? r0 : (c == 1 ? r1 : r2);
        bool ok = true;
        if (border) {
          const int hi = hi0 + line - c * SM_IH, wi = wi0 + 4 * vi;
          ok = (unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W;
        }
        const Vec q = rsrc[it * SM_THREADS];
        const float f0 = ok ? ((float)q.x - sc) * rc : 0.f, f1 = ok ? ((float)q.y - sc) * rc : 0.f;
        const float f2 = ok ? ((float)q.z - sc) * rc : 0.f, f3 = ok ? ((float)q.w - sc) * rc : 0.f;
        const __nv_bfloat162 lo2 = __floats2bfloat162_rn(f0, f1), hi2 = __floats2bfloat162_rn(f2, f3);
        *reinterpret_cast<uint2*>(tile + line * SM_IWP + 4 * vi) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&lo2), *reinterpret_cast<const uint32_t*>(&hi2));
      }
      vi += P_DV; line += P_DL;
      if (vi >= P_VPL) { vi -= P_VPL; ++line; }
    }
    // the raw slots of this thread are free again: request the next tile now, it lands while this one is multiplied
    if (tix + (int)gridDim.x < total_tiles) issue_raw(tix + gridDim.x);
  } else
  if constexpr (LDR == 1) {
    // Same staging as the loader below with the index arithmetic taken out of the per-vector path (ncu, source level: 469 M
    // warp-instructions per launch for 13 M MMAs, ~100 per staged vector: two divisions by constants, the bounds tests and
    // the 64-bit address, all done twice): item i + SM_THREADS is (line + DL, vector + DV) with one carry, the image-border
    // tests exist only in the tiles that touch the border (22 % of them at 640 x 640), validity is kept as a bit per vector.
    constexpr int VPL = SM_TW * 2 / 4 + 1, NL = 3 * SM_IH, NV = NL * VPL, PER = (NV + SM_THREADS - 1) / SM_THREADS;
    constexpr int DL = SM_THREADS / VPL, DV = SM_THREADS - DL * VPL;
    static_assert((PER - 1) * SM_THREADS <= NV, "only the last pass may run past the tile");
    using Vec = typename std::conditional<sizeof(TIn) == 1, uchar4, float4>::type;
    Vec v[PER];
    const int hi0 = 2 * ho0 - 1, wi0 = 2 * wo0 - 4;
    const TIn* xb = x + (long long)b * 3 * H * W;
    const bool interior = hi0 >= 0 && hi0 + SM_IH <= H && wi0 >= 0 && wi0 + 4 * VPL <= W;
    const int line0 = tid / VPL, vi0 = tid - line0 * VPL;
    unsigned okmask = 0;
    auto stage = [&](auto border_tag) {
      constexpr bool BORDER = decltype(border_tag)::value;
      {
        int line = line0, vi = vi0;
#pragma unroll
        for (int it = 0; it < PER; ++it) {
          const int c = (line >= SM_IH ? 1 : 0) + (line >= 2 * SM_IH ? 1 : 0);
          const int hi = hi0 + line - c * SM_IH, wi = wi0 + 4 * vi;
          bool ok = it < PER - 1 || line < NL;
          if (BORDER) ok = ok && (unsigned)hi < (unsigned)H && (unsigned)wi < (unsigned)W;
          if (ok) {
            v[it] = __ldg(reinterpret_cast<const Vec*>(xb + (long long)((c * H + hi) * W + wi)));
            okmask |= 1u << it;
          } else {
            v[it] = Vec{0, 0, 0, 0};
          }
          vi += DV; line += DL;
          if (vi >= VPL) { vi -= VPL; ++line; }
        }
      }
      {
        int line = line0, vi = vi0;
#pragma unroll
        for (int it = 0; it < PER; ++it) {
          if (it < PER - 1 || line < NL) {
            const int c = (line >= SM_IH ? 1 : 0) + (line >= 2 * SM_IH ? 1 : 0);
            const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2), rc = c == 0 ? r0 : (c == 1 ? r1 : r2);
            const bool ok = !BORDER || ((okmask >> it) & 1u) != 0;
            const float f0 = ok ? ((float)v[it].x - sc) * rc : 0.f, f1 = ok ? ((float)v[it].y - sc) * rc : 0.f;
            const float f2 = ok ? ((float)v[it].z - sc) * rc : 0.f, f3 = ok ? ((float)v[it].w - sc) * rc : 0.f;
            const __nv_bfloat162 lo2 = __floats2bfloat162_rn(f0, f1), hi2 = __floats2bfloat162_rn(f2, f3);
            *reinterpret_cast<uint2*>(tile + line * SM_IWP + 4 * vi) =
                make_uint2(*reinterpret_cast<const uint32_t*>(&lo2), *reinterpret_cast<const uint32_t*>(&hi2));
          }
          vi += DV; line += DL;
          if (vi >= VPL) { vi -= VPL; ++line; }
        }
      }
    };
    if (interior) stage(std::false_type{}); else stage(std::true_type{});
  } else
  {
    // aligned 4-element vectors: vector v of a line covers input columns 2*wo0 - 4 + 4v .. +3;
    // tile column 0 is input column 2*wo0 - 1 (the last element of vector 0)
    constexpr int VPL = SM_TW * 2 / 4 + 1, NV = 3 * SM_IH * VPL, PER = (NV + SM_THREADS - 1) / SM_THREADS;
    using Vec = typename std::conditional<sizeof(TIn) == 1, uchar4, float4>::type;
    Vec v[PER];
    const int hi0 = 2 * ho0 - 1, wi0 = 2 * wo0 - 4;
#pragma unroll
    for (int it = 0; it < PER; ++it) {
      const int i = tid + it * SM_THREADS;
      const int line = i / VPL, vi = i - line * VPL;
      const int c = line / SM_IH, hi = hi0 + (line - c * SM_IH), wi = wi0 + 4 * vi;
      const bool ok = i < NV && hi >= 0 && hi < H && wi >= 0 && wi < W;
      if (ok) v[it] = __ldg(reinterpret_cast<const Vec*>(x + (((long long)b * 3 + c) * H + hi) * W + wi));
      else v[it] = Vec{0, 0, 0, 0};
    }
#pragma unroll
    for (int it = 0; it < PER; ++it) {
      const int i = tid + it * SM_THREADS;
      if (i >= NV) continue;
      const int line = i / VPL, vi = i - line * VPL;
      const int c = line / SM_IH, hi = hi0 + (line - c * SM_IH), wi = wi0 + 4 * vi;
      const bool ok = hi >= 0 && hi < H && wi >= 0 && wi < W;
      const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2), rc = c == 0 ? r0 : (c == 1 ? r1 : r2);
      const float f0 = ok ? ((float)v[it].x - sc) * rc : 0.f, f1 = ok ? ((float)v[it].y - sc) * rc : 0.f;
      const float f2 = ok ? ((float)v[it].z - sc) * rc : 0.f, f3 = ok ? ((float)v[it].w - sc) * rc : 0.f;
      const __nv_bfloat162 lo2 = __floats2bfloat162_rn(f0, f1), hi2 = __floats2bfloat162_rn(f2, f3);
      *reinterpret_cast<uint2*>(tile + line * SM_IWP + 4 * vi) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&lo2), *reinterpret_cast<const uint32_t*>(&hi2));
    }
  }
  __syncthreads();
  // (fragments are re-read from the staged weights per tile: 16 shared-memory loads, instead of 30 registers live across the loader)
  // weight fragments: b0 = W[co = nt*8+g][k = ks*16 + 2t, +1], b1 = ... + 8
  uint32_t wb[NT][2][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t* wr = reinterpret_cast<const uint32_t*>(wsm + (nt * 8 + g) * 32 + ks * 16 + 2 * t);
      wb[nt][ks][0] = wr[0];
      wb[nt][ks][1] = wr[4];
    }
  float bv[NT][2];   // columns (nt, 2t), (nt, 2t+1) = channels 2*NT*t + 2*nt, +1
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { bv[nt][0] = 0.5f * bias[2 * NT * t + 2 * nt]; bv[nt][1] = 0.5f * bias[2 * NT * t + 2 * nt + 1]; }

  // per-thread fragment offsets (elements).  k-step 0: two (kx = 0,1) pairs; k-step 1: four singles
  const int op0 = stem_slot_off(2 * t), op1 = stem_slot_off(2 * t + 8);
  const int os0 = stem_slot_off(16 + 2 * t), os1 = stem_slot_off(17 + 2 * t);
  const int os2 = stem_slot_off(24 + 2 * t), os3 = stem_slot_off(25 + 2 * t);
  const int ho = ho0 + warp;

  for (int mt = 0; mt < SM_TW / 16; ++mt) {
    const int px0 = mt * 16;
    if (wo0 + px0 >= Wo || ho >= Ho) break;
    const int base0 = (2 * warp) * SM_IWP + 2 * (px0 + g), base1 = base0 + 16;
    uint32_t a[2][4];
    a[0][0] = *reinterpret_cast<const uint32_t*>(tl + base0 + op0);
    a[0][1] = *reinterpret_cast<const uint32_t*>(tl + base1 + op0);
    a[0][2] = *reinterpret_cast<const uint32_t*>(tl + base0 + op1);
    a[0][3] = *reinterpret_cast<const uint32_t*>(tl + base1 + op1);
    a[1][0] = (uint32_t)tl[base0 + os0] | ((uint32_t)tl[base0 + os1] << 16);
    a[1][1] = (uint32_t)tl[base1 + os0] | ((uint32_t)tl[base1 + os1] << 16);
    a[1][2] = (uint32_t)tl[base0 + os2] | ((uint32_t)tl[base0 + os3] << 16);
    a[1][3] = (uint32_t)tl[base1 + os2] | ((uint32_t)tl[base1 + os3] << 16);
    uint32_t lo[NT], hi[NT];   // pixel g / g + 8: bf16 pairs of channels 2*NT*t + 2*nt, +1
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float c[4] = {bv[nt][0], bv[nt][1], bv[nt][0], bv[nt][1]};
      mma_bf16_16816(c, a[0], wb[nt][0][0], wb[nt][0][1]);
      mma_bf16_16816(c, a[1], wb[nt][1][0], wb[nt][1][1]);
      // c = x/2 (weights and bias are pre-halved): SiLU(x) = h + h*tanh(h)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float th;
        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(c[j]));
        c[j] = fmaf(c[j], th, c[j]);
      }
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(c[0], c[1]), p1 = __floats2bfloat162_rn(c[2], c[3]);
      lo[nt] = *reinterpret_cast<const uint32_t*>(&p0);
      hi[nt] = *reinterpret_cast<const uint32_t*>(&p1);
    }
    __nv_bfloat16* orow = dst + (((long long)b * Ho + ho) * Wo + wo0 + px0) * dCtot + dC0 + 2 * NT * t;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int px = g + 8 * half;
      if (wo0 + px0 + px >= Wo) continue;
      const uint32_t* v = half ? hi : lo;
      __nv_bfloat16* o = orow + (long long)px * dCtot;
      if (NT % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NT / 4; ++q) *reinterpret_cast<uint4*>(o + 8 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      } else {
#pragma unroll
        for (int q = 0; q < NT / 2; ++q) *reinterpret_cast<uint2*>(o + 4 * q) = make_uint2(v[2 * q], v[2 * q + 1]);
      }
    }
  }
  __syncthreads();      // every warp has read the staged tile before the next one overwrites it
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
  pdl_trigger();
  pdl_wait();
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
// (layers.py:210-217).  One CTA = one image x CBV 16-byte channel vectors (128 contiguous
// bytes per pixel when CBV = 8, so global accesses are full lines): the H x W x CBV plane
// lives in shared memory in the storage type (max is exact in bf16) and each of the three
// chained pools is a separable row-max / column-max pass over it.  Reads channels [0,c) of
// the concat buffer and writes [c,2c), [2c,3c), [3c,4c) of the same buffer.
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ uint4 vmax(const uint4& a, const uint4& b);
template <>
__device__ __forceinline__ uint4 vmax<float>(const uint4& a, const uint4& b) {
  uint4 r;
  r.x = __float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(b.x)));
  r.y = __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(b.y)));
  r.z = __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(b.z)));
  r.w = __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(b.w)));
  return r;
}
template <>
__device__ __forceinline__ uint4 vmax<__nv_bfloat16>(const uint4& a, const uint4& b) {
  uint4 r;
  const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
  __nv_bfloat162* z = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) z[i] = __hmax2(x[i], y[i]);
  return r;
}

template <typename T>
__global__ void __launch_bounds__(256)
pool_kernel(T* buf, int H, int W, int Ctot, int C0, int C, int cbv) {
  constexpr int V = Elem<T>::kVec;
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint4 psm[];   // [2][H*W][cbv]
  const int HW = H * W, items = HW * cbv;
  uint4* cur = psm;
  uint4* tmp = psm + items;
  const int c = blockIdx.x * cbv * V;
  const int b = blockIdx.y;
  T* base = buf + (long long)b * HW * Ctot + C0 + c;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int p = i / cbv, v = i - p * cbv;
    cur[i] = *reinterpret_cast<const uint4*>(base + (long long)p * Ctot + v * V);
  }
  __syncthreads();
  for (int stage = 1; stage <= 3; ++stage) {
    for (int i = threadIdx.x; i < items; i += blockDim.x) {   // row max -> tmp
      const int p = i / cbv;
      const int x = p % W;
      uint4 m = cur[i];
      if (x >= 1) m = vmax<T>(m, cur[i - cbv]);
      if (x >= 2) m = vmax<T>(m, cur[i - 2 * cbv]);
      if (x + 1 < W) m = vmax<T>(m, cur[i + cbv]);
      if (x + 2 < W) m = vmax<T>(m, cur[i + 2 * cbv]);
      tmp[i] = m;
    }
    __syncthreads();
    const int rs = W * cbv;
    for (int i = threadIdx.x; i < items; i += blockDim.x) {   // column max -> cur (and out)
      const int p = i / cbv, v = i - p * cbv;
      const int y = p / W;
      uint4 m = tmp[i];
      if (y >= 1) m = vmax<T>(m, tmp[i - rs]);
      if (y >= 2) m = vmax<T>(m, tmp[i - 2 * rs]);
      if (y + 1 < H) m = vmax<T>(m, tmp[i + rs]);
      if (y + 2 < H) m = vmax<T>(m, tmp[i + 2 * rs]);
      *reinterpret_cast<uint4*>(base + (long long)p * Ctot + stage * C + v * V) = m;
      cur[i] = m;   // cur is only read by the row pass, which finished before the barrier above
    }
    __syncthreads();
  }
}

// Same op for maps whose row of channel vectors (W * cbv) fits the CTA: a thread keeps ONE (x, channel vector) column
// and walks the rows y = y0, y0 + rpp, ...: the (pixel, vector) split and the four x-border tests are computed once,
// every pass is shifts, adds and compares.  (ncu on the flat-index kernel above at 20x20x256, batch 256: 63 M
// warp-instructions of which 9 M were the maxima; i / cbv, p % W and p / W are runtime divisions on every item of all
// seven passes; the plane load stalled on one global load per loop iteration.)
template <typename T>
__global__ void __launch_bounds__(256)
pool_rows_kernel(T* buf, int H, int W, int Ctot, int C0, int C, int cbv_log2) {
  constexpr int V = Elem<T>::kVec;
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint4 psm[];   // [2][H*W][cbv]
  const int cbv = 1 << cbv_log2;
  const int cols = W << cbv_log2, items = H * cols;
  const int rpp = 256 / cols;                    // rows per pass (host: cols <= 256)
  const int y0 = threadIdx.x / cols, within = threadIdx.x - y0 * cols;
  const int Hn = y0 < rpp ? H : 0;               // idle tail threads (256 - rpp * cols) own no rows, but take part in the barriers
  uint4* cur = psm;
  uint4* tmp = psm + items;
  const int x = within >> cbv_log2, v = within & (cbv - 1);
  const int ol1 = x >= 1 ? -cbv : 0, ol2 = x >= 2 ? -2 * cbv : 0, or1 = x + 1 < W ? cbv : 0, or2 = x + 2 < W ? 2 * cbv : 0;
  T* gcol = buf + (long long)blockIdx.y * H * W * Ctot + C0 + blockIdx.x * cbv * V + (long long)x * Ctot + v * V;
  const long long grow = (long long)W * Ctot;    // elements between rows of the map
  const int step = rpp * cols;
  {
    int y = y0;
    for (; y + 3 * rpp < Hn; y += 4 * rpp) {      // four loads in flight per thread
      const T* g = gcol + (long long)y * grow;
      const uint4 a = *reinterpret_cast<const uint4*>(g);
      const uint4 b = *reinterpret_cast<const uint4*>(g + (long long)rpp * grow);
      const uint4 c = *reinterpret_cast<const uint4*>(g + 2LL * rpp * grow);
      const uint4 d = *reinterpret_cast<const uint4*>(g + 3LL * rpp * grow);
      const int i = y * cols + within;
      cur[i] = a; cur[i + step] = b; cur[i + 2 * step] = c; cur[i + 3 * step] = d;
    }
    for (; y < Hn; y += rpp) cur[y * cols + within] = *reinterpret_cast<const uint4*>(gcol + (long long)y * grow);
  }
  __syncthreads();
  for (int stage = 1; stage <= 3; ++stage) {
    for (int y = y0, i = y0 * cols + within; y < Hn; y += rpp, i += step) {   // row max -> tmp
      // a border tap re-reads the centre (max with itself): five independent loads, no branches
      const uint4 c0 = cur[i], l1 = cur[i + ol1], l2 = cur[i + ol2], r1 = cur[i + or1], r2 = cur[i + or2];
      tmp[i] = vmax<T>(vmax<T>(vmax<T>(l2, l1), vmax<T>(r1, r2)), c0);
    }
    __syncthreads();
    T* gout = gcol + stage * C;
    for (int y = y0, i = y0 * cols + within; y < Hn; y += rpp, i += step) {   // column max -> cur (and out)
      const uint4 c0 = tmp[i], u1 = tmp[y >= 1 ? i - cols : i], u2 = tmp[y >= 2 ? i - 2 * cols : i];
      const uint4 d1 = tmp[y + 1 < H ? i + cols : i], d2 = tmp[y + 2 < H ? i + 2 * cols : i];
      const uint4 m = vmax<T>(vmax<T>(vmax<T>(u2, u1), vmax<T>(d1, d2)), c0);
      *reinterpret_cast<uint4*>(gout + (long long)y * grow) = m;
      cur[i] = m;   // cur is only read by the row pass, which finished before the barrier above
    }
    __syncthreads();
  }
}

// nearest x2: dst(b, y, x, :) = src(b, y/2, x/2, :)   (layers.py:240)
template <typename T>
__global__ void __launch_bounds__(256)
up_kernel(const T* __restrict__ src, int sH, int sW, int sCtot, int sC0, T* __restrict__ dst, int dCtot, int dC0,
          int C, long long total) {
  pdl_trigger();
  pdl_wait();
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
  pdl_trigger();
  pdl_wait();
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
  pdl_trigger();
  pdl_wait();
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
  launch_k(stem_kernel<T, TIN>, grid, dim3(ST_THREADS), smem, s, (const TIN*)op.nchw, H, W, (T*)op.dst.ptr, op.dst.ctot, op.dst.c0, \
                                                     Cpad, (const float*)op.w, op.bias, op.sub[0], op.sub[1], op.sub[2], \
                                                     op.div[0], op.div[1], op.div[2])
  const bool u8 = op.impl == LY_STEM_IN_U8, lb = op.impl == LY_STEM_IN_LB;
  if (op.dtype == LY_BF16 && Cpad % 8 == 0 && Cpad <= 80 && W % 4 == 0 && op.dst.ctot % 8 == 0 && op.dst.c0 % 8 == 0 &&
      reinterpret_cast<uintptr_t>(op.nchw) % 16 == 0 && reinterpret_cast<uintptr_t>(op.dst.ptr) % 16 == 0) {
    const long long stem_tiles = (long long)((op.dst.W + SM_TW - 1) / SM_TW) * ((op.dst.H + SM_TH - 1) / SM_TH) * op.B;
    LY_CHECK_ARG(stem_tiles <= 0x7FFFFFFF, "stem: too many tiles");
    static const int stem_persist = getenv("LY_STEM_PERSIST") ? atoi(getenv("LY_STEM_PERSIST")) : 1;
    dim3 g2((unsigned)(stem_persist ? std::min<long long>(stem_tiles, 4LL * sm_count()) : stem_tiles), 1, 1);
    static const int stem_ldr = getenv("LY_STEM_LOADER") ? atoi(getenv("LY_STEM_LOADER")) : 1;
#define LY_STEM_MMA(TIN, NT) LY_STEM_MMA_L(TIN, NT, 0)
#define LY_STEM_MMA_L(TIN, NT, LDR)                                                                                   \
  launch_k(stem_mma_kernel<TIN, NT, LDR>, g2, dim3(SM_THREADS), 0, s, (const TIN*)op.nchw, H, W, (__nv_bfloat16*)op.dst.ptr, op.dst.ctot, \
                                                     op.dst.c0, (const float*)op.w, op.bias, op.sub[0], op.sub[1],    \
                                                     op.sub[2], op.div[0], op.div[1], op.div[2], op.nh, op.kdp, op.hd, op.B)
#define LY_STEM_NT(NT)                                                 \
  case NT:                                                             \
    if (lb) LY_STEM_MMA(ly_lb_desc, NT);                                                                        \
    else if (stem_ldr == 2) { if (u8) LY_STEM_MMA_L(uint8_t, NT, 2); else LY_STEM_MMA_L(float, NT, 2); }            \
    else if (stem_ldr) { if (u8) LY_STEM_MMA_L(uint8_t, NT, 1); else LY_STEM_MMA_L(float, NT, 1); }                 \
    else { if (u8) LY_STEM_MMA(uint8_t, NT); else LY_STEM_MMA(float, NT); }                                           \
    return post_launch("stem_mma");
    switch (Cpad / 8) {
      LY_STEM_NT(2) LY_STEM_NT(4) LY_STEM_NT(6) LY_STEM_NT(8) LY_STEM_NT(10)
      default: break;   // other widths: CUDA-core kernel below
    }
#undef LY_STEM_NT
#undef LY_STEM_MMA
#undef LY_STEM_MMA_L
  }
  LY_CHECK_ARG(!lb, "stem: the fused letterbox loader needs the bf16 tensor-core stem (Cout_pad in {16,32,48,64,80}, W %% 4 == 0)");
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
  launch_k(dw_kernel<T, K, S, TH, TW>, dim3(grid_for(total, 128)), dim3(128), 0, s,
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
  const int nvec = op.src.c / V;
  // channel vectors per CTA, dividing the channel count.  The kernel is a chain of 7 barrier-separated passes over a small
  // plane, i.e. latency-bound: smaller planes = more CTAs per SM to hide it (measured at 20x20x256, batch 256: see DESIGN)
  static const int plane_kb = getenv("LY_POOL_PLANE_KB") ? atoi(getenv("LY_POOL_PLANE_KB")) : 52;   // 100 KB: 0.140 ms, 52: 0.102, 26: 0.107, 13: 0.122
  int cbv = 8;
  while (cbv > 1 && (nvec % cbv != 0 || (size_t)2 * op.src.H * op.src.W * cbv * 16 > (size_t)plane_kb * 1024)) cbv >>= 1;
  const size_t smem = (size_t)2 * op.src.H * op.src.W * cbv * 16;
  LY_CHECK_ARG(smem <= 200 * 1024, "pool: feature map %dx%d too large for the shared-memory plane", op.src.H, op.src.W);
  static std::atomic<unsigned long long> attr_devs{0};   // per DEVICE: the attribute does not carry over to another GPU
  if (first_on_device(attr_devs)) {
    LY_CUDA(cudaFuncSetAttribute(pool_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  dim3 grid(nvec / cbv, op.B);
  static const int pool_rows = getenv("LY_POOL_ROWS") ? atoi(getenv("LY_POOL_ROWS")) : 1;
  if (pool_rows && op.src.W * cbv <= 256) {
    int lg = 0;
    while ((1 << lg) < cbv) ++lg;
    static std::atomic<unsigned long long> attr_devs_r{0};
    if (first_on_device(attr_devs_r)) {
      LY_CUDA(cudaFuncSetAttribute(pool_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    launch_k(pool_rows_kernel<T>, grid, dim3(256), smem, s, (T*)op.src.ptr, op.src.H, op.src.W, op.src.ctot, op.src.c0, op.src.c, lg);
    return post_launch("sppf_pool");
  }
  launch_k(pool_kernel<T>, grid, dim3(256), smem, s, (T*)op.src.ptr, op.src.H, op.src.W, op.src.ctot, op.src.c0, op.src.c, cbv);
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
  launch_k(up_kernel<T>, dim3(grid_for(total)), dim3(256), 0, s, (const T*)op.src.ptr, op.src.H, op.src.W, op.src.ctot, op.src.c0,
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
    launch_k(export_kernel<float>, dim3(grid_for(total)), dim3(256), 0, s, (const float*)op.src.ptr, HW, op.src.ctot, op.src.c0, op.nchw,
                                                         op.nchw_ctot, op.nchw_c0, op.nchw_c, total);
  else
    launch_k(export_kernel<__nv_bfloat16>, dim3(grid_for(total)), dim3(256), 0, s, (const __nv_bfloat16*)op.src.ptr, HW, op.src.ctot,
                                                                 op.src.c0, op.nchw, op.nchw_ctot, op.nchw_c0, op.nchw_c, total);
  return post_launch("export_nchw");
}

int32_t launch_import(const ly_op& op, cudaStream_t s) {
  LY_CHECK_ARG(op.dst.ptr && op.nchw, "import: null pointer");
  const int HW = op.dst.H * op.dst.W;
  const long long total = (long long)op.B * op.dst.c * HW;
  if (op.dtype == LY_F32)
    launch_k(import_kernel<float>, dim3(grid_for(total)), dim3(256), 0, s, op.nchw, HW, op.nchw_ctot, op.nchw_c0, op.nchw_c,
                                                         (float*)op.dst.ptr, op.dst.ctot, op.dst.c0, op.dst.c, total);
  else
    launch_k(import_kernel<__nv_bfloat16>, dim3(grid_for(total)), dim3(256), 0, s, op.nchw, HW, op.nchw_ctot, op.nchw_c0, op.nchw_c,
                                                                 (__nv_bfloat16*)op.dst.ptr, op.dst.ctot, op.dst.c0, op.dst.c, total);
  return post_launch("import_nchw");
}

}  // namespace ly
