// Depthwise k x k convolution (bf16 hot path), TMA-staged.
//
// The op is HBM-bound (reads and writes every activation once, ~9-49 MACs per element), so
// the kernel is built around keeping many bytes in flight without spending registers on
// them: a producer warp streams (channel-block x halo-tile) boxes of the NHWC input into a
// shared-memory ring with 4-D TMA tile loads (out-of-image halo = hardware zero fill = the
// conv's zero padding), up to 8 compute warps consume them: a warp owns a 4x4 patch of output
// pixels and a lane 2 of the box's 64 channels, accumulation is fp32.  Persistent grid, static
// round-robin over tiles.  Channel counts that are not a multiple of 64 take the register-tiled
// CUDA-core kernel in elementwise.cu.
//
// Reference semantics: Conv(g=C) + BN (+SiLU) (+shortcut), leanyolo/models/yolov10/layers.py
// :51-88, 274-300, 455; RepVGGDW arrives here already merged into one 7x7 (modules.py).
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include "common.cuh"
#include "tma.cuh"
#include "tc.cuh"

namespace ly {

namespace {

constexpr int kMaxDwWarps = 8;
constexpr int kMaxStagesDw = 6;

struct DwParams {
  CUtensorMap tmIn;
  int npx, npy, tw, th, iwt, iht;          // tile = npx x npy patches of 4x4 outputs; raw box iwt x iht pixels
  int tiles_x, tiles_y, tiles_c, total_tiles;
  int Ho, Wo, C;
  int stages, stage_bytes, box_bytes;
  int act;
  int rev;               // walk the tiles last to first (see g_reverse)
  const __nv_bfloat16* w;
  const float* bias;
  __nv_bfloat16* dst; int dCtot, dC0;
  const __nv_bfloat16* res; int rCtot, rC0;
};

__device__ __forceinline__ float2 bf2_to_f2(uint32_t raw) {
  return make_float2(__uint_as_float(raw << 16), __uint_as_float(raw & 0xffff0000u));
}

// One compute warp = one 4x4 patch of output pixels; a lane owns 2 of the tile's 64 channels, so
// every shared-memory read is a conflict-free 128-byte pixel row, every global store / shortcut
// load a full 128-byte line, and an input pixel is read from shared memory ~once per patch
// (K=3: each input row once, all 9 taps in registers; K=7: one filter row of taps at a time).
template <int K, int S>
__global__ void __launch_bounds__(32 * (kMaxDwWarps + 1)) dw_tma_kernel(const __grid_constant__ DwParams p) {
  constexpr int IP = 3 * S + K;                     // input patch edge: 6 (3x3 s1), 9 (3x3 s2), 10 (7x7 s1)
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * kMaxStagesDw];
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStagesDw + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncw = p.npx * p.npy;                    // compute warps; warp ncw is the TMA producer
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmIn) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), ncw);
    }
    fence_barrier_init();
  }
  pdl_trigger();
  pdl_wait();
  __syncthreads();

  auto split = [&](int tile, int& xt, int& yt, int& ct, int& b) {
    int t = p.rev ? p.total_tiles - 1 - tile : tile;
    xt = t % p.tiles_x; t /= p.tiles_x;
    yt = t % p.tiles_y; t /= p.tiles_y;
    ct = t % p.tiles_c;
    b = t / p.tiles_c;
  };

  if (warp == ncw) {
    // ------------------------------------------------ producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int xt, yt, ct, b;
        split(tile, xt, yt, ct, b);
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), (uint32_t)p.box_bytes);
        tma_load_4d(smem_base + stage * p.stage_bytes, &p.tmIn, full_bar(stage), ct * 64, xt * p.tw * S - K / 2,
                    yt * p.th * S - K / 2, b);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // -------------------------------------------------- compute
  const int pyi = warp / p.npx, pxi = warp - pyi * p.npx;
  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    int xt, yt, ct, b;
    split(tile, xt, yt, ct, b);
    const int c = ct * 64 + 2 * lane;
    float acc[4][4][2];
    {
      const float2 b2 = __ldg(reinterpret_cast<const float2*>(p.bias + c));
#pragma unroll
      for (int y = 0; y < 4; ++y)
#pragma unroll
        for (int x = 0; x < 4; ++x) { acc[y][x][0] = b2.x; acc[y][x][1] = b2.y; }
    }
    const uint32_t* wg = reinterpret_cast<const uint32_t*>(p.w + c);   // tap t at wg[t * C / 2]
    const int wstride = p.C >> 1;
    float2 wv[K <= 3 ? K * K : K];
    if (K <= 3) {
#pragma unroll
      for (int t = 0; t < K * K; ++t) wv[t] = bf2_to_f2(__ldg(wg + t * wstride));
    }
    mbar_wait(full_bar(stage), phase);
    // halo pixel (X, Y) of the raw tile sits at ((Y * iwt + X) * 64 + channel) * 2 bytes
    const uint8_t* rt = smem_gen + (size_t)stage * p.stage_bytes + ((size_t)((4 * pyi * S) * p.iwt + 4 * pxi * S) * 64 + 2 * lane) * 2;
    if (K <= 3) {
#pragma unroll
      for (int iy = 0; iy < IP; ++iy) {
        float2 in[IP];
#pragma unroll
        for (int ix = 0; ix < IP; ++ix) in[ix] = bf2_to_f2(*reinterpret_cast<const uint32_t*>(rt + (size_t)(iy * p.iwt + ix) * 128));
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const int d = iy - ky;
          if (d >= 0 && d % S == 0 && d / S < 4) {
#pragma unroll
            for (int kx = 0; kx < K; ++kx)
#pragma unroll
              for (int ox = 0; ox < 4; ++ox) {
                acc[d / S][ox][0] = fmaf(in[ox * S + kx].x, wv[ky * K + kx].x, acc[d / S][ox][0]);
                acc[d / S][ox][1] = fmaf(in[ox * S + kx].y, wv[ky * K + kx].y, acc[d / S][ox][1]);
              }
          }
        }
      }
    } else {
      uint32_t wn[K];                       // next filter row, loaded one iteration ahead (L1/L2 latency off the FMA chain)
#pragma unroll
      for (int kx = 0; kx < K; ++kx) wn[kx] = __ldg(wg + kx * wstride);
#pragma unroll 1
      for (int ky = 0; ky < K; ++ky) {
#pragma unroll
        for (int kx = 0; kx < K; ++kx) wv[kx] = bf2_to_f2(wn[kx]);
        if (ky + 1 < K) {
#pragma unroll
          for (int kx = 0; kx < K; ++kx) wn[kx] = __ldg(wg + ((ky + 1) * K + kx) * wstride);
        }
#pragma unroll
        for (int oy = 0; oy < 4; ++oy) {
          float2 in[IP];
          const uint8_t* rr = rt + (size_t)((oy * S + ky) * p.iwt) * 128;
#pragma unroll
          for (int ix = 0; ix < IP; ++ix) in[ix] = bf2_to_f2(*reinterpret_cast<const uint32_t*>(rr + (size_t)ix * 128));
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int ox = 0; ox < 4; ++ox) {
              acc[oy][ox][0] = fmaf(in[ox * S + kx].x, wv[kx].x, acc[oy][ox][0]);
              acc[oy][ox][1] = fmaf(in[ox * S + kx].y, wv[kx].y, acc[oy][ox][1]);
            }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(stage));   // this warp is done reading the stage
    if (++stage == p.stages) { stage = 0; phase ^= 1u; }

    // stores: one 64-bit base per patch, 32-bit offsets per pixel (ncu: the per-pixel 64-bit index arithmetic
    // was ~40 % of the 1 600 instructions this kernel spent per 4x4 patch)
    const int oy0 = yt * p.th + 4 * pyi, ox0 = xt * p.tw + 4 * pxi;
    const size_t opix0 = ((size_t)b * p.Ho + oy0) * p.Wo + ox0;
    __nv_bfloat16* dbase = p.dst + opix0 * p.dCtot + p.dC0 + c;
    const __nv_bfloat16* rbase = p.res ? p.res + opix0 * p.rCtot + p.rC0 + c : nullptr;
    const bool act = p.act != 0;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      if (oy0 + y >= p.Ho) continue;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        if (ox0 + x >= p.Wo) continue;
        const int rel = y * p.Wo + x;
        float v0 = acc[y][x][0], v1 = acc[y][x][1];
        if (act) { v0 = silu_tanh(v0); v1 = silu_tanh(v1); }
        if (rbase) {
          const float2 r2 = bf2_to_f2(*reinterpret_cast<const uint32_t*>(rbase + rel * p.rCtot));
          v0 += r2.x; v1 += r2.y;
        }
        *reinterpret_cast<__nv_bfloat162*>(dbase + rel * p.dCtot) = __floats2bfloat162_rn(v0, v1);
      }
    }
  }
}

template <int K, int S>
int32_t launch_cfg(const ly_op& op, cudaStream_t st) {
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("dw_tma: cuTensorMapEncodeTiled entry point not available"); return LY_E_CUDA; }
  DwParams p;
  memset(&p, 0, sizeof(p));
  p.Ho = op.dst.H; p.Wo = op.dst.W; p.C = op.src.c;
  // tile = npx x npy patches (one compute warp each): raw box <= 48 KB, then most patches, fewest tiles
  {
    long long best_key = -1;
    for (int npx = 1; npx <= kMaxDwWarps; ++npx)
      for (int npy = 1; npx * npy <= kMaxDwWarps; ++npy) {
        const int tw = 4 * npx, th = 4 * npy;
        const long long box = (long long)((tw - 1) * S + K) * ((th - 1) * S + K) * 128;
        if (box > 48 * 1024 || (tw - 1) * S + K > 256 || (th - 1) * S + K > 256) continue;
        const long long tiles = (long long)((p.Wo + tw - 1) / tw) * ((p.Ho + th - 1) / th);
        // useful outputs per unit of work: total warp-patches = tiles * npx * npy (lower is better), then smaller box
        const long long key = tiles * npx * npy * 1000000LL + tiles * 1000 + box / 1024;
        if (best_key < 0 || key < best_key) { best_key = key; p.npx = npx; p.npy = npy; }
      }
    LY_CHECK_ARG(best_key >= 0, "dw_tma: no tile configuration fits");
  }
  p.tw = 4 * p.npx; p.th = 4 * p.npy;
  p.iwt = (p.tw - 1) * S + K; p.iht = (p.th - 1) * S + K;
  p.tiles_x = (p.Wo + p.tw - 1) / p.tw;
  p.tiles_y = (p.Ho + p.th - 1) / p.th;
  p.tiles_c = p.C / 64;
  const long long total = (long long)p.tiles_x * p.tiles_y * p.tiles_c * op.B;
  LY_CHECK_ARG(total <= 0x7FFFFFFF, "dw_tma: too many tiles");
  p.total_tiles = (int)total;
  p.box_bytes = p.iwt * p.iht * 128;
  p.stage_bytes = (p.box_bytes + 127) / 128 * 128;
  // K=7 is FMA/latency-bound (49 taps per output): three CTAs per SM; the 3x3 kernels are
  // bandwidth-bound: two CTAs with deeper rings
  const int ctas_per_sm = K == 7 ? 3 : 2;
  int stages = ((K == 7 ? 72 : 100) * 1024) / p.stage_bytes;
  if (stages > kMaxStagesDw) stages = kMaxStagesDw;
  if (stages < 2) stages = 2;
  p.stages = stages;
  p.act = op.act;
  p.rev = g_reverse;
  p.w = (const __nv_bfloat16*)op.w; p.bias = op.bias;
  p.dst = (__nv_bfloat16*)op.dst.ptr; p.dCtot = op.dst.ctot; p.dC0 = op.dst.c0;
  p.res = (const __nv_bfloat16*)op.res.ptr; p.rCtot = op.res.ctot; p.rC0 = op.res.c0;
  {
    char* base = (char*)op.src.ptr + (size_t)op.src.c0 * 2;
    cuuint64_t dims[4] = {(cuuint64_t)op.src.c, (cuuint64_t)op.src.W, (cuuint64_t)op.src.H, (cuuint64_t)op.B};
    cuuint64_t strides[3] = {(cuuint64_t)op.src.ctot * 2, (cuuint64_t)op.src.ctot * 2 * op.src.W,
                             (cuuint64_t)op.src.ctot * 2 * op.src.W * op.src.H};
    cuuint32_t box[4] = {64, (cuuint32_t)p.iwt, (cuuint32_t)p.iht, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&p.tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("dw_tma: cuTensorMapEncodeTiled failed with %d", (int)r); return LY_E_CUDA; }
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 128;
  static std::atomic<unsigned long long> attr_devs{0};   // per DEVICE: the attribute does not carry over to another GPU
  if (first_on_device(attr_devs)) {
    LY_CUDA(cudaFuncSetAttribute(dw_tma_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const int sms = sm_count();
  int grid = ctas_per_sm * sms;
  if (grid > p.total_tiles) grid = p.total_tiles;
  launch_k(dw_tma_kernel<K, S>, dim3(grid), dim3(32 * (p.npx * p.npy + 1)), smem, st, p);
  return post_launch("dwconv_tma");
}

// ---------------------------------------------------------------------------------------
// 3x3 stride 1: a compute thread owns one 16-byte channel vector of a horizontal strip of SL
// output pixels (fewer, wider stores per output than the patch kernel above, which wins on
// the small stride-1 maps that remain after the dw->1x1 fusion).
// ---------------------------------------------------------------------------------------
constexpr int kComputeThreads = 128;
constexpr int kThreadsDw = kComputeThreads + 32;
constexpr int kMaxStagesV = 6;

struct DwVParams {
  CUtensorMap tmIn;
  int tiles_x, tiles_y, tiles_c, total_tiles;
  int Ho, Wo, C;
  int stages, stage_bytes, box_bytes;
  int act;
  int rev;               // walk the tiles last to first (see g_reverse)
  const __nv_bfloat16* w;
  const float* bias;
  __nv_bfloat16* dst; int dCtot, dC0;
  const __nv_bfloat16* res; int rCtot, rC0;
};

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

template <int K, int S, int CBV, int TW_T, int TH_T>
__global__ void __launch_bounds__(kThreadsDw) dw_strip_kernel(const __grid_constant__ DwVParams p) {
  constexpr int CB = CBV * 8;                       // channels per block
  constexpr int IWt = (TW_T - 1) * S + K, IHt = (TH_T - 1) * S + K;
  constexpr int NW = kComputeThreads / CBV;         // pixel workers
  constexpr int SL = TW_T * TH_T / NW;              // outputs per worker (a horizontal strip)
  constexpr int SPR = TW_T / SL;                    // strips per tile row
  constexpr int IWS = (SL - 1) * S + K;             // input columns a strip touches
  static_assert(SL >= 1 && SL * NW == TW_T * TH_T && SPR * SL == TW_T, "bad dw tile configuration");
  (void)IHt;

  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * kMaxStagesV];
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStagesV + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmIn) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kComputeThreads / 32);
    }
    fence_barrier_init();
  }
  pdl_trigger();
  pdl_wait();
  __syncthreads();

  if (warp == kComputeThreads / 32) {
    // ------------------------------------------------ producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int t = p.rev ? p.total_tiles - 1 - tile : tile;
        const int xt = t % p.tiles_x; t /= p.tiles_x;
        const int yt = t % p.tiles_y; t /= p.tiles_y;
        const int ct = t % p.tiles_c;
        const int b = t / p.tiles_c;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), (uint32_t)p.box_bytes);
        tma_load_4d(smem_base + stage * p.stage_bytes, &p.tmIn, full_bar(stage), ct * CB, xt * TW_T * S - K / 2,
                    yt * TH_T * S - K / 2, b);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }

  // -------------------------------------------------- compute
  const int cv = threadIdx.x % CBV;
  const int wk = threadIdx.x / CBV;
  const int sy = wk / SPR;
  const int sx0 = (wk % SPR) * SL;
  int stage = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
    int t = p.rev ? p.total_tiles - 1 - tile : tile;
    const int xt = t % p.tiles_x; t /= p.tiles_x;
    const int yt = t % p.tiles_y; t /= p.tiles_y;
    const int ct = t % p.tiles_c;
    const int b = t / p.tiles_c;
    const int c = ct * CB + cv * 8;

    float acc[SL][8];
    {
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + c + 4));
#pragma unroll
      for (int x = 0; x < SL; ++x) {
        acc[x][0] = b0.x; acc[x][1] = b0.y; acc[x][2] = b0.z; acc[x][3] = b0.w;
        acc[x][4] = b1.x; acc[x][5] = b1.y; acc[x][6] = b1.z; acc[x][7] = b1.w;
      }
    }
    mbar_wait(full_bar(stage), phase);
    const uint8_t* tile_s = smem_gen + stage * p.stage_bytes;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      float wv[K][8];
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p.w + (ky * K + kx) * p.C + c));
        unpack8(raw, wv[kx]);
      }
      const uint8_t* row_s = tile_s + ((size_t)((sy * S + ky) * IWt + sx0 * S) * CB + cv * 8) * 2;
#pragma unroll
      for (int ix = 0; ix < IWS; ++ix) {
        float in[8];
        unpack8(*reinterpret_cast<const uint4*>(row_s + (size_t)ix * CB * 2), in);
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int d = ix - kx;
          if (d >= 0 && d % S == 0 && d / S < SL) {
#pragma unroll
            for (int j = 0; j < 8; j += 2)      // packed FFMA2: half the issue slots of the FMA stream
              ffma2(acc[d / S][j], acc[d / S][j + 1], in[j], in[j + 1], wv[kx][j], wv[kx][j + 1], acc[d / S][j], acc[d / S][j + 1]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(stage));   // this warp is done reading the stage
    if (++stage == p.stages) { stage = 0; phase ^= 1u; }

    const int oy = yt * TH_T + sy;
    if (oy < p.Ho) {
      const int ox0 = xt * TW_T + sx0;
      const size_t opix0 = ((size_t)b * p.Ho + oy) * p.Wo + ox0;       // one 64-bit base per strip, 32-bit offsets per pixel
      __nv_bfloat16* dbase = p.dst + opix0 * p.dCtot + p.dC0 + c;
      const __nv_bfloat16* rbase = p.res ? p.res + opix0 * p.rCtot + p.rC0 + c : nullptr;
      const bool act = p.act != 0;
#pragma unroll
      for (int x = 0; x < SL; ++x) {
        if (ox0 + x >= p.Wo) continue;
        if (act) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[x][j] = silu_tanh(acc[x][j]);
        }
        if (rbase) {
          float rv[8];
          load_vec<__nv_bfloat16>(rbase + x * p.rCtot, rv);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[x][j] += rv[j];
        }
        store_vec<__nv_bfloat16>(dbase + x * p.dCtot, acc[x]);
      }
    }
  }
}

template <int K, int S, int CBV, int TW_T, int TH_T>
int32_t launch_strip(const ly_op& op, cudaStream_t st) {
  constexpr int CB = CBV * 8;
  constexpr int IWt = (TW_T - 1) * S + K, IHt = (TH_T - 1) * S + K;
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("dw_tma: cuTensorMapEncodeTiled entry point not available"); return LY_E_CUDA; }
  DwVParams p;
  memset(&p, 0, sizeof(p));
  p.Ho = op.dst.H; p.Wo = op.dst.W; p.C = op.src.c;
  p.tiles_x = (p.Wo + TW_T - 1) / TW_T;
  p.tiles_y = (p.Ho + TH_T - 1) / TH_T;
  p.tiles_c = p.C / CB;
  const long long total = (long long)p.tiles_x * p.tiles_y * p.tiles_c * op.B;
  LY_CHECK_ARG(total <= 0x7FFFFFFF, "dw_tma: too many tiles");
  p.total_tiles = (int)total;
  p.box_bytes = IWt * IHt * CB * 2;
  p.stage_bytes = (p.box_bytes + 127) / 128 * 128;
  static const int dws_kb = getenv("LY_DWS_SMEM_KB") ? atoi(getenv("LY_DWS_SMEM_KB")) : 36;
  static const int dws_ctas = getenv("LY_DWS_CTAS") ? atoi(getenv("LY_DWS_CTAS")) : 6;
  // (round 2 sweep, KB per CTA / CTAs per SM: 72/3 0.404, 54/4 0.361, 40/5 0.446, 36/6 0.382, 100/2 0.464 ms on 640 ch @40^2; 36/6 is the
  //  best or within 1 % of the best on all four shapes tried)
  int stages = (dws_kb * 1024) / p.stage_bytes;
  if (stages > kMaxStagesV) stages = kMaxStagesV;
  if (stages < 2 && 2 * p.stage_bytes <= 200 * 1024) stages = 2;   // big 7x7 halo tiles: two CTAs per SM instead
  LY_CHECK_ARG(stages >= 2, "dw_tma: tile does not fit in shared memory");
  p.stages = stages;
  p.act = op.act;
  p.rev = g_reverse;
  p.w = (const __nv_bfloat16*)op.w; p.bias = op.bias;
  p.dst = (__nv_bfloat16*)op.dst.ptr; p.dCtot = op.dst.ctot; p.dC0 = op.dst.c0;
  p.res = (const __nv_bfloat16*)op.res.ptr; p.rCtot = op.res.ctot; p.rC0 = op.res.c0;
  {
    char* base = (char*)op.src.ptr + (size_t)op.src.c0 * 2;
    cuuint64_t dims[4] = {(cuuint64_t)op.src.c, (cuuint64_t)op.src.W, (cuuint64_t)op.src.H, (cuuint64_t)op.B};
    cuuint64_t strides[3] = {(cuuint64_t)op.src.ctot * 2, (cuuint64_t)op.src.ctot * 2 * op.src.W,
                             (cuuint64_t)op.src.ctot * 2 * op.src.W * op.src.H};
    cuuint32_t box[4] = {CB, IWt, IHt, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&p.tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("dw_tma: cuTensorMapEncodeTiled failed with %d", (int)r); return LY_E_CUDA; }
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 128;
  static std::atomic<unsigned long long> attr_devs{0};   // per DEVICE: the attribute does not carry over to another GPU
  if (first_on_device(attr_devs)) {
    LY_CUDA(cudaFuncSetAttribute(dw_strip_kernel<K, S, CBV, TW_T, TH_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  const int sms = sm_count();
  // several CTAs per SM: more loads in flight, the stores of one overlap the compute of the others
  int grid = dws_ctas * sms;
  if (grid > p.total_tiles) grid = p.total_tiles;
  launch_k(dw_strip_kernel<K, S, CBV, TW_T, TH_T>, dim3(grid), dim3(kThreadsDw), smem, st, p);
  return post_launch("dwconv_tma");
}


// ---------------------------------------------------------------------------------------
// 7x7 stride 1 (the merged RepVGGDW of the large-kernel CIB blocks): 49 taps per output make
// this the one FMA-bound depthwise conv.  (ncu on the generic kernel above: 40 % of the stall
// samples on global weight loads inside the tap loop, every input row re-read from shared
// memory once per (ky, oy) pair, issue slots 60 % busy at 31 % FMA-pipe utilisation.)  Here
//  * a CTA keeps ONE 64-channel block for its whole life, so a lane's 49 x 2 weights and its
//    bias stay in registers;
//  * the loop runs over INPUT rows: a row of the halo patch is loaded and unpacked once and
//    feeds every output row it touches;
//  * all arithmetic is packed FFMA2 on the lane's two channels.
// (Round 2: keeping the 49 taps in shared memory instead of registers, 72 instead of 154 registers and two CTAs = 20
//  compute warps per SM, measured the same 0.207 ms: the kernel is not occupancy-bound; ncu: IPC 1.9 of 4, 109 M
//  warp-instructions.  Removed again.)
// ---------------------------------------------------------------------------------------
constexpr int kDw7Warps = 10;

struct Dw7Params {
  CUtensorMap tmIn;
  int npx, npy, tw, th, iwt, iht;
  int tiles_x, tiles_y, tiles_c, tiles_per_cb, ctas_per_cb;
  int Ho, Wo, C;
  int stages, stage_bytes, box_bytes;
  int act;
  int rev;               // walk the tiles last to first (see g_reverse)
  const __nv_bfloat16* w;
  const float* bias;
  __nv_bfloat16* dst; int dCtot, dC0;
  const __nv_bfloat16* res; int rCtot, rC0;
};

__device__ __forceinline__ void ffma2_f2(float2& d, const float2& a, const float2& b) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(d.x), "+f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}

template <int PH>
__global__ void __launch_bounds__(32 * (kDw7Warps + 1)) dw7_kernel(const __grid_constant__ Dw7Params p) {
  constexpr int K = 7, IPX = 4 + K - 1, IPY = PH + K - 1;       // halo patch of a 4 x PH output patch
  constexpr int kWBytes = 0;                                   // (weights live in registers)
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) unsigned long long bars[2 * kMaxStagesDw];
  const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stage0 = smem_base + kWBytes;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kMaxStagesDw + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ncw = p.npx * p.npy;                    // compute warps; warp ncw is the TMA producer
  const int cb = blockIdx.x % p.tiles_c, cta = blockIdx.x / p.tiles_c;
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmIn) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), ncw);
    }
    fence_barrier_init();
  }
  pdl_trigger();
  // weights are parameters: fetched before waiting on the previous kernel.  The lane's 49 taps stay
  // in registers as packed bf16 pairs and are unpacked at the point of use (2 ALU ops next to 4..8
  // FFMA2): fp32 pairs in shared memory cost one LDS.64 per tap and output row, and made the
  // kernel shared-memory-bandwidth bound (measured 0.24 ms vs 0.30 for the generic kernel).
  uint32_t wreg[K * K];
  {
    const uint32_t* wg = reinterpret_cast<const uint32_t*>(p.w + cb * 64 + 2 * lane);
#pragma unroll
    for (int t = 0; t < K * K; ++t) wreg[t] = __ldg(wg + (size_t)t * (p.C >> 1));
  }
  pdl_wait();
  __syncthreads();

  auto split = [&](int t, int& xt, int& yt, int& b) {
    if (p.rev) t = p.tiles_per_cb - 1 - t;
    xt = t % p.tiles_x; t /= p.tiles_x;
    yt = t % p.tiles_y;
    b = t / p.tiles_y;
  };

  if (warp == ncw) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cta; t < p.tiles_per_cb; t += p.ctas_per_cb) {
        int xt, yt, b;
        split(t, xt, yt, b);
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), (uint32_t)p.box_bytes);
        tma_load_4d(stage0 + stage * p.stage_bytes, &p.tmIn, full_bar(stage), cb * 64, xt * p.tw - K / 2, yt * p.th - K / 2, b);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
    return;
  }
  if (warp > ncw) return;

  const int pyi = warp / p.npx, pxi = warp - pyi * p.npx;
  const int c = cb * 64 + 2 * lane;
  const float2 bias2 = __ldg(reinterpret_cast<const float2*>(p.bias + c));
  int stage = 0;
  uint32_t phase = 0;
  for (int t = cta; t < p.tiles_per_cb; t += p.ctas_per_cb) {
    int xt, yt, b;
    split(t, xt, yt, b);
    float2 acc[PH][4];
#pragma unroll
    for (int y = 0; y < PH; ++y)
#pragma unroll
      for (int x = 0; x < 4; ++x) acc[y][x] = bias2;
    mbar_wait(full_bar(stage), phase);
    // halo pixel (X, Y) of the raw tile sits at ((Y * iwt + X) * 64 + channel) * 2 bytes
    const uint8_t* rt = smem_gen + kWBytes + (size_t)stage * p.stage_bytes + ((size_t)((PH * pyi) * p.iwt + 4 * pxi) * 64 + 2 * lane) * 2;
#pragma unroll
    for (int iy = 0; iy < IPY; ++iy) {
      float2 in[IPX];
#pragma unroll
      for (int ix = 0; ix < IPX; ++ix) in[ix] = bf2_to_f2(*reinterpret_cast<const uint32_t*>(rt + (size_t)(iy * p.iwt + ix) * 128));
#pragma unroll
      for (int oy = 0; oy < PH; ++oy) {
        const int ky = iy - oy;
        if (ky >= 0 && ky < K) {
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const float2 w2 = bf2_to_f2(wreg[ky * K + kx]);
#pragma unroll
            for (int ox = 0; ox < 4; ++ox) ffma2_f2(acc[oy][ox], in[ox + kx], w2);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(stage));   // this warp is done reading the stage
    if (++stage == p.stages) { stage = 0; phase ^= 1u; }

    // stores: one 64-bit base per patch, 32-bit offsets per pixel (the per-pixel 64-bit index
    // arithmetic of the generic kernel was 30 % of this kernel's instructions)
    const int oy0 = yt * p.th + PH * pyi, ox0 = xt * p.tw + 4 * pxi;
    const size_t opix0 = ((size_t)b * p.Ho + oy0) * p.Wo + ox0;
    __nv_bfloat16* dbase = p.dst + opix0 * p.dCtot + p.dC0 + c;
    const __nv_bfloat16* rbase = p.res ? p.res + opix0 * p.rCtot + p.rC0 + c : nullptr;
    const bool act = p.act != 0;
#pragma unroll
    for (int y = 0; y < PH; ++y) {
      if (oy0 + y >= p.Ho) continue;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        if (ox0 + x >= p.Wo) continue;
        const int rel = y * p.Wo + x;
        float v0 = acc[y][x].x, v1 = acc[y][x].y;
        if (act) { v0 = silu_tanh(v0); v1 = silu_tanh(v1); }
        if (rbase) {
          const float2 r2 = bf2_to_f2(*reinterpret_cast<const uint32_t*>(rbase + rel * p.rCtot));
          v0 += r2.x; v1 += r2.y;
        }
        *reinterpret_cast<__nv_bfloat162*>(dbase + rel * p.dCtot) = __floats2bfloat162_rn(v0, v1);
      }
    }
  }
}

template <int PH>
int32_t launch_dw7(const ly_op& op, cudaStream_t st) {
  constexpr int K = 7;
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("dw_tma: cuTensorMapEncodeTiled entry point not available"); return LY_E_CUDA; }
  Dw7Params p;
  memset(&p, 0, sizeof(p));
  p.Ho = op.dst.H; p.Wo = op.dst.W; p.C = op.src.c;
  {
    long long best_key = -1;
    for (int npx = 1; npx <= kDw7Warps; ++npx)
      for (int npy = 1; npx * npy <= kDw7Warps; ++npy) {
        const int tw = 4 * npx, th = PH * npy;
        const long long box = (long long)(tw + K - 1) * (th + K - 1) * 128;
        if (box > 40 * 1024 || tw + K - 1 > 256 || th + K - 1 > 256) continue;
        const long long tiles = (long long)((p.Wo + tw - 1) / tw) * ((p.Ho + th - 1) / th);
        const long long key = tiles * npx * npy * 1000000LL + tiles * 1000 + box / 1024;   // fewest warp-patches, then fewest tiles
        if (best_key < 0 || key < best_key) { best_key = key; p.npx = npx; p.npy = npy; }
      }
    LY_CHECK_ARG(best_key >= 0, "dw7: no tile configuration fits");
  }
  p.tw = 4 * p.npx; p.th = PH * p.npy;
  p.iwt = p.tw + K - 1; p.iht = p.th + K - 1;
  p.tiles_x = (p.Wo + p.tw - 1) / p.tw;
  p.tiles_y = (p.Ho + p.th - 1) / p.th;
  p.tiles_c = p.C / 64;
  const long long per_cb = (long long)p.tiles_x * p.tiles_y * op.B;
  LY_CHECK_ARG(per_cb <= 0x7FFFFFFF, "dw7: too many tiles");
  p.tiles_per_cb = (int)per_cb;
  p.box_bytes = p.iwt * p.iht * 128;
  p.stage_bytes = (p.box_bytes + 127) / 128 * 128;
  // registers (the 49 packed taps + the accumulators) allow 65536 / (threads * ~170) CTAs per SM; the
  // ring takes what shared memory is left: the loads of a 33 KB halo box take about as long as its
  // 49-tap compute, so two stages starve (measured: 22 % of the samples in the barrier wait)
  const int threads = 32 * (p.npx * p.npy + 1);
  const int ctas_per_sm = std::max(1, std::min(3, 65536 / (threads * 176)));
  static const int stages_env = getenv("LY_DW7_STAGES") ? atoi(getenv("LY_DW7_STAGES")) : 0;
  p.stages = stages_env ? stages_env : (int)std::min<long long>(kMaxStagesDw, (200LL * 1024 / ctas_per_sm - 256) / p.stage_bytes);
  LY_CHECK_ARG(p.stages >= 2, "dw7: tile does not fit in shared memory");
  const size_t smem = (size_t)p.stages * p.stage_bytes + 128;
  p.act = op.act;
  p.rev = g_reverse;
  p.w = (const __nv_bfloat16*)op.w; p.bias = op.bias;
  p.dst = (__nv_bfloat16*)op.dst.ptr; p.dCtot = op.dst.ctot; p.dC0 = op.dst.c0;
  p.res = (const __nv_bfloat16*)op.res.ptr; p.rCtot = op.res.ctot; p.rC0 = op.res.c0;
  {
    char* base = (char*)op.src.ptr + (size_t)op.src.c0 * 2;
    cuuint64_t dims[4] = {(cuuint64_t)op.src.c, (cuuint64_t)op.src.W, (cuuint64_t)op.src.H, (cuuint64_t)op.B};
    cuuint64_t strides[3] = {(cuuint64_t)op.src.ctot * 2, (cuuint64_t)op.src.ctot * 2 * op.src.W,
                             (cuuint64_t)op.src.ctot * 2 * op.src.W * op.src.H};
    cuuint32_t box[4] = {64, (cuuint32_t)p.iwt, (cuuint32_t)p.iht, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&p.tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("dw7: cuTensorMapEncodeTiled failed with %d", (int)r); return LY_E_CUDA; }
  }
  static std::atomic<unsigned long long> attr_devs{0};   // per DEVICE: the attribute does not carry over to another GPU
  if (first_on_device(attr_devs)) {
    LY_CUDA(cudaFuncSetAttribute(dw7_kernel<PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  // every CTA owns one channel block: the grid is a multiple of the channel-block count
  const int sms = sm_count();
  long long per = (long long)ctas_per_sm * sms / p.tiles_c;
  if (per < 1) per = 1;
  if (per > per_cb) per = per_cb;
  p.ctas_per_cb = (int)per;
  launch_k(dw7_kernel<PH>, dim3((unsigned)(per * p.tiles_c)), dim3(32 * (p.npx * p.npy + 1)), smem, st, p);
  return post_launch("dwconv7");
}

// waste of covering an Ho x Wo map with TW x TH tiles (1.0 = none)
double cover(int Ho, int Wo, int tw, int th) {
  return (double)((Wo + tw - 1) / tw * tw) * ((Ho + th - 1) / th * th) / ((double)Ho * Wo);
}

}  // namespace

bool dw_tma_supported(const ly_op& op) {
  if (op.dtype != LY_BF16) return false;
  if (!((op.k == 3 && (op.stride == 1 || op.stride == 2)) || (op.k == 7 && op.stride == 1))) return false;
  if (op.src.c % 64 || op.src.c0 % 8 || op.src.ctot % 8 || op.dst.c0 % 8 || op.dst.ctot % 8) return false;
  if (op.res.ptr && (op.res.c0 % 8 || op.res.ctot % 8)) return false;
  if (reinterpret_cast<uintptr_t>(op.src.ptr) % 16) return false;
  return true;
}

int32_t launch_dw_tma(const ly_op& op, cudaStream_t s) {
  if (op.k == 3 && op.stride == 1) {
    const bool wide = cover(op.dst.H, op.dst.W, 20, 4) < cover(op.dst.H, op.dst.W, 16, 8);
    return wide ? launch_strip<3, 1, 8, 20, 4>(op, s) : launch_strip<3, 1, 8, 16, 8>(op, s);
  }
  if (op.k == 3 && op.stride == 2) return launch_cfg<3, 2>(op, s);
  static const int dw7_mode = getenv("LY_DW7") ? atoi(getenv("LY_DW7")) : 2;   // 0: generic patch kernel, 2 / 4: dw7_kernel with 4x2 / 4x4 patches
  if (dw7_mode == 2) return launch_dw7<2>(op, s);
  if (dw7_mode == 4) return launch_dw7<4>(op, s);
  return launch_cfg<7, 1>(op, s);
}

}  // namespace ly
