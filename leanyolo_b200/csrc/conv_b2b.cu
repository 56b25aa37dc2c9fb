// Back-to-back GEMM: Conv 3x3 (+BN, SiLU) -> Conv2d 1x1 (+bias) of a box-regression stack (head.py:86-92) in ONE
// launch, the 3x3 result never leaves the SM.
//
//   D1[128 px, C1] = sum over 9 taps A[px + tap, C] * W1[tap][C1, C]^T      (band mode of conv_tc.cu: one TMA box of R+2
//                                                                             padded image rows serves all nine taps)
//   A2 = bf16(SiLU(D1 + b1))                                                 (middle warps: TMEM -> registers -> the
//                                                                             128-byte-swizzled K-major A tile of GEMM 2)
//   D2[128 px, C2] = A2 * W2[C2, C1]^T + b2 -> public NCHW fp32 tensor       (final warps)
//
// Why it pays where the generic chain kernel (chain_tc.cu) did not: the 3x3 stage is bound by its MMA issue (36 MMAs x
// ~50 cycles per 128-pixel tile, conv_tc.cu / DESIGN.md "finding 2"), the 1x1 adds 4 MMAs to that stream, and both
// epilogues are tile-parallel (two tile groups of four quarter warps each, four accumulator stages per GEMM), so neither
// waits for the other.  Layer by layer the pair costs a 64-channel bf16 round trip through HBM and a second launch.
//
// Stride-2 variant (backbone cv1 -> c2.cv1: 3x3/s2 32->64 then 1x1 64->64, the 64-channel 160x160 intermediate of yolov10s, 1.7 GB
// of HBM traffic per step at batch 256, never exists): the band is loaded as FOUR parity planes (even / odd input rows x even /
// odd input columns, TMA elementStrides = 2), each [plane row][P pixels] with the same pitch P as the accumulator rows
// (m = oy*P + ox).  Tap (ky, kx) of output pixel m then reads row m + r_off*P + j_off of ONE plane: ky = 1 the even rows, ky =
// 0 / 2 the odd rows at r_off 0 / 1; kx = 1 the even columns at j_off 1, kx = 0 / 2 the odd columns at j_off 0 / 1 (plane column j
// holds input column 2j - 2 resp. 2j - 1, so column 0 is the left zero padding).  Same constant-offset descriptors as stride 1.
//
// Roles (one CTA per SM, persistent over (band, image) units):
//   warp 0       TMA: both weight sets once (resident), the input band of every unit (2 stages)
//   warp 1       tcgen05.mma issuer: GEMM 1 of tile i, then GEMM 2 of tile i-1 (its A2 tile is ready by then)
//   warps 2-9    middle: 2 tile groups x 4 TMEM lane quarters: D1 -> +bias, SiLU -> bf16 -> A2[group]
//   warps 10-17  final:  2 tile groups x 4 quarters: D2 -> +bias (SiLU optional) -> NCHW fp32
// TMEM: 4 stages x C1 columns (D1) + 4 stages x C2 columns (D2) <= 512.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "common.cuh"
#include "tma.cuh"
#include "tc.cuh"

namespace ly {

namespace {

constexpr int kThreadsB = 64 + 32 * 16;
constexpr int kAcc = 4;                       // accumulator stages per GEMM
constexpr uint32_t kSmemMaxB = 227 * 1024;
#ifdef LY_TC_EXP
#define BEXP(p, bit) (((p).exp & (bit)) != 0)
#else
#define BEXP(p, bit) false
#endif

struct BParams {
  CUtensorMap tmA, tmW1, tmW2;
  int C, C1, C2;                 // input / middle / output channels (multiples of 16; C, C1 in {32, 64})
  int H, W, B;                   // OUTPUT extent of the 3x3 (= input extent at stride 1, half of it at stride 2)
  int s2;                        // the 3x3 has stride 2: the band is four parity planes (see the header comment)
  int pw, npieces, piece1;       // stride 2: columns per TMA box, boxes per plane row, first column of the second box
  uint32_t tap_off16[9];         // A-operand offset of every tap from the tile's first row, 16-byte units
  uint32_t plane16[4];           // stride 2: offsets of the planes EE, EO, OE, OO inside a stage, 16-byte units
  __nv_bfloat16* dst; int dCtot, dC0, st256;   // NHWC bf16 destination (instead of nchw)
  int band_r, band_w, band_mt, bands, total_units;
  uint32_t mg_bw, mg_bands;
  int a_stage, a_box;            // bytes
  int w1_slab, w2_bytes;         // bytes of one tap slab of W1 (C1 rows x C channels), of W2
  int a2_stage;                  // bytes of one A2 tile (128 rows x C1 channels)
  uint32_t idesc1, idesc2, hi_a, hi_a2;
  int act2;
  int exp;                       // -DLY_TC_EXP builds: 1 skip GEMM 1, 2 skip the final stores, 4 skip GEMM 2, 8 skip the middle math
  const float* b1; const float* b2;
  float* nchw; int nCtot, nC0, nC;
};

// barrier block layout (8-byte slots)
enum { kAFull = 0, kAEmpty = 2, kWFull = 4, kD1Full = 5, kD1Empty = 9, kA2Full = 13, kA2Empty = 15, kD2Full = 17, kD2Empty = 21,
       kTmemSlot = 25, kBarSlots = 26 };

// The 9 x KS MMAs of one tile of GEMM 1 as straight-line code: every operand of the loop lives in a register (the first version
// rebuilt the descriptors from the parameter bank inside the loop: 65 cycles per iteration on the single issuing thread, more
// than the MMA itself).
template <int KS>
__device__ __forceinline__ void b2b_gemm1(uint32_t d, uint32_t am, uint32_t w1lo, uint32_t slab16, const uint32_t (&toff)[9],
                                          uint64_t hi, uint32_t idesc, bool skip) {
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uint32_t at = am + toff[tap];
    const uint32_t bt = w1lo + (uint32_t)tap * slab16;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk)
      if (!skip) umma_bf16(d, hi | (uint64_t)(at + 2 * kk), hi | (uint64_t)(bt + 2 * kk), idesc, (tap | kk) != 0 ? 1u : 0u);
  }
}

__global__ void __launch_bounds__(kThreadsB, 1) conv_b2b_kernel(const __grid_constant__ BParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // carve: [A band x 2][W1 resident: 9 slabs][W2][A2 x 2][biases][barriers]
  const uint32_t a_base = base;
  const uint32_t w1_base = a_base + 2u * (uint32_t)p.a_stage;
  const uint32_t w2_base = w1_base + 9u * (uint32_t)p.w1_slab;
  const uint32_t a2_base = w2_base + (uint32_t)p.w2_bytes;
  const uint32_t f_off = (a2_base - base) + 2u * (uint32_t)p.a2_stage;
  float* b1_s = reinterpret_cast<float*>(gen + f_off);
  float* b2_s = b1_s + p.C1;
  const uint32_t bar = base + f_off + (uint32_t)(p.C1 + p.C2) * 4u;
  auto B = [&](int slot) { return bar + 8u * (uint32_t)slot; };
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (B(kTmemSlot) - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  for (int i = threadIdx.x; i < p.C1; i += kThreadsB) b1_s[i] = 0.5f * p.b1[i];               // SiLU(x) = h + h*tanh(h), h = x/2
  for (int i = threadIdx.x; i < p.C2; i += kThreadsB) b2_s[i] = (p.act2 ? 0.5f : 1.0f) * p.b2[i];
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmW1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmW2) : "memory");
    for (int s = 0; s < 2; ++s) {
      mbar_init(B(kAFull + s), 1);
      mbar_init(B(kAEmpty + s), 1);
      mbar_init(B(kA2Full + s), 4);
      mbar_init(B(kA2Empty + s), 1);
    }
    for (int s = 0; s < kAcc; ++s) {
      mbar_init(B(kD1Full + s), 1);
      mbar_init(B(kD1Empty + s), 4);
      mbar_init(B(kD2Full + s), 1);
      mbar_init(B(kD2Empty + s), 4);
    }
    mbar_init(B(kWFull), 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(B(kTmemSlot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t d2_col0 = (uint32_t)(kAcc * p.C1);
  const int mt_n = p.band_mt;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      mbar_expect_tx(B(kWFull), 9u * (uint32_t)(p.C1 * p.C * 2) + (uint32_t)(p.C2 * p.C1 * 2));
      for (int tap = 0; tap < 9; ++tap) tma_load_2d(w1_base + (uint32_t)tap * p.w1_slab, &p.tmW1, B(kWFull), tap * p.C, 0);
      tma_load_2d(w2_base, &p.tmW2, B(kWFull), 0, 0);
      int sa = 0;
      uint32_t pa = 0;
      for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
        const uint32_t b = p.mg_bands ? __umulhi((uint32_t)unit, p.mg_bands) : (uint32_t)unit;
        const int band = unit - (int)b * p.bands;
        mbar_wait(B(kAEmpty + sa), pa ^ 1u);
        if (BEXP(p, 16) && unit >= (int)blockIdx.x + 2 * (int)gridDim.x) {      // knock-out: no band loads after the first two
          mbar_arrive(B(kAFull + sa));
          if (++sa == 2) { sa = 0; pa ^= 1u; }
          continue;
        }
        mbar_expect_tx(B(kAFull + sa), (uint32_t)p.a_box);
        const uint32_t sb = a_base + (uint32_t)sa * p.a_stage;
        if (!p.s2) {
          tma_load_4d(sb, &p.tmA, B(kAFull + sa), 0, -1, band * p.band_r - 1, (int)b);
        } else {
          // four parity planes (even / odd input rows x even / odd input columns), every plane row loaded by 1-row boxes of
          // pw columns (a box may traverse at most 256 input columns); rows and pieces start at multiples of 8 pixels
          const uint32_t rowb = (uint32_t)(p.band_w * p.C * 2), colb = (uint32_t)(p.C * 2);
          const int oy0 = band * p.band_r;
          for (int pl = 0; pl < 4; ++pl) {
            const int odd_r = pl >> 1, odd_c = pl & 1;
            const int rows = p.band_r + odd_r;
            for (int r = 0; r < rows; ++r)
              for (int pc = 0; pc < p.npieces; ++pc) {
                const int j0 = pc ? p.piece1 : 0;                      // plane column of the box's first pixel
                tma_load_4d(sb + (p.plane16[pl] << 4) + (uint32_t)r * rowb + (uint32_t)j0 * colb, &p.tmA, B(kAFull + sa), 0,
                            2 * j0 - 2 + odd_c,          // plane column j holds input column 2j - 2 (even planes) / 2j - 1 (odd planes)
                            2 * (oy0 + r) - odd_r, (int)b);   // plane row r holds input row 2(oy0 + r) (even) / 2(oy0 + r) - 1 (odd)
              }
          }
        }
        if (++sa == 2) { sa = 0; pa ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ================================
    if (elect_one()) {
      const uint32_t row16 = (uint32_t)p.C >> 3;                         // one pixel row of the band, 16-byte units
      uint32_t toff[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) toff[t] = p.tap_off16[t];
      const int ks1 = p.C / 16, ks2 = p.C1 / 16;
      const uint64_t hi1 = (uint64_t)p.hi_a << 32, hi2 = (uint64_t)p.hi_a2 << 32;
      const uint32_t id1 = p.idesc1, id2 = p.idesc2, slab16 = (uint32_t)p.w1_slab >> 4, w1lo = (w1_base >> 4) | (1u << 16);
      const uint32_t a_lo0 = (a_base >> 4) | (1u << 16), a_stage16 = (uint32_t)p.a_stage >> 4;
      const uint32_t a2lo0 = (a2_base >> 4) | (1u << 16), a2_stage16 = (uint32_t)p.a2_stage >> 4, w2lo = (w2_base >> 4) | (1u << 16);
      const uint32_t c1 = (uint32_t)p.C1, c2 = (uint32_t)p.C2;
      mbar_wait(B(kWFull), 0);
      tc_fence_after();
      long long seq = 0;
#ifdef LY_TC_PROFILE
      long long w_afull = 0, w_d1e = 0, w_a2f = 0, w_d2e = 0, t_i1 = 0; const long long istart = clock64();
#define IT(x) const long long x = clock64()
#define IADD(acc, a, b) acc += (b) - (a)
#else
#define IT(x)
#define IADD(acc, a, b)
#endif
      auto gemm2 = [&](long long j) {
        const int g = (int)(j & 1), st = (int)(j & (kAcc - 1));
        IT(g0);
        mbar_wait(B(kA2Full + g), (uint32_t)((j >> 1) & 1));
        IT(g1);
        mbar_wait(B(kD2Empty + st), (uint32_t)(((j >> 2) & 1) ^ 1));
        IT(g2);
        IADD(w_a2f, g0, g1); IADD(w_d2e, g1, g2);
        tc_fence_after();
        const uint32_t d = tmem_base + d2_col0 + (uint32_t)st * c2;
        const uint32_t alo = a2lo0 + (uint32_t)g * a2_stage16;
#pragma unroll 4
        for (int kk = 0; kk < ks2; ++kk)
          if (!BEXP(p, 4)) umma_bf16(d, hi2 | (uint64_t)(alo + 2 * kk), hi2 | (uint64_t)(w2lo + 2 * kk), id2, kk != 0 ? 1u : 0u);
        umma_commit(B(kA2Empty + g));
        umma_commit(B(kD2Full + st));
      };
      int sa = 0;
      uint32_t pa = 0;
      for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) {
        IT(u0);
        mbar_wait(B(kAFull + sa), pa);
        tc_fence_after();
        IT(u1);
        IADD(w_afull, u0, u1);
        for (int mt = 0; mt < mt_n; ++mt) {
          const int st = (int)(seq & (kAcc - 1));
          IT(m0);
          mbar_wait(B(kD1Empty + st), (uint32_t)(((seq >> 2) & 1) ^ 1));
          tc_fence_after();
          IT(m1);
          IADD(w_d1e, m0, m1);
          const uint32_t d = tmem_base + (uint32_t)st * c1;
          const uint32_t am = a_lo0 + (uint32_t)sa * a_stage16 + (uint32_t)(mt * 128) * row16;
          if (ks1 == 4) b2b_gemm1<4>(d, am, w1lo, slab16, toff, hi1, id1, BEXP(p, 1));
          else b2b_gemm1<2>(d, am, w1lo, slab16, toff, hi1, id1, BEXP(p, 1));
          umma_commit(B(kD1Full + st));
          IT(m2);
          IADD(t_i1, m1, m2);
          if (seq >= 1) gemm2(seq - 1);
          ++seq;
        }
        umma_commit(B(kAEmpty + sa));      // the band is free once everything issued so far has completed
        if (++sa == 2) { sa = 0; pa ^= 1u; }
      }
      if (seq >= 1) gemm2(seq - 1);
#ifdef LY_TC_PROFILE
      if (blockIdx.x == 0)
        printf("[b2b prof] issuer: total %lld wait_afull %lld wait_d1empty %lld issue_gemm1 %lld wait_a2full %lld wait_d2empty %lld tiles %lld\n",
               clock64() - istart, w_afull, w_d1e, t_i1, w_a2f, w_d2e, seq);
#endif
    }
  } else if (warp < 10) {
    // ============================== middle: D1 -> A2 ==========================
    const int q = warp & 3, g = (warp - 2) >> 2;
    const uint32_t row = (uint32_t)(q * 32 + lane);
    const int nch = p.C1 >> 4;
    const uint32_t rowb = (uint32_t)p.C1 * 2u;                           // bytes per A2 row: 128 (SW128) or 64 (SW64)
    const uint32_t swz = p.C1 == 64 ? (row & 7u) : ((row >> 1) & 3u);
    const uint32_t a2row = a2_base + (uint32_t)g * p.a2_stage + row * rowb;
    long long total_items = 0;
    for (int unit = blockIdx.x; unit < p.total_units; unit += gridDim.x) total_items += mt_n;
#ifdef LY_TC_PROFILE
    long long t_wait1 = 0, t_ld = 0, t_wait2 = 0, t_body = 0, t_fence = 0; const long long tstart = clock64();
#define BT(x) const long long x = clock64()
#else
#define BT(x)
#endif
    for (long long j = g; j < total_items; j += 2) {
      const int st = (int)(j & (kAcc - 1));
      BT(c0);
      mbar_wait(B(kD1Full + st), (uint32_t)((j >> 2) & 1));
      tc_fence_after();
      BT(c1);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(st * p.C1);
      uint32_t nxt[16];
      tmem_ld16(taddr, nxt);
      mbar_wait(B(kA2Empty + g), (uint32_t)(((j >> 1) & 1) ^ 1));       // GEMM 2 of this group's previous tile has read A2[g]
      BT(c2);
      for (int c = 0; c < nch; ++c) {
        uint32_t pk[8];
        tmem_ld_wait();
        const float4* bp = reinterpret_cast<const float4*>(b1_s + c * 16);
        float v[16];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float4 bb = bp[jj];
          ffma2(v[4 * jj + 0], v[4 * jj + 1], __uint_as_float(nxt[4 * jj]), __uint_as_float(nxt[4 * jj + 1]), 0.5f, 0.5f, bb.x, bb.y);
          ffma2(v[4 * jj + 2], v[4 * jj + 3], __uint_as_float(nxt[4 * jj + 2]), __uint_as_float(nxt[4 * jj + 3]), 0.5f, 0.5f, bb.z, bb.w);
        }
        if (c + 1 < nch) {
          tmem_ld16(taddr + (c + 1) * 16, nxt);
        } else {                                   // the last chunk sits in registers: release the accumulator stage
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(B(kD1Empty + st));
        }
#pragma unroll
        for (int jj = 0; jj < 16; jj += 2) {
          if (!BEXP(p, 8)) silu2_from_half(v[jj], v[jj + 1]);
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[jj], v[jj + 1]);
          pk[jj >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a2row + ((((uint32_t)(2 * c)) ^ swz) << 4)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a2row + ((((uint32_t)(2 * c + 1)) ^ swz) << 4)), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]),
                     "r"(pk[7]) : "memory");
      }
      BT(c3);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(B(kA2Full + g));
#ifdef LY_TC_PROFILE
      const long long c4 = clock64();
      t_wait1 += c1 - c0; t_wait2 += c2 - c1; t_body += c3 - c2; t_fence += c4 - c3;
#endif
    }
#ifdef LY_TC_PROFILE
    if (blockIdx.x == 0 && lane == 0 && q == 0)
      printf("[b2b prof] middle group %d: total %lld wait_d1full %lld ld+wait_a2empty %lld body %lld fence+arrive %lld items %lld\n", g,
             clock64() - tstart, t_wait1, t_wait2, t_body, t_fence, (total_items - g + 1) / 2);
    (void)t_ld;
#endif
  } else {
    // ============================== final: D2 -> NCHW fp32 ====================
    const int q = warp & 3, g = (warp - 10) >> 2;
    const uint32_t row = (uint32_t)(q * 32 + lane);
    const int nch = p.C2 >> 4;
    const bool act2 = p.act2 != 0;
    const float pre = act2 ? 0.5f : 1.0f;
    const uint32_t hw = (uint32_t)p.H * (uint32_t)p.W;
    int it_unit = blockIdx.x, it_mt = 0;
    auto advance = [&]() { if (++it_mt == mt_n) { it_mt = 0; it_unit += gridDim.x; } };
    if (g) advance();
    for (long long j = g; it_unit < p.total_units; j += 2) {
      const uint32_t b = p.mg_bands ? __umulhi((uint32_t)it_unit, p.mg_bands) : (uint32_t)it_unit;
      const uint32_t bd = (uint32_t)it_unit - b * (uint32_t)p.bands;
      const uint32_t m = (uint32_t)it_mt * 128u + row;
      const uint32_t oy = __umulhi(m, p.mg_bw), ox = m - oy * (uint32_t)p.band_w;
      const uint32_t h = bd * (uint32_t)p.band_r + oy;
      const bool valid = ox < (uint32_t)p.W && oy < (uint32_t)p.band_r && h < (uint32_t)p.H;
      float* nrow = (valid && p.nchw) ? p.nchw + (size_t)(b * (uint32_t)p.nCtot + (uint32_t)p.nC0) * hw + (h * (uint32_t)p.W + ox) : nullptr;
      __nv_bfloat16* drow = (valid && p.dst) ? p.dst + (size_t)((b * (uint32_t)p.H + h) * (uint32_t)p.W + ox) * (uint32_t)p.dCtot + p.dC0 : nullptr;
      const int st = (int)(j & (kAcc - 1));
      mbar_wait(B(kD2Full + st), (uint32_t)((j >> 2) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + d2_col0 + (uint32_t)(st * p.C2);
      uint32_t nxt[16];
      tmem_ld16(taddr, nxt);
      for (int ch = 0; ch < nch; ++ch) {
        const int c = ch * 16;
        float v[16];
        tmem_ld_wait();
        const float4* bp = reinterpret_cast<const float4*>(b2_s + c);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float4 bb = bp[jj];
          ffma2(v[4 * jj + 0], v[4 * jj + 1], __uint_as_float(nxt[4 * jj + 0]), __uint_as_float(nxt[4 * jj + 1]), pre, pre, bb.x, bb.y);
          ffma2(v[4 * jj + 2], v[4 * jj + 3], __uint_as_float(nxt[4 * jj + 2]), __uint_as_float(nxt[4 * jj + 3]), pre, pre, bb.z, bb.w);
        }
        if (ch + 1 < nch) {
          tmem_ld16(taddr + c + 16, nxt);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(B(kD2Empty + st));
        }
        if (act2) {
#pragma unroll
          for (int jj = 0; jj < 16; jj += 2) silu2_from_half(v[jj], v[jj + 1]);
        }
        if (nrow && !BEXP(p, 2)) {
          float* np = nrow + (size_t)c * hw;
#pragma unroll
          for (int jj = 0; jj < 16; ++jj)
            if (c + jj < p.nC) np[(size_t)jj * hw] = v[jj];
        }
        if (drow && !BEXP(p, 2)) {
          if (p.st256) {
            store_bf16x16(drow + c, v);
          } else {
            store_vec<__nv_bfloat16>(drow + c, v);
            store_vec<__nv_bfloat16>(drow + c + 8, v + 8);
          }
        }
      }
      advance();
      advance();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

struct B2bState {
  BParams p;
  int grid;
  size_t smem;
};

// chain = [3x3 (SiLU, stride 1 or 2) on the whole input region, 1x1 on its result], input and middle width 32 or 64; the result
// goes to the public NCHW tensor or to an NHWC bf16 slice.  ly_chain.reserved == 2 marks a stride-2 first stage.
bool conv_b2b_supported(const ly_op& op) {
  if (op.kind != LY_OP_CHAIN || op.dtype != LY_BF16 || !op.chain) return false;
  if ((op.nchw != nullptr) == (op.dst.ptr != nullptr)) return false;
  static const int enabled = getenv("LY_B2B") ? atoi(getenv("LY_B2B")) : 1;
  if (!enabled) return false;
  const ly_chain& ch = *op.chain;
  if (ch.n_stages != 2 || ch.n_regions != 2 || ch.n_in != 1) return false;
  const int stride = ch.reserved == 2 ? 2 : 1;
  const ly_chain_stage &s0 = ch.st[0], &s1 = ch.st[1];
  const int C = op.src.c, C1 = s0.cout, C2 = s1.cout;
  if (!(C == 32 || C == 64) || !(C1 == 32 || C1 == 64) || C2 % 16 || C2 < 16 || C2 > 64) return false;
  if (ch.region_c[0] != C || ch.region_c[1] != C1) return false;
  if (s0.k != 3 || !s0.act || s0.n_src != 1 || s0.src[0].region != 0 || s0.src[0].c0 != 0 || s0.src[0].c != C) return false;
  if (s0.dst.region != 1 || s0.dst.c0 != 0 || s0.dst.c != C1 || s0.res.region >= 0) return false;
  if (s1.k != 1 || s1.n_src != 1 || s1.src[0].region != 1 || s1.src[0].c0 != 0 || s1.src[0].c != C1) return false;
  if (s1.dst.region >= 0 || s1.res.region >= 0) return false;
  if (op.src.c0 % 8 || op.src.ctot % 8) return false;
  if (stride == 1 && op.src.W + 2 > 256) return false;
  if (stride == 2 && (op.src.H % 2 || op.src.W % 2 || op.src.W / 2 + 8 > 512)) return false;
  if (op.dst.ptr && (op.dst.c != C2 || op.dst.c0 % 8 || op.dst.ctot % 8 || op.dst.H != op.src.H / stride || op.dst.W != op.src.W / stride)) return false;
  return true;
}

int32_t conv_b2b_prepare(const ly_op& op, B2bState** out) {
  LY_CHECK_ARG(conv_b2b_supported(op), "conv_b2b: unsupported chain");
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("conv_b2b: cuTensorMapEncodeTiled entry point not available"); return LY_E_CUDA; }
  const ly_chain& ch = *op.chain;
  B2bState* st = new B2bState();
  BParams& p = st->p;
  memset(&p, 0, sizeof(p));
  const int stride = ch.reserved == 2 ? 2 : 1;
  const int Hi = op.src.H, Wi = op.src.W, H = Hi / stride, W = Wi / stride;     // H, W: the OUTPUT extent
  p.s2 = stride == 2;
  p.C = op.src.c; p.C1 = ch.st[0].cout; p.C2 = ch.st[1].cout;
  p.H = H; p.W = W; p.B = op.B;
  p.act2 = ch.st[1].act;
  p.exp = getenv("LY_TC_EXP") ? atoi(getenv("LY_TC_EXP")) : 0;
  p.b1 = ch.st[0].bias; p.b2 = ch.st[1].bias;
  p.nchw = op.nchw; p.nCtot = op.nchw_ctot; p.nC0 = op.nchw_c0; p.nC = op.nchw_c;
  p.dst = (__nv_bfloat16*)op.dst.ptr; p.dCtot = op.dst.ctot; p.dC0 = op.dst.c0;
  p.st256 = op.dst.ptr && op.dst.ctot % 16 == 0 && op.dst.c0 % 16 == 0 && reinterpret_cast<uintptr_t>(op.dst.ptr) % 32 == 0;
  p.w1_slab = (p.C1 * p.C * 2 + 1023) / 1024 * 1024;
  p.w2_bytes = (p.C2 * p.C1 * 2 + 1023) / 1024 * 1024;
  p.a2_stage = 128 * p.C1 * 2;
  // accumulator-row pitch: W + 2 padded columns (stride 1) or W + 1 rounded up to 8 pixels (stride 2: plane rows and the TMA
  // pieces of a row then start on 512-byte boundaries whatever the swizzle width)
  const int BW = p.s2 ? (W + 1 + 7) / 8 * 8 : W + 2;
  const int rowb = p.C * 2;
  const long long fixed = 1024 + 9LL * p.w1_slab + p.w2_bytes + 2LL * p.a2_stage + (p.C1 + p.C2) * 4 + 8 * kBarSlots;
  const int sms = sm_count();
  auto stage_bytes = [&](int R, int mt) -> long long {
    if (!p.s2) {
      const int need_rows = mt * 128 + 2 * BW + 2;       // the tap windows of the last M tile reach 2*BW + 2 rows past its rows
      return ((long long)std::max((R + 2) * BW, need_rows) * rowb + 1023) / 1024 * 1024;
    }
    const long long ev = ((long long)R * BW * rowb + 1023) / 1024 * 1024, od = ((long long)(R + 1) * BW * rowb + 1023) / 1024 * 1024;
    return 2 * ev + 2 * od;     // (reads past the last plane land in the next stage / the weights: garbage rows, never stored)
  };
  // band height: the cheapest schedule that fits two band stages (MMA issue ~50 cycles per instruction, TMA ~5 cycles per row)
  int best_r = 0;
  double best_cost = 1e30;
  for (int R = 1; R <= H && R + 2 <= 256; ++R) {
    const int mt = ((R - 1) * BW + W + 127) / 128;
    if (fixed + 2 * stage_bytes(R, mt) > (long long)kSmemMaxB) break;
    const long long units = (long long)((H + R - 1) / R) * op.B;
    const double rows_tma = p.s2 ? (double)(4 * R + 2) * BW : (double)(R + 2) * BW;
    const double unit_cost = std::max((double)mt * (9 * (p.C / 16) * 50.0 + (p.C1 / 16) * 50.0), rows_tma * 5.0);
    const double cost = (double)((units + sms - 1) / sms) * unit_cost;
    if (cost < best_cost) { best_cost = cost; best_r = R; }
  }
  if (!best_r) { delete st; set_error("conv_b2b: no band fits in shared memory"); return LY_E_ARG; }
  p.band_r = best_r; p.band_w = BW;
  p.band_mt = ((best_r - 1) * BW + W + 127) / 128;
  p.bands = (H + best_r - 1) / best_r;
  const long long units = (long long)p.bands * op.B;
  if (units > 0x7FFFFFFF || (unsigned long long)(units + 2 * sms) * p.bands >= (1ull << 32)) { delete st; set_error("conv_b2b: too many units"); return LY_E_ARG; }
  p.total_units = (int)units;
  p.mg_bw = (uint32_t)((1ull << 32) / (uint32_t)BW + 1);
  p.mg_bands = p.bands > 1 ? (uint32_t)((1ull << 32) / (uint32_t)p.bands + 1) : 0u;
  p.a_stage = (int)stage_bytes(best_r, p.band_mt);
  const uint32_t row16 = (uint32_t)rowb >> 4;
  if (!p.s2) {
    p.a_box = (best_r + 2) * BW * rowb;
    for (int t = 0; t < 9; ++t) p.tap_off16[t] = (uint32_t)((t / 3) * BW + (t % 3)) * row16;
  } else {
    const uint32_t ev = (uint32_t)(((long long)best_r * BW * rowb + 1023) / 1024 * 1024), od = (uint32_t)(((long long)(best_r + 1) * BW * rowb + 1023) / 1024 * 1024);
    p.plane16[0] = 0; p.plane16[1] = ev >> 4; p.plane16[2] = (2 * ev) >> 4; p.plane16[3] = (2 * ev + od) >> 4;     // EE, EO, OE, OO
    for (int t = 0; t < 9; ++t) {
      const int ky = t / 3, kx = t % 3;
      const int plane = (ky == 1 ? 0 : 2) + (kx == 1 ? 0 : 1);
      const int r_off = ky == 2 ? 1 : 0, j_off = kx == 0 ? 0 : 1;
      p.tap_off16[t] = p.plane16[plane] + (uint32_t)(r_off * BW + j_off) * row16;
    }
    // TMA pieces of a plane row: one box may traverse at most 256 input columns = 128 plane columns
    p.npieces = BW <= 128 ? 1 : 2;
    p.pw = p.npieces == 1 ? BW : (BW / 2 + 7) / 8 * 8;
    p.piece1 = BW - p.pw;
    if (p.pw > 128 || p.piece1 % 8) { delete st; set_error("conv_b2b: map too wide for the stride-2 band"); return LY_E_ARG; }
    p.a_box = (2 * best_r + 2 * (best_r + 1)) * p.npieces * p.pw * rowb;
  }
  st->smem = (size_t)fixed + 2 * (size_t)p.a_stage;
  if (st->smem > kSmemMaxB) { delete st; set_error("conv_b2b: tile does not fit in shared memory"); return LY_E_ARG; }
  if (st->smem < 120 * 1024) st->smem = 120 * 1024;     // one CTA per SM (all 512 TMEM columns)

  auto desc_hi = [](int kc) -> uint32_t {
    const int swz = kc == 64 ? 2 : 4;                                  // UMMA layout type: SWIZZLE_128B / SWIZZLE_64B
    const uint32_t sbo = (uint32_t)(8 * kc * 2) >> 4;
    return (sbo & 0x3FFFu) | (1u << 14) | ((uint32_t)swz << 29);
  };
  p.hi_a = desc_hi(p.C);
  p.hi_a2 = desc_hi(p.C1);
  p.idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.C1 >> 3) << 17) | ((128u >> 4) << 24);
  p.idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.C2 >> 3) << 17) | ((128u >> 4) << 24);
  auto tswz = [](int kc) { return kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B; };
  {
    char* base = (char*)op.src.ptr + (size_t)op.src.c0 * 2;
    cuuint64_t dims[4] = {(cuuint64_t)p.C, (cuuint64_t)Wi, (cuuint64_t)Hi, (cuuint64_t)op.B};
    cuuint64_t strides[3] = {(cuuint64_t)op.src.ctot * 2, (cuuint64_t)op.src.ctot * 2 * Wi, (cuuint64_t)op.src.ctot * 2 * Wi * Hi};
    cuuint32_t box[4] = {(cuuint32_t)p.C, (cuuint32_t)BW, (cuuint32_t)(best_r + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (p.s2) { box[1] = 2 * p.pw; box[2] = 2; estr[1] = 2; estr[2] = 2; }      // one plane row piece: pw columns of one row
    CUresult r = encode(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, tswz(p.C),
                        p.C == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete st; set_error("conv_b2b: cuTensorMapEncodeTiled(A) failed with %d", (int)r); return LY_E_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)9 * p.C, (cuuint64_t)p.C1};
    cuuint64_t strides[1] = {dims[0] * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.C, (cuuint32_t)p.C1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.tmW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ch.st[0].w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        tswz(p.C), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete st; set_error("conv_b2b: cuTensorMapEncodeTiled(W1) failed with %d", (int)r); return LY_E_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.C1, (cuuint64_t)p.C2};
    cuuint64_t strides[1] = {dims[0] * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.C1, (cuuint32_t)p.C2};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.tmW2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ch.st[1].w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        tswz(p.C1), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete st; set_error("conv_b2b: cuTensorMapEncodeTiled(W2) failed with %d", (int)r); return LY_E_CUDA; }
  }
  st->grid = p.total_units < sms ? p.total_units : sms;
  static std::atomic<unsigned long long> attr_devs{0};
  if (first_on_device(attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(conv_b2b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMaxB);
    if (e != cudaSuccess) { delete st; set_error("conv_b2b: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return LY_E_CUDA; }
  }
  static const int debug = getenv("LY_TC_DEBUG") ? atoi(getenv("LY_TC_DEBUG")) : 0;
  if (debug)
    fprintf(stderr, "[conv_b2b] s%d out %dx%d C %d -> %d -> %d B %d: band R %d pitch %d mt %d bands %d units %d a_stage %d pieces %d x %d smem %zu\n", stride,
            H, W, p.C, p.C1, p.C2, op.B, p.band_r, BW, p.band_mt, p.bands, p.total_units, p.a_stage, p.npieces, p.pw, st->smem);
  *out = st;
  return LY_OK;
}

int32_t conv_b2b_launch(const B2bState* st, float* nchw_override, cudaStream_t s) {
  if (nchw_override) {
    BParams p = st->p;
    p.nchw = nchw_override;
    launch_k(conv_b2b_kernel, dim3(st->grid), dim3(kThreadsB), st->smem, s, p);
  } else {
    launch_k(conv_b2b_kernel, dim3(st->grid), dim3(kThreadsB), st->smem, s, st->p);
  }
  return post_launch("conv_b2b");
}

void conv_b2b_free(B2bState* st) { delete st; }

}  // namespace ly
