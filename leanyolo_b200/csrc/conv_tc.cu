// BN-folded NHWC bf16 implicit-GEMM convolution on Blackwell tensor cores.
//
//   D[M = B*Ho*Wo, N = Cout] = A[M, K = k*k*Cin] * W[N, K]^T  (+bias, SiLU, +residual)
//
// * A is never materialised.  The producer warp issues 4-D TMA tile loads (C, W, H, B) of
//   the NHWC activation; out-of-bounds (padding) elements are zero-filled by the TMA unit and
//   a stride-2 conv is the tensor map's elementStrides = 2.  The tile of 128 output pixels is
//   a (tw x th x tb) brick chosen per layer so that feature maps tile without waste; 1x1/s1
//   layers collapse to a flat [M, C] matrix.
//     classic mode : one box per (filter tap, channel block)          -> 9 loads / block for 3x3
//     halo mode    : 3x3, brick 8 x 16: one box of 8 x 18 pixels per (kx, channel
//                    block); the three ky taps are the same shared-memory tile viewed 8 rows
//                    (= one swizzle atom) further down -> 3 loads / block.  Stride 2: the box
//                    holds 8 (every other) columns x 33 rows and the A descriptor's 8-row
//                    group stride is doubled, so output row oy reads stored row 2*oy + ky.  The TMA unit is
//                    request-rate bound on these strided 64/128-byte rows (measured ~5 cycles
//                    per row), so this is what moves the 3x3 layers towards the MMA roofline.
//     band mode    : 3x3 stride 1, maps up to 254 pixels wide: ONE box per channel block holds
//                    R+2 full padded image rows ((W+2) x (R+2) pixels, row index = y*(W+2)+x).
//                    Output pixel m = oy*(W+2)+ox of the band reads row m + ky*(W+2) + kx for tap
//                    (ky, kx): all nine taps are the SAME tile viewed a few rows further down
//                    (the 128B/64B/32B swizzle is a function of the absolute shared-memory
//                    address, so a descriptor may start at any row), and the M = 128 tiles are
//                    consecutive runs of m.  2 of every W+2 accumulator rows are discarded, but
//                    every input byte crosses the TMA unit (R+2)/R times instead of 3.4 times.
//     fold (band)  : Cout <= 80 with resident weights.  An M=128 MMA costs about the same for N = 16..64 (the
//                    A operand's 4 KB dominate its shared-memory reads), so the nine taps of a thin layer run at a
//                    fraction of the tensor pipe.  Here the three kx taps become ONE MMA with N = 3*Cout:
//                    D[m][kx*Cout + co] = sum_{ky,ci} A[m + ky*(W+2)][ci] * W[co][ky][kx][ci]  (3x fewer MMAs, each
//                    reading A once), and the epilogue adds the three column groups of rows m, m+1, m+2:
//                    out[m][co] = D[m][co] + D[m+1][Cout+co] + D[m+2][2*Cout+co] (warp shuffles; the two rows that
//                    cross a TMEM lane quarter go through a tiny shared-memory exchange).  M tiles advance by 126 rows.
// * MMA: tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N (<=256), K=16 per
//   instruction, bf16 x bf16 -> fp32 accumulators in TMEM (double-buffered so the epilogue of
//   tile i overlaps the MMAs of tile i+1).  Operands are K-major in shared memory with the
//   32/64/128-byte swizzle that matches the channel block (16/32/64).
// * Weights: resident in shared memory for the whole kernel when they fit (<= 96 KB),
//   otherwise streamed through their own ring, one box per (tap, channel block).
// * concat / split / residual are epilogue addressing: the output goes to a channel slice
//   (dC0, dCtot) of the consumer's concat buffer, the input tensor map starts at the
//   producer's channel offset, the shortcut is read in the epilogue.
// * Persistent: one CTA per SM, static round-robin over (pixel-brick, n-tile) tiles.
//   Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..17 = epilogue (thin layers are
//   bound by epilogue latency, not bandwidth: 16 warps keep 4 per scheduler to overlap TMEM loads,
//   MUFU and stores of different chunks).
//
// Reference semantics: leanyolo/models/yolov10/layers.py:51-88 (Conv = conv+BN+SiLU).
#include <cuda.h>
#include <stdlib.h>
#include <algorithm>
#include <string.h>
#include "common.cuh"
#include "tma.cuh"
#include "tc.cuh"

namespace ly {

namespace {

constexpr int kEpiWarps = 16;    // 4 TMEM lane quarters x 4 column groups
constexpr int kThreads = 64 + 32 * kEpiWarps;    // TMA warp + MMA warp + epilogue warps
constexpr int kMaxCout = 1024;   // bias vector staged in shared memory
constexpr int kMaxStages = 12;   // per ring
constexpr int kMaxAcc = 8;       // TMEM accumulator stages (512 columns / N)
constexpr int kBarSlots = 4 * kMaxStages + 2 * kMaxAcc + 4;   // 8-byte slots of the barrier block
constexpr uint32_t kSmemBudget = 216 * 1024;

struct Params {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmD;         // output tensor (TMA-store epilogue): box = 32 channels x the 32 pixels of one epilogue warp
  // tiling
  int tw, th, tb;
  int tiles_w, tiles_h, tiles_b, tiles_n, total_tiles;
  int Wo, Ho, Bo;          // (possibly collapsed) output extents used for tiling / masking
  int hw_real;             // Ho*Wo of the real tensor (NCHW addressing)
  uint32_t mg_n, mg_w, mg_h;   // magic multipliers: x / d == __umulhi(x, mg) for the tile counts used (0: d == 1)
  int k, stride, pad;
  int kc, kc_blocks, num_kb;
  int halo;                // 1: halo mode (3 taps share one A box)  2: band mode (all 9 taps share one box)
  int band_r, band_w, band_mt, bands;   // band mode: output rows per band, padded row width W+2, M tiles per band, bands per image
  uint32_t mg_bw, mg_bands;
  int tpa, num_ka;         // taps per A stage, A loads per tile
  int a_tap_stride;        // bytes between the windows of successive taps inside an A stage
  int block_n, tmem_cols;
  int a_stages, b_stages;
  int a_stage, b_stage;    // stage strides in bytes
  int a_box, b_box;        // bytes one TMA box delivers
  int b_resident;
  int res_slot;            // bytes of private smem per epilogue thread for the prefetched shortcut (0 = direct loads)
  uint32_t idesc, desc_hi, desc_hi_a;   // desc_hi_a: A operand (its 8-row group stride doubles in stride-2 halo mode)
  int act;
  __nv_bfloat16* dst; int dCtot, dC0;
  const __nv_bfloat16* res; int rCtot, rC0;
  const __nv_bfloat16* up; int uCtot, uC0, uH, uW, Wreal;   // half-resolution pre-activation addend (upsampled x2 on the fly)
  const float* bias;
  float* nchw; int nCtot, nC0, nC;
  int cin_pad;
  int rev;                 // walk the tiles / units last to first (see g_reverse)
  int st256;               // NHWC rows are 32-byte aligned: one 256-bit store per 16-channel chunk
  int pair;                // streamed weights: two pixel tiles (2q, 2q+1) share every weight k-block (4 accumulators in TMEM)
  int fold;                // band mode, Cout <= 80: the three kx taps are folded into the MMA's N dimension (N = 3 * Cout)
  uint32_t idesc_fold;     // instruction descriptor with N = 3 * block_n
  uint32_t exch_off;       // fold: byte offset (from the barrier block) of the epilogue's boundary-row exchange slots
  int xpre;                // epilogue: cross-tile TMEM prefetch (LY_TC_XPRE=1, default off)
  int acc;                 // TMEM accumulator stages (each block_n columns; pair: 2 * block_n, fold: 3 * block_n)
  int epi_groups;          // 0: column-parallel epilogue; 4: tile groups of the tile-parallel epilogue (needs acc >= 4)
  int ld256;               // shortcut / addend rows are 32-byte aligned: one 256-bit load per 16-channel chunk
  int ts;                  // TMA-store epilogue: staging slots per epilogue warp (0 = direct stores)
  uint32_t stg_off;        // byte offset (from the barrier block, 1024-aligned) of the staging slots: [warp][slot][32 rows x 64 B]
  int exp;                 // -DLY_TC_EXP builds only (LY_TC_EXP=mask): 1 skip the MMAs, 2 skip the epilogue's work, 4 skip the TMA loads after the first pass
};

// Knock-out experiments (-DLY_TC_EXP): which role bounds a layer?  Results are garbage, only the time is meaningful.
#ifdef LY_TC_EXP
#define EXP_ON(p, bit) (((p).exp & (bit)) != 0)
#else
#define EXP_ON(p, bit) false
#endif

// Optional per-role cycle accounting (-DLY_TC_PROFILE): CTA 0 prints where each role waited.
#ifdef LY_TC_PROFILE
#define PROF_DECL(name) long long name = 0
#define PROF_T0() const long long t0__ = clock64()
#define PROF_ADD(name) name += clock64() - t0__
#else
#define PROF_DECL(name)
#define PROF_T0()
#define PROF_ADD(name)
#endif

// barrier addresses inside the barrier block (see the carve-up in the kernel)
__device__ __forceinline__ uint32_t bar_afull(uint32_t bb, int s) { return bb + 8u * s; }
__device__ __forceinline__ uint32_t bar_aempty(uint32_t bb, int s) { return bb + 8u * (kMaxStages + s); }
__device__ __forceinline__ uint32_t bar_bfull(uint32_t bb, int s) { return bb + 8u * (2 * kMaxStages + s); }
__device__ __forceinline__ uint32_t bar_bempty(uint32_t bb, int s) { return bb + 8u * (3 * kMaxStages + s); }
__device__ __forceinline__ uint32_t bar_tfull(uint32_t bb, int s) { return bb + 8u * (4 * kMaxStages + s); }
__device__ __forceinline__ uint32_t bar_tempty(uint32_t bb, int s) { return bb + 8u * (4 * kMaxStages + kMaxAcc + s); }
__device__ __forceinline__ uint32_t bar_bres(uint32_t bb) { return bb + 8u * (4 * kMaxStages + 2 * kMaxAcc); }

// tile index -> (n tile, brick) without integer division (magic multipliers from the host)
__device__ __forceinline__ int phys(const Params& p, int t) { return p.rev ? p.total_tiles - 1 - t : t; }
__device__ __forceinline__ void split_tile(const Params& p, int tile_idx, int& nt, int& wt, int& ht, int& bt) {
  uint32_t t = (uint32_t)phys(p, tile_idx);   // multiplier 0 encodes a divisor of 1
  uint32_t qn = p.mg_n ? __umulhi(t, p.mg_n) : t; nt = (int)(t - qn * (uint32_t)p.tiles_n); t = qn;
  uint32_t qw = p.mg_w ? __umulhi(t, p.mg_w) : t; wt = (int)(t - qw * (uint32_t)p.tiles_w); t = qw;
  uint32_t qh = p.mg_h ? __umulhi(t, p.mg_h) : t; ht = (int)(t - qh * (uint32_t)p.tiles_h); bt = (int)qh;
}

// The single MMA-issuing thread.  Specialised on the k-steps per channel block, the taps per
// A stage and weight residency so that the issue loop is straight-line code: a descriptor is
// (constant high word | 14-bit address field), and stepping K by 16 elements or moving to the
// next tap / stage is an integer add on the low word.  (Measured: a clean issue loop sustains
// one M=128 MMA every ~50 cycles for N <= 64; anything slower is issue overhead.)
template <int KSTEPS, int TPA, bool BRES>
__device__ __forceinline__ void mma_role(const Params& p, uint32_t a_base, uint32_t b_base, uint32_t bb, uint32_t tmem_base) {
  const uint32_t hi = p.desc_hi, hi_a = p.desc_hi_a, idesc = p.idesc;
  const int num_ka = p.num_ka, a_stages = p.a_stages, b_stages = p.b_stages, total = p.total_tiles;
  const uint32_t a_stage16 = (uint32_t)p.a_stage >> 4, b_stage16 = (uint32_t)p.b_stage >> 4, tap16 = (uint32_t)p.a_tap_stride >> 4;
  const uint32_t a_lo0 = (a_base >> 4) | (1u << 16), b_lo0 = (b_base >> 4) | (1u << 16);
  const uint32_t block_n = (uint32_t)p.block_n;
  const int acc = p.acc;
  if (BRES) {
    mbar_wait_spin(bar_bres(bb), 0);
    tc_fence_after();
  }
  int sa = 0, sb = 0, as = 0;
  uint32_t pa = 0, pb = 0, aphase = 0;
  PROF_DECL(w_afull); PROF_DECL(w_tempty);
#ifdef LY_TC_PROFILE
  const long long istart = clock64();
#endif
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    { PROF_T0(); mbar_wait_spin(bar_tempty(bb, as), aphase ^ 1u); PROF_ADD(w_tempty); }
    tc_fence_after();
    const uint32_t d_tmem = tmem_base + (uint32_t)as * block_n;
    for (int ka = 0; ka < num_ka; ++ka) {
      { PROF_T0(); mbar_wait_spin(bar_afull(bb, sa), pa); PROF_ADD(w_afull); }
      tc_fence_after();
      const uint32_t alo = a_lo0 + (uint32_t)sa * a_stage16;
#pragma unroll
      for (int tt = 0; tt < TPA; ++tt) {
        uint32_t blo;
        if (BRES) {
          blo = b_lo0 + (uint32_t)(ka * TPA + tt) * b_stage16;
        } else {
          mbar_wait_spin(bar_bfull(bb, sb), pb);
          tc_fence_after();
          blo = b_lo0 + (uint32_t)sb * b_stage16;
        }
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
          const uint64_t da = ((uint64_t)hi_a << 32) | (uint64_t)(alo + tt * tap16 + 2 * kk);
          const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)(blo + 2 * kk);
          if (!EXP_ON(p, 1)) umma_bf16(d_tmem, da, db, idesc, (tt | kk) != 0 ? 1u : (ka != 0 ? 1u : 0u));
        }
        if (!BRES) {
          umma_commit(bar_bempty(bb, sb));
          if (++sb == b_stages) { sb = 0; pb ^= 1u; }
        }
      }
      umma_commit(bar_aempty(bb, sa));
      if (++sa == a_stages) { sa = 0; pa ^= 1u; }
    }
    umma_commit(bar_tfull(bb, as));
    if (++as == acc) { as = 0; aphase ^= 1u; }
  }
#ifdef LY_TC_PROFILE
  if (blockIdx.x == 0) printf("[tc prof] issuer: total %lld wait_afull %lld wait_tempty %lld\n", clock64() - istart, w_afull, w_tempty);
#endif
}

// Pair mode issuer (streamed weights, one n tile, 4 * N <= 512 TMEM columns).  The 3x3 N = 128 layers stream
// 295 KB of weights per 128-pixel tile; at 10 TB/s of L2 -> SM bandwidth that, not the tensor pipe, bounded
// them (3x3 128->128 @40^2: 3584 tiles x 349 KB in 0.12 ms = 10.4 TB/s).  Every weight k-block now feeds the
// MMAs of TWO pixel tiles (consecutive A stages), halving the weight traffic per output.
template <int KSTEPS, int TPA>
__device__ __forceinline__ void mma_role_pair(const Params& p, uint32_t a_base, uint32_t b_base, uint32_t bb, uint32_t tmem_base) {
  const uint32_t hi = p.desc_hi, hi_a = p.desc_hi_a, idesc = p.idesc;
  const int num_ka = p.num_ka, a_stages = p.a_stages, b_stages = p.b_stages, total = p.total_tiles;
  const uint32_t a_stage16 = (uint32_t)p.a_stage >> 4, b_stage16 = (uint32_t)p.b_stage >> 4, tap16 = (uint32_t)p.a_tap_stride >> 4;
  const uint32_t a_lo0 = (a_base >> 4) | (1u << 16), b_lo0 = (b_base >> 4) | (1u << 16);
  const uint32_t block_n = (uint32_t)p.block_n;
  const int acc = p.acc;
  const int pairs = (total + 1) >> 1;
  int sa = 0, sb = 0, as = 0;
  uint32_t pa = 0, pb = 0, aphase = 0;
  for (int q = blockIdx.x; q < pairs; q += gridDim.x) {
    const bool two = 2 * q + 1 < total;
    mbar_wait_spin(bar_tempty(bb, as), aphase ^ 1u);
    tc_fence_after();
    const uint32_t d0 = tmem_base + (uint32_t)(2 * as) * block_n, d1 = d0 + block_n;
    for (int ka = 0; ka < num_ka; ++ka) {
      const int s0 = sa;
      mbar_wait_spin(bar_afull(bb, s0), pa);
      if (++sa == a_stages) { sa = 0; pa ^= 1u; }
      const int s1 = sa;
      if (two) {
        mbar_wait_spin(bar_afull(bb, s1), pa);
        if (++sa == a_stages) { sa = 0; pa ^= 1u; }
      }
      tc_fence_after();
      const uint32_t alo0 = a_lo0 + (uint32_t)s0 * a_stage16, alo1 = a_lo0 + (uint32_t)s1 * a_stage16;
#pragma unroll
      for (int tt = 0; tt < TPA; ++tt) {
        mbar_wait_spin(bar_bfull(bb, sb), pb);
        tc_fence_after();
        const uint32_t blo = b_lo0 + (uint32_t)sb * b_stage16;
        const uint32_t acc = (tt != 0 || ka != 0) ? 1u : 0u;
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
          const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)(blo + 2 * kk);
          umma_bf16(d0, ((uint64_t)hi_a << 32) | (uint64_t)(alo0 + tt * tap16 + 2 * kk), db, idesc, kk != 0 ? 1u : acc);
        }
        if (two) {
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk) {
            const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)(blo + 2 * kk);
            umma_bf16(d1, ((uint64_t)hi_a << 32) | (uint64_t)(alo1 + tt * tap16 + 2 * kk), db, idesc, kk != 0 ? 1u : acc);
          }
        }
        umma_commit(bar_bempty(bb, sb));
        if (++sb == b_stages) { sb = 0; pb ^= 1u; }
      }
      umma_commit(bar_aempty(bb, s0));
      if (two) umma_commit(bar_aempty(bb, s1));
    }
    umma_commit(bar_tfull(bb, as));
    if (++as == acc) { as = 0; aphase ^= 1u; }
  }
}

// Band mode issuer.  Unit = one band of one image; its kc_blocks A stages stay resident while every
// (n tile, M tile) of the band is accumulated: D[mt] = sum over (cb, tap) A[stage cb, rows mt*128 + tap offset ...] * W[tap, cb].
template <int KSTEPS, bool BRES>
__device__ __forceinline__ void mma_role_band(const Params& p, uint32_t a_base, uint32_t b_base, uint32_t bb, uint32_t tmem_base) {
  const uint32_t hi = p.desc_hi, idesc = p.idesc;
  const int a_stages = p.a_stages, b_stages = p.b_stages, total = p.total_tiles, kcb = p.kc_blocks;
  const uint32_t a_stage16 = (uint32_t)p.a_stage >> 4, b_stage16 = (uint32_t)p.b_stage >> 4;
  const uint32_t row16 = (uint32_t)p.kc >> 3;                       // one pixel row of the tile, in 16-byte units
  const uint32_t a_lo0 = (a_base >> 4) | (1u << 16), b_lo0 = (b_base >> 4) | (1u << 16);
  const uint32_t block_n = (uint32_t)p.block_n;
  const int acc = p.acc;
  const uint32_t bw16 = (uint32_t)p.band_w * row16;
  if (BRES) {
    mbar_wait_spin(bar_bres(bb), 0);
    tc_fence_after();
  }
  int sa = 0, sb = 0, as = 0;
  uint32_t pa = 0, pb = 0, aphase = 0;
  PROF_DECL(w_afull); PROF_DECL(w_tempty); PROF_DECL(t_issue); PROF_DECL(t_commit);
#ifdef LY_TC_PROFILE
  const long long istart = clock64();
#endif
  for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
    {  // all channel blocks of this band have landed
      PROF_T0();
      int s = sa; uint32_t ph = pa;
      for (int cb = 0; cb < kcb; ++cb) {
        mbar_wait_spin(bar_afull(bb, s), ph);
        if (++s == a_stages) { s = 0; ph ^= 1u; }
      }
      tc_fence_after();
      PROF_ADD(w_afull);
    }
    for (int nt = 0; nt < p.tiles_n; ++nt)
      for (int mt = 0; mt < p.band_mt; ++mt) {
        { PROF_T0(); mbar_wait_spin(bar_tempty(bb, as), aphase ^ 1u); PROF_ADD(w_tempty); }
        tc_fence_after();
#ifdef LY_TC_PROFILE
        const long long ti0 = clock64();
#endif
        const uint32_t d_tmem = tmem_base + (uint32_t)as * block_n;
        int s = sa;
        for (int cb = 0; cb < kcb; ++cb) {
          const uint32_t alo = a_lo0 + (uint32_t)s * a_stage16 + (uint32_t)(mt * 128) * row16;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t blo;
            if (BRES) {
              blo = b_lo0 + (uint32_t)(tap * kcb + cb) * b_stage16;
            } else {
              mbar_wait_spin(bar_bfull(bb, sb), pb);
              tc_fence_after();
              blo = b_lo0 + (uint32_t)sb * b_stage16;
            }
            const uint32_t at = alo + (uint32_t)(tap / 3) * bw16 + (uint32_t)(tap % 3) * row16;
#pragma unroll
            for (int kk = 0; kk < KSTEPS; ++kk) {
              const uint64_t da = ((uint64_t)hi << 32) | (uint64_t)(at + 2 * kk);
              const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)(blo + 2 * kk);
              if (!EXP_ON(p, 1)) umma_bf16(d_tmem, da, db, idesc, (tap | kk) != 0 ? 1u : (cb != 0 ? 1u : 0u));
            }
            if (!BRES) {
              umma_commit(bar_bempty(bb, sb));
              if (++sb == b_stages) { sb = 0; pb ^= 1u; }
            }
          }
          if (++s == a_stages) s = 0;
        }
#ifdef LY_TC_PROFILE
        const long long ti1 = clock64();
        t_issue += ti1 - ti0;
#endif
        umma_commit(bar_tfull(bb, as));
#ifdef LY_TC_PROFILE
        t_commit += clock64() - ti1;
#endif
        if (++as == acc) { as = 0; aphase ^= 1u; }
      }
    for (int cb = 0; cb < kcb; ++cb) {   // the band's tiles are free once everything issued so far has completed
      umma_commit(bar_aempty(bb, sa));
      if (++sa == a_stages) { sa = 0; pa ^= 1u; }
    }
  }
#ifdef LY_TC_PROFILE
  if (blockIdx.x == 0) printf("[tc prof] band issuer: total %lld wait_afull %lld wait_tempty %lld issue %lld commit %lld\n", clock64() - istart, w_afull, w_tempty, t_issue, t_commit);
#endif
}

// Band mode with the kx taps folded into N (resident weights only): per M tile 3 * kc_blocks * KSTEPS MMAs of N = 3*Cout.
// Weight slab (ky, cb) = [3*Cout rows (kx-major) x kc] at index ky*kcb + cb; M tiles advance by 126 rows (the epilogue
// needs rows m+1, m+2 of the same accumulator).
template <int KSTEPS>
__device__ __forceinline__ void mma_role_fold(const Params& p, uint32_t a_base, uint32_t b_base, uint32_t bb, uint32_t tmem_base) {
  const uint32_t hi = p.desc_hi, idesc = p.idesc_fold;
  const int a_stages = p.a_stages, total = p.total_tiles, kcb = p.kc_blocks;
  const uint32_t a_stage16 = (uint32_t)p.a_stage >> 4, b_stage16 = (uint32_t)p.b_stage >> 4;
  const uint32_t row16 = (uint32_t)p.kc >> 3;
  const uint32_t a_lo0 = (a_base >> 4) | (1u << 16), b_lo0 = (b_base >> 4) | (1u << 16);
  const uint32_t acc_cols = 3u * (uint32_t)p.block_n;
  const int acc = p.acc;
  const uint32_t bw16 = (uint32_t)p.band_w * row16;
  mbar_wait_spin(bar_bres(bb), 0);
  tc_fence_after();
  int sa = 0, as = 0;
  uint32_t pa = 0, aphase = 0;
  for (int unit = blockIdx.x; unit < total; unit += gridDim.x) {
    {
      int s = sa; uint32_t ph = pa;
      for (int cb = 0; cb < kcb; ++cb) {
        mbar_wait_spin(bar_afull(bb, s), ph);
        if (++s == a_stages) { s = 0; ph ^= 1u; }
      }
      tc_fence_after();
    }
    for (int mt = 0; mt < p.band_mt; ++mt) {
      mbar_wait_spin(bar_tempty(bb, as), aphase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)as * acc_cols;
      int s = sa;
      for (int cb = 0; cb < kcb; ++cb) {
        const uint32_t alo = a_lo0 + (uint32_t)s * a_stage16 + (uint32_t)(mt * 126) * row16;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          const uint32_t blo = b_lo0 + (uint32_t)(ky * kcb + cb) * b_stage16;
          const uint32_t at = alo + (uint32_t)ky * bw16;
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk) {
            const uint64_t da = ((uint64_t)hi << 32) | (uint64_t)(at + 2 * kk);
            const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)(blo + 2 * kk);
            umma_bf16(d_tmem, da, db, idesc, (ky | kk) != 0 ? 1u : (cb != 0 ? 1u : 0u));
          }
        }
        if (++s == a_stages) s = 0;
      }
      umma_commit(bar_tfull(bb, as));
      if (++as == acc) { as = 0; aphase ^= 1u; }
    }
    for (int cb = 0; cb < kcb; ++cb) {
      umma_commit(bar_aempty(bb, sa));
      if (++sa == a_stages) { sa = 0; pa ^= 1u; }
    }
  }
}

// ------------------------------------------------------------------------------- epilogue
// 16 warps: warp -> (TMEM lane quarter q = warp % 4, column group cg = (warp - 2) / 4).  A thread
// owns one pixel row; in round i the four warps of a quarter take the 16-column chunks 4i+cg.
// The TMEM load of round i+1 is in flight while round i is activated and stored; the accumulator
// stage is released as soon as the last chunk sits in registers.
// Thin layers are bound by THIS instruction stream (ncu: ~300 warp-instructions per warp and tile
// of which ~60 are the activation math), so the role is specialised at compile time on
//   MAP   pixel mapping: 0 flat (1x1/s1: pixel = tile*128 + row), 1 brick / halo, 2 band
//   ADD   epilogue addend: 0 none, 1 shortcut (prefetched slots), 2 half-resolution addend
//         (prefetched slots), 3 shortcut (direct loads), 4 half-resolution addend (direct loads)
//   NCHW  public fp32 NCHW output instead of the NHWC bf16 buffer
// and every address is 32-bit pixel index x 32-bit pitch (one IMAD.WIDE).
// (Measured dead ends for the NHWC store: a TMA store from a swizzled staging tile and a
// shared-memory transpose to 128-byte coalesced stores were both slower than storing
// 2 x 16 bytes per lane straight from the TMEM layout: the extra barrier per tile costs more
// than the partial-sector writes, and the TMA unit is already row-rate bound on the loads.)
template <int MAP, int ADD, bool NCHW, bool FOLD = false>
__device__ __forceinline__ void epilogue_role(const Params& p, uint8_t* smem_raw, uint32_t bar_base, uint32_t tmem_base, const float* s_bias) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;
  const int cg = (warp - 2) >> 2;
  const int nchunks = p.block_n >> 4;
  const int rounds = (nchunks + 3) >> 2;
  const uint32_t row = (uint32_t)(q * 32 + lane);
  const uint32_t dw = row % (uint32_t)p.tw;
  const uint32_t dh = (row / (uint32_t)p.tw) % (uint32_t)p.th;
  const uint32_t db = row / (uint32_t)(p.tw * p.th);
  const bool act = p.act != 0;
  const float pre = act ? 0.5f : 1.0f;   // SiLU(x) = h + h*tanh(h), h = x/2: the halving rides on the bias FMA
  constexpr bool kSlots = ADD == 1 || ADD == 2;
  constexpr bool kUp = ADD == 2 || ADD == 4;
  constexpr bool kRes = ADD == 1 || ADD == 3;

  // Work items in the order the issuer accumulates them: one tile per step of the persistent loop
  // (flat / brick), or (unit, n tile, M tile) in band mode.  `it_*` is the state of the next item.
  struct Loc { bool more, valid, first, last; uint32_t lin; int n0, sub; };
  int it_unit = blockIdx.x, it_nt = 0, it_mt = 0;
  const bool pair = MAP == 1 && p.pair != 0;
  auto next_item = [&]() -> Loc {
    Loc L;
    L.more = pair ? 2 * it_unit < p.total_tiles : it_unit < p.total_tiles;
    L.valid = false; L.lin = 0; L.n0 = 0; L.sub = 0; L.first = L.last = true;
    if (!L.more) return L;
    if (MAP == 2) {
      const uint32_t pu = (uint32_t)phys(p, it_unit);
      const uint32_t b = p.mg_bands ? __umulhi(pu, p.mg_bands) : pu;
      const uint32_t bd = pu - b * (uint32_t)p.bands;
      const uint32_t m = (uint32_t)it_mt * (FOLD ? 126u : 128u) + row;
      const uint32_t oy = __umulhi(m, p.mg_bw), ox = m - oy * (uint32_t)p.band_w;
      const uint32_t h = bd * (uint32_t)p.band_r + oy;
      L.valid = ox < (uint32_t)p.Wo && oy < (uint32_t)p.band_r && h < (uint32_t)p.Ho && (!FOLD || row < 126u);
      L.lin = (b * (uint32_t)p.Ho + h) * (uint32_t)p.Wo + ox;
      L.n0 = it_nt * p.block_n;
      if (++it_mt == p.band_mt) { it_mt = 0; if (++it_nt == p.tiles_n) { it_nt = 0; it_unit += gridDim.x; } }
    } else if (MAP == 0) {
      const uint32_t t = (uint32_t)phys(p, it_unit);
      const uint32_t qn = p.mg_n ? __umulhi(t, p.mg_n) : t;
      L.n0 = (int)(t - qn * (uint32_t)p.tiles_n) * p.block_n;
      L.lin = qn * 128u + row;
      L.valid = L.lin < (uint32_t)p.Wo;
      it_unit += gridDim.x;
    } else {
      int nt, wt, ht, bt;
      int tile = it_unit;
      if (pair) {   // it_unit counts PAIRS: tiles 2q (accumulator 0) and 2q + 1 (accumulator 1)
        tile = 2 * it_unit + it_mt;
        L.sub = it_mt; L.first = it_mt == 0; L.last = it_mt == 1 || tile + 1 >= p.total_tiles;
      }
      split_tile(p, tile, nt, wt, ht, bt);
      const uint32_t w = (uint32_t)(wt * p.tw) + dw, h = (uint32_t)(ht * p.th) + dh, b = (uint32_t)(bt * p.tb) + db;
      L.valid = w < (uint32_t)p.Wo && h < (uint32_t)p.Ho && b < (uint32_t)p.Bo;
      L.lin = (b * (uint32_t)p.Ho + h) * (uint32_t)p.Wo + w;
      L.n0 = nt * p.block_n;
      if (pair && !L.last) it_mt = 1;
      else { it_mt = 0; it_unit += gridDim.x; }
    }
    return L;
  };
  // half-resolution source row of output pixel `lin` (the conv is 1x1, so lin is the real pixel index)
  auto up_row = [&](uint32_t lin, int n0) -> const __nv_bfloat16* {
    const uint32_t ub = lin / (uint32_t)p.hw_real, urem = lin - ub * (uint32_t)p.hw_real;
    const uint32_t uh = urem / (uint32_t)p.Wreal, uw = urem - uh * (uint32_t)p.Wreal;
    return p.up + (size_t)((ub * (uint32_t)p.uH + (uh >> 1)) * (uint32_t)p.uW + (uw >> 1)) * (uint32_t)p.uCtot + p.uC0 + n0;
  };
  // Shortcut / addend operand: each thread cp.async's its own 32 bytes per chunk of the NEXT tile
  // into a private shared-memory slot while it works on the current tile, so the DRAM latency of
  // the uncoalesced read is off the epilogue's critical path (measured: a C2f bottleneck at 160^2
  // spent 35 % of its samples waiting for that load).
  // Slot layout: piece-major ([stage][2*round + half][epilogue thread] x 16 bytes), so the 32 lanes of a
  // warp touch 32 consecutive 16-byte words: thread-major slots (64-byte stride per lane) cost a 4-way
  // bank conflict on every cp.async write and LDS.128 read (ncu: 11.8 M conflicts, 28 % short-scoreboard stalls).
  const uint32_t r_base = bar_base + 8u * kBarSlots;
  const uint32_t rslot = r_base + (uint32_t)(threadIdx.x - 64) * 16u;
  constexpr uint32_t kPiece = 32u * kEpiWarps * 16u;                    // bytes between pieces
  const uint32_t rstage = (uint32_t)(32 * kEpiWarps) * p.res_slot;
  // Half-resolution addend: horizontally adjacent output pixels (w even, w + 1) share their source pixel.  In
  // the flat mapping they are adjacent lanes (the row width is even), so the odd lane skips its own request
  // and reads the even lane's slot: half the scattered 32-byte L2 requests of this epilogue.
  auto up_dup = [&](uint32_t lin) -> bool { return kUp && MAP == 0 && lane > 0 && (((lin % (uint32_t)p.Wreal) & 1u) != 0u); };
  auto res_prefetch = [&](const Loc& L, uint32_t dst) {
    if (L.more && L.valid && !up_dup(L.lin)) {
      const __nv_bfloat16* rr = kUp ? up_row(L.lin, L.n0) : p.res + (size_t)L.lin * (uint32_t)p.rCtot + p.rC0 + L.n0;
      for (int i = 0, ch = cg; ch < nchunks; ++i, ch += 4) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (2 * i) * kPiece), "l"(rr + ch * 16) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (2 * i + 1) * kPiece), "l"(rr + ch * 16 + 8) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int as = 0;
  const int acc = p.acc;
  uint32_t aphase = 0, rs = 0;
  // fold: exchange slots for the rows that cross a TMEM lane quarter: [parity][cg][q][48 floats]
  float* exch = reinterpret_cast<float*>(smem_raw + ((bar_base + p.exch_off) - smem_u32(smem_raw)));
  uint32_t xpar = 0;
  PROF_DECL(w_tfull);
#ifdef LY_TC_PROFILE
  const long long estart = clock64();
#endif
  Loc cur = next_item();
  if (kSlots) res_prefetch(cur, rslot);
  // Cross-tile prefetch: when a warp finishes its last TMEM load of a tile and the NEXT tile's accumulator is already
  // complete, its first chunk is requested right away, so that the TMEM round trip (and the barrier poll) overlap the
  // activation / store of the current tile instead of heading the next one.
  uint32_t nxt[16];
  bool tpre = false;
  const bool xpre = p.xpre != 0 && !pair && !FOLD;
  while (cur.more) {
    const Loc nxtloc = next_item();
    const bool valid = cur.valid;
    const uint32_t lin = cur.lin;
    const int n0 = cur.n0;
    if (kSlots) {
      res_prefetch(nxtloc, rslot + (rs ^ 1u) * rstage);
      asm volatile("cp.async.wait_group 1;" ::: "memory");   // this tile's addend has landed
    }
    const uint8_t* rsm = smem_raw + ((rslot + rs * rstage) - smem_u32(smem_raw)) - (up_dup(lin) ? 16 : 0);
    const __nv_bfloat16* arow = nullptr;                     // direct-load addend row
    if (ADD == 3 && valid) arow = p.res + (size_t)lin * (uint32_t)p.rCtot + p.rC0 + n0;
    if (ADD == 4 && valid) arow = up_row(lin, n0);
    __nv_bfloat16* drow = nullptr;
    float* nrow = nullptr;
    if (NCHW) {
      if (valid) {
        const uint32_t nb = lin / (uint32_t)p.hw_real;        // image index, pixel inside the image
        nrow = p.nchw + (size_t)(nb * (uint32_t)p.nCtot + (uint32_t)(p.nC0 + n0)) * (uint32_t)p.hw_real + (lin - nb * (uint32_t)p.hw_real);
      }
    } else if (valid) {
      drow = p.dst + (size_t)lin * (uint32_t)p.dCtot + p.dC0 + n0;
    }
    if (cur.first && !tpre) { PROF_T0(); mbar_wait(bar_tfull(bar_base, as), aphase); PROF_ADD(w_tfull); }
    tc_fence_after();
    if (EXP_ON(p, 2)) {
      if (cur.last) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty(bar_base, as));
        if (++as == acc) { as = 0; aphase ^= 1u; }
      }
      rs ^= 1u;
      cur = nxtloc;
      continue;
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(((pair ? 2 * as : as) + cur.sub) * (FOLD ? 3 * p.block_n : p.block_n));
    if (cg < nchunks && !tpre) tmem_ld16(taddr + cg * 16, nxt);
    tpre = false;
    bool released = false;
    for (int rd = 0; rd < rounds; ++rd) {
      const int ch = rd * 4 + cg;
      const int c = ch * 16;
      const bool has = ch < nchunks;
      float v[16];
      if (has) {
        if (FOLD) {
          // out[m] = D[m][c..] + D[m+1][Cout + c..] + D[m+2][2*Cout + c..]: every thread loads the three column groups of
          // ITS row; rows m+1 / m+2 are the next lanes (shuffle), for lanes 30/31 the first lanes of the next quarter
          uint32_t tq[16];
          float* xs = exch + (((xpar * 4 + cg) * 4 + q) * 48);
          tmem_ld16(taddr + p.block_n + c, tq);          // D[row][Cout + c ..]: row m+1's share of pixel m
          tmem_ld_wait();
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) xs[j] = __uint_as_float(tq[j]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float a = __shfl_down_sync(0xFFFFFFFFu, __uint_as_float(tq[j]), 1);
            if (lane < 31) nxt[j] = __float_as_uint(__fadd_rn(__uint_as_float(nxt[j]), a));
          }
          tmem_ld16(taddr + 2 * p.block_n + c, tq);      // D[row][2*Cout + c ..]: row m+2's share
          tmem_ld_wait();
          if (lane < 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) xs[16 + 16 * lane + j] = __uint_as_float(tq[j]);
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float b2 = __shfl_down_sync(0xFFFFFFFFu, __uint_as_float(tq[j]), 2);
            if (lane < 30) nxt[j] = __float_as_uint(__fadd_rn(__uint_as_float(nxt[j]), b2));
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + cg) : "memory");
          if (lane >= 30) {     // rows m+1 / m+2 live in the next TMEM lane quarter (q = 3: rows 126, 127 are not outputs)
            const float* xn = exch + (((xpar * 4 + cg) * 4 + ((q + 1) & 3)) * 48);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = lane == 31 ? xn[j] : 0.f;                       // lane 30 got row m+1 by shuffle already
              const float b2 = lane == 31 ? xn[32 + j] : xn[16 + j];
              nxt[j] = __float_as_uint(__fadd_rn(__fadd_rn(__uint_as_float(nxt[j]), a), b2));
            }
          }
          xpar ^= 1u;
        } else {
          tmem_ld_wait();
        }
        const float4* bp = reinterpret_cast<const float4*>(s_bias + n0 + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 bb = bp[j];
          ffma2(v[4 * j + 0], v[4 * j + 1], __uint_as_float(nxt[4 * j + 0]), __uint_as_float(nxt[4 * j + 1]), pre, pre, bb.x, bb.y);
          ffma2(v[4 * j + 2], v[4 * j + 3], __uint_as_float(nxt[4 * j + 2]), __uint_as_float(nxt[4 * j + 3]), pre, pre, bb.z, bb.w);
        }
        if (ch + 4 < nchunks) tmem_ld16(taddr + c + 64, nxt);   // next round's chunk, in flight during the activation
      }
      if (!released && ch + 4 >= nchunks) {   // this warp's last TMEM load has completed (or it has none)
        released = true;
        if (cur.last) {                       // (pair mode: the accumulator set is free after its second tile)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty(bar_base, as));
        }
        if (xpre && has && nxtloc.more && cg < nchunks) {
          const int as2 = as + 1 == acc ? 0 : as + 1;
          const uint32_t ph2 = as2 == 0 ? aphase ^ 1u : aphase;
          if (mbar_test(bar_tfull(bar_base, as2), ph2)) {
            tc_fence_after();
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as2 * p.block_n) + cg * 16, nxt);
            tpre = true;
          }
        }
      }
      if (has) {
        if (kUp && valid) {
          float uv[16];
          if (kSlots) {
            load_vec<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(rsm + (2 * rd) * kPiece), uv);
            load_vec<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(rsm + (2 * rd + 1) * kPiece), uv + 8);
          } else {
            load_vec<__nv_bfloat16>(arow + c, uv);
            load_vec<__nv_bfloat16>(arow + c + 8, uv + 8);
          }
#pragma unroll
          for (int j = 0; j < 16; j += 2) ffma2(v[j], v[j + 1], uv[j], uv[j + 1], pre, pre, v[j], v[j + 1]);
        }
        if (act) {
#pragma unroll
          for (int j = 0; j < 16; j += 2) silu2_from_half(v[j], v[j + 1]);
        }
        if (kRes && valid) {
          float rv[16];
          if (kSlots) {
            load_vec<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(rsm + (2 * rd) * kPiece), rv);
            load_vec<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(rsm + (2 * rd + 1) * kPiece), rv + 8);
          } else {
            load_vec<__nv_bfloat16>(arow + c, rv);
            load_vec<__nv_bfloat16>(arow + c + 8, rv + 8);
          }
#pragma unroll
          for (int j = 0; j < 16; j += 2) fadd2(v[j], v[j + 1], v[j], v[j + 1], rv[j], rv[j + 1]);
        }
        if (NCHW) {
          if (nrow) {
            float* np = nrow + (size_t)c * (uint32_t)p.hw_real;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + c + j < p.nC) np[(size_t)j * (uint32_t)p.hw_real] = v[j];
          }
        } else if (drow) {
          if (p.st256) {
            store_bf16x16(drow + c, v);
          } else {
            store_vec<__nv_bfloat16>(drow + c, v);
            store_vec<__nv_bfloat16>(drow + c + 8, v + 8);
          }
        }
      }
    }
    if (cur.last && ++as == acc) { as = 0; aphase ^= 1u; }
    rs ^= 1u;
    cur = nxtloc;
  }
  if (kSlots) asm volatile("cp.async.wait_group 0;" ::: "memory");
#ifdef LY_TC_PROFILE
  if (blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 6))
    printf("[tc prof] epilogue warp %d: total %lld wait_tfull %lld\n", warp, clock64() - estart, w_tfull);
#endif
}

// ------------------------------------------------------------------ tile-parallel epilogue
// Knock-out runs (tools/exp_knockout.sh, -DLY_TC_EXP) and per-role cycle accounting (-DLY_TC_PROFILE) showed that the
// column-parallel role above costs ~1400-1600 cycles per 128-pixel tile whether the tile has 32 or 64 columns: with
// N <= 64 every warp handles ONE 16-column chunk per tile (N = 32: half of the warps have none), so the per-tile work
// of a warp (item arithmetic, barrier wait, arrival, pointers: ~130 instructions) is paid 16 times per tile for ~45
// instructions of real work each.  Here the 16 warps are G = 4 tile groups: a tile group owns every 4th tile of the
// issue order and its four warps (one per TMEM lane quarter) walk ALL the chunks of their rows, so the per-tile overhead
// is paid 4 times per tile and a thread's NHWC stores are consecutive 32-byte pieces of one pixel row.  Four tiles are in
// flight in the epilogue, which the TMEM stages (p.acc >= 4, i.e. N <= 128) make possible.
// The shortcut / half-resolution addend is read directly, one chunk ahead (the slots of the other role would need
// nchunks x 32 bytes per thread).
// TMA store of one staged [32 pixels x 32 channels] block (64-byte rows, SWIZZLE_64B).  ncu on the thin layers
// (profiles/r2_final_ncu_conv_tc_datapipe.txt): a `st.global.v8.b32` whose 32 lanes hit 32 different 128-byte lines costs ~49
// wavefronts of the l1tex LSU data stage (l1tex__data_pipe_lsu_wavefronts, peak one per cycle); the MMAs' shared-memory
// operands go through the TC side of the same stage (A: 32 + B: N/4 wavefronts per instruction, also one per cycle: the ~50
// cycle floor of an N <= 64 MMA).  The two counters overlap only partly (their busy fractions add up to 96-135 % on the
// thin layers), and the NHWC stores were 25-35 % of the LSU side in every conv kernel.  Staged through shared memory (4
// conflict-free STS.128 wavefronts per 512 bytes) and written by the TMA unit they cost about a third of that, and the
// tensor map's bounds clip the rows / channels that fall outside the tensor (no per-thread masks).
__device__ __forceinline__ void tma_store_4d(uint32_t src, const CUtensorMap* tm, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// 16 bf16 channels of the shortcut / addend row.  32 lanes x 32 bytes in 32 different lines: like the stores, one 256-bit
// request per lane costs about half the l1tex wavefronts of two 128-bit ones (ncu: the shortcut layers 3x3 64->64 @80^2
// run 29.0 M LSU wavefronts against 16.7 M without the shortcut and 23.9 M for the MMA operands).
#define LY_LD8(qual)                                                                                                   \
  asm volatile(qual " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                                                \
               : "=r"(a0.x), "=r"(a0.y), "=r"(a0.z), "=r"(a0.w), "=r"(a1.x), "=r"(a1.y), "=r"(a1.z), "=r"(a1.w) : "l"(ptr) : "memory")
__device__ __forceinline__ void ld_addend(const __nv_bfloat16* ptr, int wide, uint4& a0, uint4& a1) {
  if (wide == 1) { LY_LD8("ld.global.v8.b32"); }
  else if (wide == 2) { LY_LD8("ld.global.L2::64B.v8.b32"); }
  else if (wide == 3) { LY_LD8("ld.global.L1::no_allocate.v8.b32"); }
  else if (wide == 4) { LY_LD8("ld.global.L1::no_allocate.L2::64B.v8.b32"); }
  else {
    a0 = *reinterpret_cast<const uint4*>(ptr);
    a1 = *reinterpret_cast<const uint4*>(ptr + 8);
  }
}
#undef LY_LD8

template <int MAP, int ADD, bool NCHW>
__device__ __forceinline__ void epilogue_tiles(const Params& p, uint32_t bar_base, uint32_t tmem_base, const float* s_bias) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;
  constexpr int G = 4, CS = 1, cs = 0;     // (N > 128 keeps the column-parallel role: its per-tile overhead is amortised over 4 rounds)
  const int tg = (warp - 2) >> 2;          // tile group 0..3
  const int nchunks = p.block_n >> 4;
  // pair mode: tiles 2q / 2q+1 are the two halves of accumulator set q % acc: 2 * acc slots of block_n columns, one
  // tfull / tempty barrier per SET (both tiles complete together; 8 warps release it)
  const bool pair = MAP == 1 && p.pair != 0;
  const int acc = pair ? 2 * p.acc : p.acc;
  const uint32_t row = (uint32_t)(q * 32 + lane);
  const bool act = p.act != 0;
  const float pre = act ? 0.5f : 1.0f;
  constexpr bool kUp = ADD == 4;
  constexpr bool kRes = ADD == 3;
  uint32_t dw = 0, dh = 0, db = 0;
  if (MAP == 1) { dw = row % (uint32_t)p.tw; dh = (row / (uint32_t)p.tw) % (uint32_t)p.th; db = row / (uint32_t)(p.tw * p.th); }
  const int ts = (!NCHW && MAP != 2) ? p.ts : 0;                   // staging slots of this warp (0: direct stores)
  const uint32_t stg = bar_base + p.stg_off + (uint32_t)(warp - 2) * (uint32_t)ts * 2048u + (uint32_t)lane * 64u;
  const uint32_t sw = (uint32_t)((lane >> 1) & 3);                 // SWIZZLE_64B: 16-byte chunk index ^= (row / 2) % 4
  uint32_t stg_cnt = 0;

  // position in the issue order: (unit, n tile, M tile) in band mode, one tile per step otherwise
  int it_unit = blockIdx.x, it_nt = 0, it_mt = 0;
  auto advance = [&]() {
    if (MAP == 2) {
      if (++it_mt == p.band_mt) { it_mt = 0; if (++it_nt == p.tiles_n) { it_nt = 0; it_unit += gridDim.x; } }
    } else if (pair) {
      if (++it_mt == 2) { it_mt = 0; it_unit += gridDim.x; }
    } else {
      it_unit += gridDim.x;
    }
  };
  for (int i = 0; i < tg; ++i) advance();
  int as = tg;
  uint32_t aphase = 0;
  PROF_DECL(w_tfull);
#ifdef LY_TC_PROFILE
  const long long estart = clock64();
#endif
  while ((pair ? 2 * it_unit + it_mt : it_unit) < p.total_tiles) {
    // ---- locate this thread's pixel
    bool valid;
    uint32_t lin;
    int n0;
    int tc1 = 0, tc2 = 0, tc3 = 0;      // TMA-store coordinates of this lane's pixel (lane 0: the origin of the warp's box)
    if (MAP == 2) {
      const uint32_t pu = (uint32_t)phys(p, it_unit);
      const uint32_t b = p.mg_bands ? __umulhi(pu, p.mg_bands) : pu;
      const uint32_t bd = pu - b * (uint32_t)p.bands;
      const uint32_t m = (uint32_t)it_mt * 128u + row;
      const uint32_t oy = __umulhi(m, p.mg_bw), ox = m - oy * (uint32_t)p.band_w;
      const uint32_t h = bd * (uint32_t)p.band_r + oy;
      valid = ox < (uint32_t)p.Wo && oy < (uint32_t)p.band_r && h < (uint32_t)p.Ho;
      lin = (b * (uint32_t)p.Ho + h) * (uint32_t)p.Wo + ox;
      n0 = it_nt * p.block_n;
    } else if (MAP == 0) {
      const uint32_t t = (uint32_t)phys(p, it_unit);
      const uint32_t qn = p.mg_n ? __umulhi(t, p.mg_n) : t;
      n0 = (int)(t - qn * (uint32_t)p.tiles_n) * p.block_n;
      lin = qn * 128u + row;
      valid = lin < (uint32_t)p.Wo;
      tc1 = (int)lin;
    } else {
      int nt, wt, ht, bt;
      split_tile(p, pair ? 2 * it_unit + it_mt : it_unit, nt, wt, ht, bt);
      const uint32_t w = (uint32_t)(wt * p.tw) + dw, h = (uint32_t)(ht * p.th) + dh, b = (uint32_t)(bt * p.tb) + db;
      valid = w < (uint32_t)p.Wo && h < (uint32_t)p.Ho && b < (uint32_t)p.Bo;
      lin = (b * (uint32_t)p.Ho + h) * (uint32_t)p.Wo + w;
      n0 = nt * p.block_n;
      tc1 = (int)w; tc2 = (int)h; tc3 = (int)b;
    }
    const __nv_bfloat16* arow = nullptr;
    if (kRes && valid) arow = p.res + (size_t)lin * (uint32_t)p.rCtot + p.rC0 + n0;
    if (kUp && valid) {
      const uint32_t ub = lin / (uint32_t)p.hw_real, urem = lin - ub * (uint32_t)p.hw_real;
      const uint32_t uh = urem / (uint32_t)p.Wreal, uw = urem - uh * (uint32_t)p.Wreal;
      arow = p.up + (size_t)((ub * (uint32_t)p.uH + (uh >> 1)) * (uint32_t)p.uW + (uw >> 1)) * (uint32_t)p.uCtot + p.uC0 + n0;
    }
    __nv_bfloat16* drow = nullptr;
    float* nrow = nullptr;
    if (NCHW) {
      if (valid) {
        const uint32_t nb = lin / (uint32_t)p.hw_real;
        nrow = p.nchw + (size_t)(nb * (uint32_t)p.nCtot + (uint32_t)(p.nC0 + n0)) * (uint32_t)p.hw_real + (lin - nb * (uint32_t)p.hw_real);
      }
    } else if (valid) {
      drow = p.dst + (size_t)lin * (uint32_t)p.dCtot + p.dC0 + n0;
    }
    // the addend of the first TWO chunks is requested before the accumulator wait, later chunks two chunks ahead: one
    // chunk of lead (~150 cycles of work) does not cover the DRAM latency of these scattered 32-byte reads
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0, b0 = a0, b1 = a0;
    if ((kRes || kUp) && arow && cs < nchunks) ld_addend(arow + cs * 16, p.ld256, a0, a1);
    if ((kRes || kUp) && arow && cs + CS < nchunks) ld_addend(arow + (cs + CS) * 16, p.ld256, b0, b1);
    const int bs = pair ? as >> 1 : as;          // barrier index of this accumulator
    { PROF_T0(); mbar_wait(bar_tfull(bar_base, bs), aphase); PROF_ADD(w_tfull); }
    tc_fence_after();
    if (EXP_ON(p, 2)) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty(bar_base, bs));
    } else {
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.block_n);
    uint32_t nxt[16];
    if (cs < nchunks) tmem_ld16(taddr + cs * 16, nxt);
    else {   // (a column group without chunks: N < 16 * CS)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty(bar_base, bs));
    }
    for (int ch = cs; ch < nchunks; ch += CS) {
      const int c = ch * 16;
      float v[16];
      tmem_ld_wait();
      const float4* bp = reinterpret_cast<const float4*>(s_bias + n0 + c);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 bb = bp[j];
        ffma2(v[4 * j + 0], v[4 * j + 1], __uint_as_float(nxt[4 * j + 0]), __uint_as_float(nxt[4 * j + 1]), pre, pre, bb.x, bb.y);
        ffma2(v[4 * j + 2], v[4 * j + 3], __uint_as_float(nxt[4 * j + 2]), __uint_as_float(nxt[4 * j + 3]), pre, pre, bb.z, bb.w);
      }
      const uint4 r0 = a0, r1 = a1;
      if (kRes || kUp) { a0 = b0; a1 = b1; }
      if (ch + CS < nchunks) {
        tmem_ld16(taddr + c + 16 * CS, nxt);   // next chunk in flight during the activation / store of this one
        if ((kRes || kUp) && arow && ch + 2 * CS < nchunks) ld_addend(arow + c + 32 * CS, p.ld256, b0, b1);
      } else {                                  // the last chunk of this warp sits in registers: release the stage
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty(bar_base, bs));
      }
      if (kRes || kUp) {
        float rv[16];
        const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
        const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f0 = __bfloat1622float2(h0[j]), f1 = __bfloat1622float2(h1[j]);
          rv[2 * j] = f0.x; rv[2 * j + 1] = f0.y; rv[8 + 2 * j] = f1.x; rv[8 + 2 * j + 1] = f1.y;
        }
        if (kUp) {    // pre-activation addend (the low-resolution half of a folded upsample + concat + 1x1)
#pragma unroll
          for (int j = 0; j < 16; j += 2) ffma2(v[j], v[j + 1], rv[j], rv[j + 1], pre, pre, v[j], v[j + 1]);
          if (act) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) silu2_from_half(v[j], v[j + 1]);
          }
        } else {      // shortcut: added after the activation
          if (act) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) silu2_from_half(v[j], v[j + 1]);
          }
#pragma unroll
          for (int j = 0; j < 16; j += 2) fadd2(v[j], v[j + 1], v[j], v[j + 1], rv[j], rv[j + 1]);
        }
      } else if (act) {
#pragma unroll
        for (int j = 0; j < 16; j += 2) silu2_from_half(v[j], v[j + 1]);
      }
      if (NCHW) {
        if (nrow) {
          float* np = nrow + (size_t)c * (uint32_t)p.hw_real;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (n0 + c + j < p.nC) np[(size_t)j * (uint32_t)p.hw_real] = v[j];
        }
      } else if (ts) {
        if ((ch & 1) == 0) {               // first chunk of a 32-channel group: the slot's previous store must have been read
          if (lane == 0) {
            if (ts > 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          __syncwarp();
        }
        const uint32_t slot = stg + (ts > 1 ? (stg_cnt & 1u) * 2048u : 0u);
        uint32_t w8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
          w8[j] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        const uint32_t k0 = (uint32_t)(ch & 1) * 2u;
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(slot + ((k0 ^ sw) << 4)), "r"(w8[0]), "r"(w8[1]), "r"(w8[2]), "r"(w8[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(slot + (((k0 + 1u) ^ sw) << 4)), "r"(w8[4]), "r"(w8[5]), "r"(w8[6]), "r"(w8[7]) : "memory");
        if ((ch & 1) != 0 || ch + 1 >= nchunks) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) tma_store_4d(slot, &p.tmD, n0 + (ch & ~1) * 16, tc1, tc2, tc3);
          ++stg_cnt;
        }
      } else if (drow) {
        if (p.st256) {
          store_bf16x16(drow + c, v);
        } else {
          store_vec<__nv_bfloat16>(drow + c, v);
          store_vec<__nv_bfloat16>(drow + c + 8, v + 8);
        }
      }
    }
    }
    for (int i = 0; i < G; ++i) advance();
    as += G;
    if (as >= acc) { as -= acc; aphase ^= 1u; }
  }
  if (ts && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all of this warp's stores have been written
#ifdef LY_TC_PROFILE
  if (blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 6))
    printf("[tc prof] tile epilogue warp %d: total %lld wait_tfull %lld\n", warp, clock64() - estart, w_tfull);
#endif
}

// ------------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A ring][B ring or resident B][barriers]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + (uint32_t)p.a_stages * p.a_stage;
  const uint32_t b_bytes_total = p.fold ? (uint32_t)p.num_kb * p.b_box
                                        : (p.b_resident ? (uint32_t)p.num_kb * p.b_stage : (uint32_t)p.b_stages * p.b_stage);
  const uint32_t bar_base = b_base + b_bytes_total;
  auto afull_bar = [&](int s) { return bar_base + 8u * s; };
  auto aempty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto bfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto bempty_bar = [&](int s) { return bar_base + 8u * (3 * kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (4 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (4 * kMaxStages + kMaxAcc + s); };
  const uint32_t bres_bar = bar_base + 8u * (4 * kMaxStages + 2 * kMaxAcc);
  const uint32_t tmem_slot = bar_base + 8u * (4 * kMaxStages + 2 * kMaxAcc + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();   // the next kernel's CTAs may be scheduled as SMs drain (they block in their own pdl_wait)
  __shared__ __align__(16) float s_bias[kMaxCout];
  for (int i = threadIdx.x; i < p.tiles_n * p.block_n; i += kThreads) s_bias[i] = p.act ? 0.5f * p.bias[i] : p.bias[i];

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmB) : "memory");
    if (p.ts) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmD) : "memory");
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(afull_bar(s), 1);
      mbar_init(aempty_bar(s), 1);
      mbar_init(bfull_bar(s), 1);
      mbar_init(bempty_bar(s), 1);
    }
    for (int s = 0; s < kMaxAcc; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), p.epi_groups ? (kEpiWarps / p.epi_groups) * (p.pair ? 2 : 1) : kEpiWarps);
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // everything above touched only parameters (weights' bias, tensor maps) and on-chip state;
  // activations written by the previous kernel are read from here on
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      // k-block order (shared with the MMA issuer, which simply counts k-blocks):
      //   classic: for tap (ky, kx): for cb            halo: for cb: for kx: [A box], ky = 0..2
      const int KK = p.k, kcb = p.kc_blocks, kc = p.kc, cin = p.cin_pad;
      if (p.fold) {
        // slab (ky, cb) = three [Cout x kc] boxes (kx = 0, 1, 2) stacked along N
        mbar_expect_tx(bres_bar, (uint32_t)p.num_kb * p.b_box);
        const uint32_t sub = (uint32_t)p.b_box;
        for (int ky = 0; ky < 3; ++ky)
          for (int cb = 0; cb < kcb; ++cb)
            for (int kx = 0; kx < 3; ++kx)
              tma_load_2d(b_base + (uint32_t)(ky * kcb + cb) * p.b_stage + (uint32_t)kx * sub, &p.tmB, bres_bar, (ky * 3 + kx) * cin + cb * kc, 0);
      } else if (p.b_resident) {
        mbar_expect_tx(bres_bar, (uint32_t)p.num_kb * p.b_box);
        uint32_t dstb = b_base;
        if (p.halo == 1) {
          for (int cb = 0; cb < kcb; ++cb)
            for (int kx = 0; kx < 3; ++kx)
              for (int ky = 0; ky < 3; ++ky, dstb += p.b_stage)
                tma_load_2d(dstb, &p.tmB, bres_bar, (ky * 3 + kx) * cin + cb * kc, 0);
        } else {
          for (int tap = 0; tap < KK * KK; ++tap)
            for (int cb = 0; cb < kcb; ++cb, dstb += p.b_stage) tma_load_2d(dstb, &p.tmB, bres_bar, tap * cin + cb * kc, 0);
        }
      }
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      PROF_DECL(w_aempty); PROF_DECL(w_bempty); PROF_DECL(p_total);
#ifdef LY_TC_PROFILE
      const long long pstart = clock64();
#endif
      const bool bres = p.b_resident != 0;
      const uint32_t a_box = (uint32_t)p.a_box, b_box = (uint32_t)p.b_box;
      int exp_loaded = 0; (void)exp_loaded;
      auto load_a = [&](int c0, int w, int h, int b) {
        { PROF_T0(); mbar_wait(aempty_bar(sa), pa ^ 1u); PROF_ADD(w_aempty); }
        if (EXP_ON(p, 4) && exp_loaded >= p.a_stages) {
          mbar_arrive(afull_bar(sa));
        } else {
          ++exp_loaded;
          mbar_expect_tx(afull_bar(sa), a_box);
          tma_load_4d(a_base + sa * p.a_stage, &p.tmA, afull_bar(sa), c0, w, h, b);
        }
        if (++sa == p.a_stages) { sa = 0; pa ^= 1u; }
      };
      auto load_b = [&](int kcol, int n0) {
        { PROF_T0(); mbar_wait(bempty_bar(sb), pb ^ 1u); PROF_ADD(w_bempty); }
        mbar_expect_tx(bfull_bar(sb), b_box);
        tma_load_2d(b_base + sb * p.b_stage, &p.tmB, bfull_bar(sb), kcol, n0);
        if (++sb == p.b_stages) { sb = 0; pb ^= 1u; }
      };
      if (p.halo == 2) {
        // band mode: unit = (band, image); one box per channel block, then (if the weights stream)
        // the weight slabs in the issuer's order.  The A boxes of the NEXT unit are requested before
        // this unit's weight slabs so that the band prefetch runs a whole unit ahead.
        auto unit_a = [&](int unit) {
          unit = phys(p, unit);
          const uint32_t b = p.mg_bands ? __umulhi((uint32_t)unit, p.mg_bands) : (uint32_t)unit;
          const int band = unit - (int)b * p.bands;
          for (int cb = 0; cb < kcb; ++cb) load_a(cb * kc, -1, band * p.band_r - 1, (int)b);
        };
        if ((int)blockIdx.x < p.total_tiles) unit_a(blockIdx.x);
        for (int unit = blockIdx.x; unit < p.total_tiles; unit += gridDim.x) {
          if (unit + (int)gridDim.x < p.total_tiles) unit_a(unit + gridDim.x);
          if (!bres)
            for (int nt = 0; nt < p.tiles_n; ++nt)
              for (int mt = 0; mt < p.band_mt; ++mt)
                for (int cb = 0; cb < kcb; ++cb)
                  for (int tap = 0; tap < 9; ++tap) load_b(tap * cin + cb * kc, nt * p.block_n);
        }
      } else if (p.pair) {
        // pair mode: the A boxes of tiles 2q and 2q+1 alternate in the ring, each weight slab is requested once
        const int pairs = (p.total_tiles + 1) >> 1;
        for (int q = blockIdx.x; q < pairs; q += gridDim.x) {
          const bool two = 2 * q + 1 < p.total_tiles;
          int nt, wt, ht, bt, nt1, wt1 = 0, ht1 = 0, bt1 = 0;
          split_tile(p, 2 * q, nt, wt, ht, bt);
          if (two) split_tile(p, 2 * q + 1, nt1, wt1, ht1, bt1);
          const int w0 = wt * p.tw * p.stride - p.pad, h0 = ht * p.th * p.stride - p.pad, b0 = bt * p.tb;
          const int w1 = wt1 * p.tw * p.stride - p.pad, h1 = ht1 * p.th * p.stride - p.pad, b1 = bt1 * p.tb;
          if (p.halo) {
            for (int cb = 0; cb < kcb; ++cb)
              for (int kx = 0; kx < 3; ++kx) {
                load_a(cb * kc, w0 + kx, h0, b0);
                if (two) load_a(cb * kc, w1 + kx, h1, b1);
                for (int ky = 0; ky < 3; ++ky) load_b((ky * 3 + kx) * cin + cb * kc, 0);
              }
          } else {
            for (int ky = 0; ky < KK; ++ky)
              for (int kx = 0; kx < KK; ++kx)
                for (int cb = 0; cb < kcb; ++cb) {
                  load_a(cb * kc, w0 + kx, h0 + ky, b0);
                  if (two) load_a(cb * kc, w1 + kx, h1 + ky, b1);
                  load_b((ky * KK + kx) * cin + cb * kc, 0);
                }
          }
        }
      } else
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int nt, wt, ht, bt;
        split_tile(p, tile, nt, wt, ht, bt);
        const int w0 = wt * p.tw * p.stride - p.pad, h0 = ht * p.th * p.stride - p.pad, b0 = bt * p.tb;
        const int n0 = nt * p.block_n;
        if (p.halo) {
          // the box starts one row above the brick and is th+2 rows tall; ky selects a window of it
          for (int cb = 0; cb < kcb; ++cb)
            for (int kx = 0; kx < 3; ++kx) {
              load_a(cb * kc, w0 + kx, h0, b0);
              if (!bres)
                for (int ky = 0; ky < 3; ++ky) load_b((ky * 3 + kx) * cin + cb * kc, n0);
            }
        } else {
          for (int ky = 0; ky < KK; ++ky)
            for (int kx = 0; kx < KK; ++kx)
              for (int cb = 0; cb < kcb; ++cb) {
                load_a(cb * kc, w0 + kx, h0 + ky, b0);
                if (!bres) load_b((ky * KK + kx) * cin + cb * kc, n0);
              }
        }
      }
#ifdef LY_TC_PROFILE
      p_total = clock64() - pstart;
      if (blockIdx.x == 0) printf("[tc prof] producer: total %lld wait_aempty %lld wait_bempty %lld\n", p_total, w_aempty, w_bempty);
#endif
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ================================
    if (elect_one()) {
      const int ksteps = p.kc / 16;
      if (p.fold) {
        if (ksteps == 4) mma_role_fold<4>(p, a_base, b_base, bar_base, tmem_base);
        else if (ksteps == 2) mma_role_fold<2>(p, a_base, b_base, bar_base, tmem_base);
        else mma_role_fold<1>(p, a_base, b_base, bar_base, tmem_base);
      } else if (p.halo == 2) {
        if (ksteps == 4) { if (p.b_resident) mma_role_band<4, true>(p, a_base, b_base, bar_base, tmem_base); else mma_role_band<4, false>(p, a_base, b_base, bar_base, tmem_base); }
        else if (ksteps == 2) { if (p.b_resident) mma_role_band<2, true>(p, a_base, b_base, bar_base, tmem_base); else mma_role_band<2, false>(p, a_base, b_base, bar_base, tmem_base); }
        else { if (p.b_resident) mma_role_band<1, true>(p, a_base, b_base, bar_base, tmem_base); else mma_role_band<1, false>(p, a_base, b_base, bar_base, tmem_base); }
      } else if (p.pair) {
        if (ksteps == 4) { if (p.tpa == 3) mma_role_pair<4, 3>(p, a_base, b_base, bar_base, tmem_base); else mma_role_pair<4, 1>(p, a_base, b_base, bar_base, tmem_base); }
        else if (ksteps == 2) { if (p.tpa == 3) mma_role_pair<2, 3>(p, a_base, b_base, bar_base, tmem_base); else mma_role_pair<2, 1>(p, a_base, b_base, bar_base, tmem_base); }
        else { if (p.tpa == 3) mma_role_pair<1, 3>(p, a_base, b_base, bar_base, tmem_base); else mma_role_pair<1, 1>(p, a_base, b_base, bar_base, tmem_base); }
      } else
#define LY_MMA_CASE(KS)                                                                       \
  if (ksteps == KS) {                                                                         \
    if (p.tpa == 3) { if (p.b_resident) mma_role<KS, 3, true>(p, a_base, b_base, bar_base, tmem_base);   \
                      else mma_role<KS, 3, false>(p, a_base, b_base, bar_base, tmem_base); }  \
    else            { if (p.b_resident) mma_role<KS, 1, true>(p, a_base, b_base, bar_base, tmem_base);   \
                      else mma_role<KS, 1, false>(p, a_base, b_base, bar_base, tmem_base); }  \
  }
      { LY_MMA_CASE(4) else LY_MMA_CASE(2) else LY_MMA_CASE(1) }
#undef LY_MMA_CASE
    }
  } else {
    // ============================== epilogue (16 warps) =======================
    const int add = p.res ? (p.res_slot ? 1 : 3) : (p.up ? (p.res_slot ? 2 : 4) : 0);
    const int map = p.halo == 2 ? 2 : ((p.k == 1 && p.stride == 1) ? 0 : 1);
#define LY_EPI(M, A, N) epilogue_role<M, A, N>(p, smem_raw, bar_base, tmem_base, s_bias)
#define LY_EPI_MAP(M)                                           \
  if (p.nchw) LY_EPI(M, 0, true);                               \
  else if (add == 0) LY_EPI(M, 0, false);                       \
  else if (add == 1) LY_EPI(M, 1, false);                       \
  else if (add == 2) LY_EPI(M, 2, false);                       \
  else if (add == 3) LY_EPI(M, 3, false);                       \
  else LY_EPI(M, 4, false);
#define LY_EPT(M, A, N) epilogue_tiles<M, A, N>(p, bar_base, tmem_base, s_bias)
#define LY_EPT_MAP(M)                                           \
  if (p.nchw) LY_EPT(M, 0, true);                               \
  else if (add == 0) LY_EPT(M, 0, false);                       \
  else if (add == 3) LY_EPT(M, 3, false);                       \
  else LY_EPT(M, 4, false);
    if (p.epi_groups) {
      if (map == 0) { LY_EPT_MAP(0) } else if (map == 1) { LY_EPT_MAP(1) } else { LY_EPT_MAP(2) }
    } else
#undef LY_EPT_MAP
    if (p.fold) {
      if (add == 0) epilogue_role<2, 0, false, true>(p, smem_raw, bar_base, tmem_base, s_bias);
      else if (add == 1) epilogue_role<2, 1, false, true>(p, smem_raw, bar_base, tmem_base, s_bias);
      else epilogue_role<2, 3, false, true>(p, smem_raw, bar_base, tmem_base, s_bias);
    } else if (map == 0) { LY_EPI_MAP(0) } else if (map == 1) { LY_EPI_MAP(1) } else { LY_EPI_MAP(2) }
#undef LY_EPI_MAP
#undef LY_EPI
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------- host side
int pow2_ge(int v) { int p = 32; while (p < v) p <<= 1; return p; }

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

}  // namespace

struct ConvTcState {
  Params p;
  int grid;
  size_t smem;
};

bool conv_tc_supported(const ly_op& op) {
  if (op.dtype != LY_BF16 || op.kind != LY_OP_CONV) return false;
  if (!(op.k == 1 || op.k == 3) || !(op.stride == 1 || op.stride == 2)) return false;
  if (op.src.c % 16 || op.src.c0 % 8 || op.src.ctot % 8) return false;
  if (op.dst.ptr && (op.dst.c % 16 || op.dst.c0 % 8 || op.dst.ctot % 8)) return false;
  if (op.res.ptr && (op.res.c0 % 8 || op.res.ctot % 8)) return false;
  if (op.up.ptr && (op.up.c0 % 8 || op.up.ctot % 8 || op.up.H * 2 != op.src.H / op.stride || op.up.W * 2 != op.src.W / op.stride)) return false;
  return true;
}

int32_t conv_tc_prepare(const ly_op& op, ConvTcState** out) {
  LY_CHECK_ARG(conv_tc_supported(op), "conv_tc: unsupported op (bf16, k in {1,3}, stride in {1,2}, 16-channel granularity)");
  LY_CHECK_ARG(op.src.ptr && op.w && op.bias && (op.dst.ptr || op.nchw), "conv_tc: null pointer");
  LY_CHECK_ARG(!(op.nchw && (op.dst.ptr || op.res.ptr || op.up.ptr)), "conv_tc: an NCHW output takes no NHWC destination, shortcut or addend");
  LY_CHECK_ARG(op.src.H % op.stride == 0 && op.src.W % op.stride == 0, "conv_tc: H,W must divide by the stride");
  LY_CHECK_ARG((op.dst.ptr ? op.dst.c : op.nchw_c) <= kMaxCout, "conv_tc: Cout > %d not supported", kMaxCout);
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("conv_tc: cuTensorMapEncodeTiled entry point not available"); return LY_E_CUDA; }

  ConvTcState* st = new ConvTcState();
  Params& p = st->p;
  memset(&p, 0, sizeof(p));
  const int Cin = op.src.c;
  const int Cout = op.dst.ptr ? op.dst.c : (op.nchw_c + 15) / 16 * 16;
  const int Ho = op.src.H / op.stride, Wo = op.src.W / op.stride;
  p.k = op.k; p.stride = op.stride; p.pad = op.k / 2; p.act = op.act; p.cin_pad = Cin;
  p.rev = g_reverse;
  p.hw_real = Ho * Wo;
  p.kc = Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 16);
  p.kc_blocks = Cin / p.kc;
  p.num_kb = op.k * op.k * p.kc_blocks;
  // N tile: the largest multiple of 16 that divides Cout and is <= 256
  // (ops with a prefetched addend keep N <= 128 so that the per-thread prefetch slots stay small)
  const int bn_max = op.up.ptr ? 128 : 256;
  int bn = 16;
  for (int c = 16; c <= bn_max && c <= Cout; c += 16)
    if (Cout % c == 0) bn = c;
  p.block_n = bn;
  p.tiles_n = Cout / bn;
  p.tmem_cols = pow2_ge(2 * bn);

  // pixel brick
  const bool flat = (op.k == 1 && op.stride == 1);
  int dimW, dimH, dimB;
  if (flat) { dimW = op.B * Ho * Wo; dimH = 1; dimB = 1; } else { dimW = Wo; dimH = Ho; dimB = op.B; }
  int best_tw = 128, best_th = 1, best_tb = 1;
  double best_cover = 1e30;
  if (!flat) {   // flat: 128 consecutive pixels per tile (the epilogue's MAP = 0 assumes it)
    double best_cost = 1e30;
    for (int tw = 128; tw >= 1; tw >>= 1)
      for (int th = 128 / tw; th >= 1; th >>= 1) {
        const int tb = 128 / (tw * th);
        if (tw * op.stride > 256 || th * op.stride > 256 || tb > 256) continue;
        const double cover = (double)((dimW + tw - 1) / tw * tw) * ((dimH + th - 1) / th * th) * ((dimB + tb - 1) / tb * tb);
        // prefer wide bricks (longer contiguous runs) on ties, penalise batch-spanning bricks slightly
        const double cost = cover * (1.0 + 1e-3 * (tb > 1) + 1e-4 * (128 / tw));
        if (cost < best_cost) { best_cost = cost; best_cover = cover; best_tw = tw; best_th = th; best_tb = tb; }
      }
  }
  // halo mode: 3x3 stride 1 with the 8 x 16 brick, unless that brick wastes > 35 % more pixels
  static const int halo_ok = env_int("LY_TC_HALO", 1);
  static const int halo2_ok = env_int("LY_TC_HALO_S2", 1);
  // (stride 2: only for 64-byte channel rows, where the TMA request rate binds: 32->64 @320^2 0.65 -> 0.51 ms;
  //  with 128-byte rows the classic 9-load mode is as fast or faster)
  if (halo_ok && op.k == 3 && (op.stride == 1 || (halo2_ok && p.kc <= 32))) {
    const double cover = (double)((dimW + 7) / 8 * 8) * ((dimH + 15) / 16 * 16) * dimB;
    if (cover <= 1.35 * best_cover) { p.halo = 1; best_tw = 8; best_th = 16; best_tb = 1; }
  }
  // experiment (LY_TC_HALO_TW=4): 4-wide brick, so the ky windows start 4 rows = HALF a swizzle atom apart
  static const int halo_tw = env_int("LY_TC_HALO_TW", 8);
  if (p.halo && halo_tw != 8 && op.stride == 1) { best_tw = halo_tw; best_th = 128 / halo_tw; }
  p.tw = best_tw; p.th = best_th; p.tb = best_tb;
  p.Wo = dimW; p.Ho = dimH; p.Bo = dimB;
  p.tiles_w = (dimW + p.tw - 1) / p.tw;
  p.tiles_h = (dimH + p.th - 1) / p.th;
  p.tiles_b = (dimB + p.tb - 1) / p.tb;
  long long total = (long long)p.tiles_w * p.tiles_h * p.tiles_b * p.tiles_n;

  // ---- band mode (3x3 stride 1): pick the band height R by a cycle model of one CTA:
  //   MMA issue  : (32 + N/4) cycles per M=128,K=16 instruction (shared-memory operand reads; N = 256: 128)
  //   TMA        : ~5 cycles per box row (measured request rate on 64/128-byte rows)
  // and compare with the brick/halo tiling chosen above under the same model.
  static const int band_ok = env_int("LY_TC_BAND", 1);   // 0 off, 1 by the model, 2 whenever it fits
  const long long b_all_bytes = (long long)p.num_kb * ((bn * p.kc * 2 + 1023) / 1024 * 1024);
  const bool b_res_possible = env_int("LY_TC_B_RESIDENT", 1) && p.tiles_n == 1 && b_all_bytes <= 96 * 1024;
  // pair mode: streamed weights shared by two pixel tiles (see mma_role_pair)
  static const int pair_ok = env_int("LY_TC_PAIR", 1);
  p.pair = (pair_ok && op.k == 3 && p.tiles_n == 1 && !b_res_possible && 4 * bn <= 512) ? 1 : 0;
  if (p.pair) p.tmem_cols = pow2_ge(4 * bn);
  // fold: the three kx taps in the MMA's N dimension (see the header comment); resident weights, NHWC output only
  // Opt-in (LY_TC_FOLD=1).  Measured on 3x3 64->64 @80^2, batch 256: the issuer's wait share drops from 38 % to 6 % of the
  // epilogue warps' time (12 MMAs of N = 192 instead of 36 of N = 64), but the epilogue, which was already co-critical
  // (~1600 cycles per tile and warp), grows to ~3000 (two more TMEM round trips, 32 shuffles, the quarter-boundary
  // exchange + named barrier): 0.158 -> 0.185 ms, with a shortcut 0.166 -> 0.291 ms; 32->32 @160^2 0.311 -> 0.479 ms.
  // Round 2, second attempt (commit "fold with a tile-parallel shuffle epilogue", removed again): one exchange barrier per tile
  // instead of one per chunk, rotating shuffles, all chunks of a row per warp: correct, 0.143 -> 0.170 ms.  ncu on that run: the
  // l1tex data stage moves one 128-byte wavefront per cycle for the MMAs' operand reads (A: 32 + B: N/4 wavefronts per
  // instruction: THAT is the ~50-cycle floor of an N <= 64 MMA) and one per cycle for the LSU side (shuffles, LDS / STS, global
  // stores); the fold removes 768 operand wavefronts per tile and adds 512 shuffles + ~300 exchange LDS/STS on the other side:
  // LSU 65.6 % + TC 30.4 % busy, tensor pipe 36 %, and the epilogue warps, not the issuer, set the pace.
  // tcgen05.shift (tools/tmem_probe.cu: one instruction shifts 8 columns by one lane inside each 32-lane quarter, ~24 cycles
  // each) would cost 24 shifts = as much tensor-pipe time as the fold saves, and cannot cross the quarter boundary either.
  static const int fold_env = env_int("LY_TC_FOLD", 0);
  const bool fold_ok = fold_env && op.k == 3 && op.stride == 1 && p.tiles_n == 1 && 3 * bn <= 256 && (bn * p.kc * 2) % 1024 == 0 &&
                       b_res_possible && !p.pair && !op.up.ptr && !op.nchw && op.dst.ptr;
  const int m_step = fold_ok ? 126 : 128;
  const long long exch_bytes = fold_ok ? 2LL * 4 * 4 * 48 * 4 : 0;
  // (the tile-parallel epilogue reads the shortcut directly: no prefetch slots)
  const bool will_ept = env_int("LY_TC_EPT", 1) && !p.pair && !fold_ok && std::min(512 / bn, env_int("LY_TC_ACC", kMaxAcc)) >= 4;
  if (band_ok && !p.pair && op.k == 3 && op.stride == 1 && Wo + 2 <= 256 && 2 * p.kc_blocks <= kMaxStages) {
    const int sms = sm_count();
    const int nf = 3 * bn;
    const double cyc_mma = fold_ok ? (nf <= 128 ? 32.0 + nf / 4.0 : nf / 2.0) * (p.kc / 16) / 3.0     // per k-block (tap) per M tile
                                   : (bn <= 128 ? 32.0 + bn / 4.0 : bn / 2.0) * (p.kc / 16);
    const double tma_row = 5.0;
    auto waves = [&](long long units) { return (double)((units + sms - 1) / sms); };
    const double cost_now = waves(total) * (p.halo ? (double)std::max(p.num_kb * cyc_mma, 3.0 * p.kc_blocks * (p.tw * (p.th + 2)) * tma_row)
                                                   : (double)std::max(p.num_kb * cyc_mma, 9.0 * p.kc_blocks * 128 * tma_row));
    const int BW = Wo + 2;
    const long long fixed = 8 * kBarSlots + 1024 + exch_bytes + (op.res.ptr && bn <= 128 && !will_ept ? 2LL * 32 * kEpiWarps * (((bn / 16 + 3) / 4) * 32) : 0);
    const long long b_bytes = b_res_possible ? b_all_bytes : 3LL * ((bn * p.kc * 2 + 1023) / 1024 * 1024);
    int best_r = 0; double best_cost = 1e30;
    for (int R = 1; R <= Ho && R + 2 <= 256; ++R) {
      const long long stage = ((long long)(R + 2) * BW * p.kc * 2 + 1023) / 1024 * 1024;
      if (fixed + b_bytes + 2LL * p.kc_blocks * stage > (long long)kSmemBudget) break;
      const int mt = ((R - 1) * BW + Wo + m_step - 1) / m_step;
      const long long units = (long long)((Ho + R - 1) / R) * op.B;
      const double unit_cost = std::max((double)p.tiles_n * mt * p.num_kb * cyc_mma, (double)p.kc_blocks * (R + 2) * BW * tma_row);
      const double cost = waves(units) * unit_cost;
      if (cost < best_cost) { best_cost = cost; best_r = R; }
    }
    if (best_r && (band_ok == 2 || best_cost < cost_now)) {
      p.halo = 2;
      p.band_r = best_r; p.band_w = BW;
      p.band_mt = ((best_r - 1) * BW + Wo + m_step - 1) / m_step;
      p.bands = (Ho + best_r - 1) / best_r;
      if (fold_ok) {
        p.fold = 1;
        p.tmem_cols = pow2_ge(2 * 3 * bn);
      }
      p.tw = 128; p.th = 1; p.tb = 1;
      total = (long long)p.bands * op.B;
    }
  }
  if (total > 0x7FFFFFFF) { delete st; set_error("conv_tc: too many tiles"); return LY_E_ARG; }
  p.total_tiles = (int)total;
  {
    // x / d == __umulhi(x, floor(2^32 / d) + 1) whenever x * d < 2^32; 0 encodes d == 1
    auto magic = [](uint32_t d) -> uint32_t { return d <= 1 ? 0u : (uint32_t)((1ull << 32) / d + 1); };
    p.mg_n = magic((uint32_t)p.tiles_n); p.mg_w = magic((uint32_t)p.tiles_w); p.mg_h = magic((uint32_t)p.tiles_h);
    const unsigned long long lim = 1ull << 32, tmax = (unsigned long long)total + 2ull * sm_count();
    if (tmax * p.tiles_n >= lim || tmax * p.tiles_w >= lim || tmax * p.tiles_h >= lim || (unsigned long long)dimW * dimH * dimB >= lim) {
      delete st; set_error("conv_tc: problem too large for 32-bit tile arithmetic"); return LY_E_ARG;
    }
    if (p.halo == 2) {
      // (a divisor of 1 cannot use the 0 encoding here: the kernel always multiplies)
      p.mg_bw = (uint32_t)((1ull << 32) / (uint32_t)p.band_w + 1);
      p.mg_bands = p.bands > 1 ? (uint32_t)((1ull << 32) / (uint32_t)p.bands + 1) : 0u;
      if (tmax * p.bands >= lim) { delete st; set_error("conv_tc: problem too large for 32-bit tile arithmetic"); return LY_E_ARG; }
    }
  }
  p.tpa = p.halo == 1 ? 3 : 1;
  p.num_ka = p.num_kb / p.tpa;
  // halo box: 8 output columns x all the input rows the 16 output rows touch (18 at stride 1, 33 at stride 2)
  const int halo_rows = (p.th - 1) * op.stride + 3;
  const int a_rows = p.halo == 2 ? (p.band_r + 2) * p.band_w : (p.halo ? p.tw * halo_rows : 128);
  p.a_tap_stride = p.tw * p.kc * 2;        // one brick row down (tw = 8: one swizzle atom)

  // shared-memory pipeline
  p.a_box = a_rows * p.kc * 2;
  p.a_stage = (p.a_box + 1023) / 1024 * 1024;
  p.b_box = bn * p.kc * 2;
  p.b_stage = (p.b_box + 1023) / 1024 * 1024;
  if (p.fold) p.b_stage = 3 * p.b_box;       // one slab = the three kx boxes stacked along N (b_box is a multiple of 1024 here)
  const long long b_all = p.fold ? (long long)p.num_kb * p.b_box : (long long)p.num_kb * p.b_stage;
  static const int resident_ok = env_int("LY_TC_B_RESIDENT", 1);
  p.b_resident = (resident_ok && p.tiles_n == 1 && b_all <= 96 * 1024) ? 1 : 0;
  const uint32_t bar_bytes = 8 * kBarSlots;
  // shortcut prefetch slots: 2 stages x 512 epilogue threads x (chunks per warp x 32 B); only while small
  // tile-parallel epilogue (see epilogue_tiles): whenever the TMEM stages allow >= 2 tiles in flight
  static const int ept_env = env_int("LY_TC_EPT", 1);
  {
    const int per = bn;   // (pair and fold keep the column-parallel role)
    const int acc_max = std::min(512 / per, std::min(env_int("LY_TC_ACC", kMaxAcc), (int)kMaxAcc));
    p.epi_groups = (ept_env && !p.fold && (p.pair ? env_int("LY_TC_EPT_PAIR", 1) != 0 : acc_max >= 4)) ? 4 : 0;   // (pair: 2 sets x 2 tiles)
  }
  static const int res_prefetch_ok = env_int("LY_TC_RES_PREFETCH", 1);
  p.res_slot = p.epi_groups ? 0 : (((op.res.ptr != nullptr) != (op.up.ptr != nullptr)) && res_prefetch_ok && bn <= 128) ? ((bn / 16 + 3) / 4) * 32 : 0;   // one 32-byte chunk per round
  const long long res_bytes = 2LL * 32 * kEpiWarps * p.res_slot;
  const long long x_bytes = p.fold ? 2LL * 4 * 4 * 48 * 4 : 0;      // fold: boundary-row exchange slots of the epilogue
  p.exch_off = (uint32_t)(bar_bytes + res_bytes);
  // TMA-store epilogue (tile-parallel role, flat / brick pixel mappings, NHWC output): 1 or 2 staging slots of 2 KB per warp
  // Only when the weights are resident: with streamed weights the 64 KB of slots cost pipeline stages (measured: pair-mode
  // 3x3 128->128 @40^2 0.111 -> 0.160 ms with two slots, neutral with one), with resident weights the A ring has room
  // to spare (1x1 256->128 @80^2 0.217 -> 0.201 ms, 3x3/s2 32->64 @320^2 0.469 -> 0.406 ms, the upsample-folded 1x1
  // 128->128 @80^2 0.230 -> 0.193 ms).  N tiles must be whole 32-channel groups (a partial group is clipped by the tensor
  // map only at the END of the channel range, not at an N-tile boundary).
  // (Tried for the column-parallel role too -- N = 256, a warp owning four consecutive chunks = two 32-channel groups: bit-correct,
  //  but slower on 7 of 10 wide 1x1 shapes (512->256 @40^2 0.118 -> 0.134 ms, 512->512 @20^2 0.060 -> 0.069, resident 128->256 @80^2
  //  0.237 -> 0.243; only 256->512 @40^2 gained, 0.155 -> 0.142): those layers stream their weights and need the shared memory
  //  for ring stages.  Not kept.)
  static const int ts_env = env_int("LY_TC_TMASTORE", 2);
  p.ts = (ts_env && p.epi_groups == 4 && p.halo != 2 && op.dst.ptr && !op.nchw && bn % 32 == 0 && (p.b_resident || ts_env >= 3))
             ? (ts_env == 1 ? 1 : 2) : 0;
  p.stg_off = (uint32_t)((bar_bytes + res_bytes + x_bytes + 1023) / 1024 * 1024);
  const long long stg_bytes = p.ts ? (long long)p.stg_off - (bar_bytes + res_bytes + x_bytes) + 32LL * kEpiWarps * 64 * p.ts : 0;
  const long long avail = (long long)kSmemBudget - bar_bytes - 1024 - res_bytes - x_bytes - stg_bytes - (p.b_resident ? b_all : 0);
  if (p.halo == 2) {
    // a band keeps kc_blocks stages for all of its tiles; the next band is prefetched meanwhile
    p.a_stages = p.b_resident ? (int)(avail / p.a_stage) : 2 * p.kc_blocks;
    p.b_stages = p.b_resident ? 1 : (int)((avail - (long long)p.a_stages * p.a_stage) / p.b_stage);
    if (p.a_stages < 2 * p.kc_blocks) { delete st; set_error("conv_tc: band does not fit in shared memory"); return LY_E_ARG; }
  } else if (p.pair) {
    // one k-step of a tile pair = 2 A stages + tpa weight slabs: size both rings in whole steps
    const long long per_step = 2LL * p.a_stage + (long long)p.tpa * p.b_stage;
    long long steps = avail / per_step;
    if (steps < 1) steps = 1;
    p.a_stages = (int)(2 * steps);
    p.b_stages = (int)((avail - (long long)p.a_stages * p.a_stage) / p.b_stage);
  } else if (p.b_resident) {
    p.a_stages = (int)(avail / p.a_stage);
    p.b_stages = 1;
  } else {
    // split the budget so that both rings hold about the same number of k-blocks
    const long long per_kb = p.a_stage / p.tpa + p.b_stage;
    long long kbs = avail / per_kb;
    p.a_stages = (int)(kbs / p.tpa);
    if (p.a_stages < 2) p.a_stages = 2;
    p.b_stages = (int)((avail - (long long)p.a_stages * p.a_stage) / p.b_stage);
  }
  if (p.a_stages > kMaxStages) p.a_stages = kMaxStages;
  if (p.b_stages > kMaxStages) p.b_stages = kMaxStages;
  if (p.a_stages < 2 || (!p.b_resident && p.b_stages < 2)) { delete st; set_error("conv_tc: tile does not fit in shared memory"); return LY_E_ARG; }
  st->smem = 1024 + (size_t)p.a_stages * p.a_stage + (p.b_resident ? (size_t)b_all : (size_t)p.b_stages * p.b_stage) + bar_bytes +
             (size_t)res_bytes + (size_t)x_bytes + (size_t)stg_bytes;
  if (st->smem < 120 * 1024) st->smem = 120 * 1024;  // force one CTA per SM (TMEM allocations must not contend)

  // TMEM accumulator stages.  Knock-out runs (-DLY_TC_EXP) showed that with two stages the thin layers are bound by the
  // issuer -> epilogue -> issuer round trip of an accumulator (commit, mbarrier wake-ups, 16 arrivals: ~1400 cycles), not
  // by the MMAs or the epilogue's work: per tile max(T_mma, T_epi, (T_mma + T_epi + L) / stages).  All 512 columns are ours
  // (one CTA per SM), so use as many stages as fit.
  {
    static const int acc_env = env_int("LY_TC_ACC", kMaxAcc);
    const int per = p.fold ? 3 * bn : (p.pair ? 2 * bn : bn);
    int acc = 512 / per;
    if (acc > acc_env) acc = acc_env;
    if (acc > kMaxAcc) acc = kMaxAcc;
    if (acc < 2) acc = 2;
    p.acc = acc;
    p.tmem_cols = pow2_ge(acc * per);
    if (p.tmem_cols > 512) { delete st; set_error("conv_tc: accumulators exceed TMEM"); return LY_E_ARG; }
  }

  // descriptors
  const int swz = p.kc == 64 ? 2 : (p.kc == 32 ? 4 : 6);       // UMMA LayoutType: SW128 / SW64 / SW32
  const uint32_t sbo = (uint32_t)(8 * p.kc * 2) >> 4;          // 8-row group stride, 16-byte units
  p.desc_hi = (sbo & 0x3FFFu) | (1u << 14) /*version = 1 (sm_100)*/ | ((uint32_t)swz << 29);
  // stride-2 halo mode: output row oy of the brick reads input row 2*oy + ky, so consecutive 8-pixel
  // row groups of the A operand are TWO stored groups apart
  const uint32_t sbo_a = sbo * (uint32_t)((p.halo && op.stride == 2) ? 2 : 1);
  p.desc_hi_a = (sbo_a & 0x3FFFu) | (1u << 14) | ((uint32_t)swz << 29);
  p.idesc = (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ | ((uint32_t)(bn >> 3) << 17) | ((128u >> 4) << 24);
  // (measured: neutral -- conv_tc total 9.45 vs 9.49 ms per step, step time within run-to-run noise; off by default)
  static const int xpre_env = env_int("LY_TC_XPRE", 0);
  p.xpre = xpre_env;
  static const int exp_env = env_int("LY_TC_EXP", 0);
  p.exp = exp_env;
  p.idesc_fold = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((3 * bn) >> 3) << 17) | ((128u >> 4) << 24);

  const CUtensorMapSwizzle tswz = p.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  {
    char* base = (char*)op.src.ptr + (size_t)op.src.c0 * 2;
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4], estr[4];
    if (flat) {
      dims[0] = Cin; dims[1] = (cuuint64_t)dimW; dims[2] = 1; dims[3] = 1;
      strides[0] = (cuuint64_t)op.src.ctot * 2; strides[1] = strides[0] * dimW; strides[2] = strides[1];
    } else {
      dims[0] = Cin; dims[1] = op.src.W; dims[2] = op.src.H; dims[3] = op.B;
      strides[0] = (cuuint64_t)op.src.ctot * 2; strides[1] = strides[0] * op.src.W; strides[2] = strides[1] * op.src.H;
    }
    // never promote an L2 fill beyond the bytes a box row really uses: a 64-byte slice of a wider
    // pixel would otherwise cost a 128-byte DRAM fetch (measured: 2x read traffic on C2f slices)
    const CUtensorMapL2promotion a_promo = p.kc == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                          : (p.kc == 32 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE);
    box[0] = p.kc; box[1] = p.tw * op.stride; box[2] = (p.halo ? halo_rows : p.th * op.stride); box[3] = p.tb;
    if (p.halo == 2) { box[1] = p.band_w; box[2] = p.band_r + 2; box[3] = 1; }
    estr[0] = 1; estr[1] = op.stride; estr[2] = p.halo ? 1 : op.stride; estr[3] = 1;   // halo: every input row is loaded
    CUresult r = encode(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        tswz, a_promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete st; set_error("conv_tc: cuTensorMapEncodeTiled(A) failed with %d", (int)r); return LY_E_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)op.k * op.k * Cin, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {dims[0] * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.kc, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)op.w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        tswz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete st; set_error("conv_tc: cuTensorMapEncodeTiled(B) failed with %d", (int)r); return LY_E_CUDA; }
  }

  if (p.ts) {
    // the 32 pixels of an epilogue warp (rows 32q .. 32q+31 of the tile) are an aligned sub-box of the pixel brick
    cuuint64_t dims[4], strides[3];
    cuuint32_t box[4] = {32, 32, 1, 1}, estr[4] = {1, 1, 1, 1};
    const cuuint64_t pitch = (cuuint64_t)op.dst.ctot * 2;
    if (flat) {
      dims[0] = (cuuint64_t)Cout; dims[1] = (cuuint64_t)dimW; dims[2] = 1; dims[3] = 1;
      strides[0] = pitch; strides[1] = pitch * dimW; strides[2] = strides[1];
    } else {
      dims[0] = (cuuint64_t)Cout; dims[1] = (cuuint64_t)Wo; dims[2] = (cuuint64_t)Ho; dims[3] = (cuuint64_t)op.B;
      strides[0] = pitch; strides[1] = pitch * Wo; strides[2] = strides[1] * Ho;
      if (p.tw >= 32) { box[1] = 32; box[2] = 1; box[3] = 1; }
      else if (p.tw * p.th >= 32) { box[1] = p.tw; box[2] = 32 / p.tw; box[3] = 1; }
      else { box[1] = p.tw; box[2] = p.th; box[3] = 32 / (p.tw * p.th); }
    }
    char* base = (char*)op.dst.ptr + (size_t)op.dst.c0 * 2;
    CUresult r = encode(&p.tmD, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete st; set_error("conv_tc: cuTensorMapEncodeTiled(D) failed with %d", (int)r); return LY_E_CUDA; }
  }
  p.dst = (__nv_bfloat16*)op.dst.ptr; p.dCtot = op.dst.ctot; p.dC0 = op.dst.c0;
  static const int st256_ok = env_int("LY_ST256", 1);
  p.st256 = st256_ok && op.dst.ptr && op.dst.ctot % 16 == 0 && op.dst.c0 % 16 == 0 && reinterpret_cast<uintptr_t>(op.dst.ptr) % 32 == 0;
  p.res = (const __nv_bfloat16*)op.res.ptr; p.rCtot = op.res.ctot; p.rC0 = op.res.c0;
  {
    static const int ld256_ok = env_int("LY_LD256", 5);
    const ly_view& av = op.res.ptr ? op.res : op.up;
    // Variant (LY_LD256 = 1..4 forces one; default 5 = choose): the rows are read once, so they bypass L1 (no_allocate: 3x3 64->64
    // @80^2 + shortcut 0.148 -> 0.140 ms); when a row slice is only 64 bytes of a wider pixel (the 32-channel C2f parts of a
    // 96-channel concat buffer) the L2 fill is also capped at 64 bytes (0.396 -> 0.323 ms; with wider rows that cap costs 8 %).
    const bool ok256 = ld256_ok && av.ptr && av.ctot % 16 == 0 && av.c0 % 16 == 0 && reinterpret_cast<uintptr_t>(av.ptr) % 32 == 0;
    p.ld256 = !ok256 ? 0 : (ld256_ok >= 5 ? (bn <= 32 && av.ctot > bn ? 4 : 3) : ld256_ok);
  }
  p.up = (const __nv_bfloat16*)op.up.ptr; p.uCtot = op.up.ctot; p.uC0 = op.up.c0; p.uH = op.up.H; p.uW = op.up.W; p.Wreal = Wo;
  p.bias = op.bias;
  p.nchw = op.nchw; p.nCtot = op.nchw_ctot; p.nC0 = op.nchw_c0; p.nC = op.nchw_c;

  const int sms = sm_count();
  const int work = p.pair ? (p.total_tiles + 1) / 2 : p.total_tiles;
  st->grid = work < sms ? work : sms;
  static const int debug = env_int("LY_TC_DEBUG", 0);
  if (debug)
    fprintf(stderr, "[conv_tc] k%d s%d %dx%d cin %d cout %d B %d: mode %d fold %d pair %d bn %d tiles_n %d kc %d a_stages %d (%d B) b_res %d b_stages %d res_slot %d "
            "band R %d mt %d bands %d units %d smem %zu\n", op.k, op.stride, op.src.H, op.src.W, Cin, Cout, op.B, p.halo, p.fold, p.pair, bn, p.tiles_n, p.kc,
            p.a_stages, p.a_stage, p.b_resident, p.b_stages, p.res_slot, p.band_r, p.band_mt, p.bands, p.total_tiles, st->smem);
  static std::atomic<unsigned long long> attr_devs{0};   // per DEVICE: the attribute does not carry over to another GPU
  if (first_on_device(attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024 - (int)sizeof(float) * kMaxCout - 1024 /* static smem: s_bias */);
    if (e != cudaSuccess) { delete st; set_error("conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return LY_E_CUDA; }
  }
  *out = st;
  return LY_OK;
}

int32_t conv_tc_launch(const ConvTcState* st, float* nchw_override, cudaStream_t s) {
  if (nchw_override) {
    Params p = st->p;
    p.nchw = nchw_override;
    launch_k(conv_tc_kernel, dim3(st->grid), dim3(kThreads), st->smem, s, p);
  } else {
    launch_k(conv_tc_kernel, dim3(st->grid), dim3(kThreads), st->smem, s, st->p);
  }
  return post_launch("conv_tc");
}

void conv_tc_free(ConvTcState* st) { delete st; }

}  // namespace ly
