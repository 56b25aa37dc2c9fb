// Fused depthwise 3x3 (+BN +SiLU) -> pointwise 1x1 (+BN +SiLU / +bias) with BOTH stages on the
// tensor cores (bf16 hot path, C <= 128 per launch, Cout <= 128).
//
// The first fused kernel (dwpw_tc.cu) computes the depthwise stage on CUDA cores.  ncu on the
// stride-8 class branch (C = 128): 24 000 warp-instructions per 128-pixel tile, of which 6 100
// FFMA, 1 000 MUFU, 1 300 LDS and ~5 000 barrier polling: the kernel is bound by the instruction
// issue rate at 2.3 TB/s while the tensor pipe idles at 7 %.  Here the depthwise conv is issued to
// the idle pipe instead:
//
//   D1[m, c] = sum over taps t  X[m + off(t), c] * w[t, c]
//
// is, per group of 16 channels g, a GEMM with a DIAGONAL 16 x 16 weight block:
//   D1[:, 16g:16g+16] += A_t[:, 16g:16g+16] (128 x 16)  *  diag(w[t, 16g:16g+16]) (16 x 16)
// i.e. one tcgen05.mma M=128 N=16 K=16 per (tap, group).  15/16 of those MACs multiply zeros,
// but the instruction is bound by reading its A operand from shared memory (128 rows x 32 bytes),
// not by math, and the pipe has nothing else to do.  The A operand of tap (ky, kx) is the SAME
// TMA-loaded halo tile viewed ky*(tw+2)+kx rows further down (the swizzle is a function of the
// absolute shared-memory address, see conv_tc.cu band mode), a 16-channel group is a 32-byte
// K-slice of the 128-byte swizzled rows.
//
//   warp 0       TMA: halo box (64 ch, tw+2, R+2) per k-block -> raw ring; weights once
//   warp 1       MMA: per k-block 36 depthwise MMAs (N=16) into a D1 slot (64 TMEM columns), then,
//                two k-blocks behind, the 1x1 GEMM (4 MMAs, K=64, N=Cout) on the activated A2 tile
//   warps 2-9    middle stage: D1 -> +bias, SiLU -> bf16 -> A2 tile (128B-swizzled K-major smem)
//   warps 10-17  epilogue: D2 -> +bias (+SiLU) -> NHWC bf16 slice or the public NCHW fp32 tensor
//
// A tile is tw x R output pixels of one image; accumulator row m = oy*(tw+2)+ox, so 2 of every
// tw+2 rows are discarded ((R-1)*(tw+2)+tw <= 128).
//
// (Round 2, two measured dead ends, both bit-correct and removed again:
//  * a TMA-store epilogue -- the four quarter warps of a column half stage a 32-channel group of the tile compacted to
//    tw x R pixels, one named barrier, one box store: 128->128 @80^2 0.259 vs 0.261 ms with two slots, 0.288 with one.
//    Unlike the resident-weight 1x1 layers of conv_tc this kernel is not bound by its store wavefronts.
//  * a hybrid depthwise stage -- the first ct taps computed by the 8 middle warps on the CUDA cores straight from the
//    swizzled raw tile (fp32 weights in shared memory, FFMA2), the other 9 - ct as diagonal MMAs: 0.264 ms at ct = 0,
//    0.328 / 0.354 / 0.385 / 0.445 / 0.465 ms at ct = 2 / 3 / 4 / 6 / 8: every tap moved costs ~280 cycles per k-block on
//    two warps per scheduler, six times the 48 cycles of the four MMAs it replaces.)
//
// Reference semantics: leanyolo/models/yolov10/head.py:95-107 (v10Detect class branch),
// layers.py:256-264 (CIB).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tma.cuh"
#include "tc.cuh"

namespace ly {

namespace {

constexpr int kMidWarps = 8, kFinWarps = 8;
constexpr int kThreads = 64 + 32 * (kMidWarps + kFinWarps);
#ifndef LY_G_NR
#define LY_G_NR 2
#endif
#ifndef LY_G_NA2
#define LY_G_NA2 4
#endif
#ifndef LY_G_LAG
#define LY_G_LAG 3
#endif
// (measured on the stride-8 class branch, C = 128: raw/A2/lag = 3/3/2 0.279 ms, 3/3/1 0.381, 2/4/2 0.254, 2/4/3 0.252,
//  3/4/3 0.266: the issuer runs ahead of the tensor pipe, so the 1x1 must lag by more than the middle stage's latency)
constexpr int kNR = LY_G_NR;    // raw halo stages
constexpr int kNA2 = LY_G_NA2;  // activated A2 stages
constexpr int kND1 = 4;         // D1 slots (64 TMEM columns each)
constexpr int kLag = LY_G_LAG;  // the 1x1 of k-block i is issued together with the depthwise MMAs of k-block i + kLag
static_assert(kLag < kNA2 && kLag < kND1 && kNR <= 4 && kNA2 <= 4, "pipeline depths");
constexpr int kA2Stage = 128 * 128;
constexpr int kDiagTile = 512;  // 16 x 16 bf16, 32-byte swizzle

struct GParams {
  CUtensorMap tmIn;   // (C, W, H, B), box (64, tw+2, R+2, 1), 128-byte swizzle
  CUtensorMap tmB;    // (K = Cin, N = Cout), box (64, Cout), 128-byte swizzle
  int tw, R, bw;
  int tiles_x, tiles_y, total_tiles;
  uint32_t mg_x, mg_y, mg_bw;
  int H, W, B;
  int kblocks, cin, cout;
  int raw_stage, raw_box, b_stage, b_box, tmem_cols;
  uint32_t idesc_dw, idesc_pw, hi128, hi32;
  int pre_act, act;
  const __nv_bfloat16* dww;   // [9][cin]
  const float* dwb;
  const float* bias;
  __nv_bfloat16* dst; int dCtot, dC0;
  float* nchw; int nCtot, nC0, nC;
  int st256;
};

__device__ __forceinline__ void g_split(const GParams& p, int tile, int& xt, int& yt, int& b) {
  uint32_t t = (uint32_t)tile;
  uint32_t qx = p.mg_x ? __umulhi(t, p.mg_x) : t; xt = (int)(t - qx * (uint32_t)p.tiles_x); t = qx;
  uint32_t qy = p.mg_y ? __umulhi(t, p.mg_y) : t; yt = (int)(t - qy * (uint32_t)p.tiles_y); b = (int)qy;
}

__global__ void __launch_bounds__(kThreads, 1) dwpw_mma_kernel(const __grid_constant__ GParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // carve: [raw ring][A2 ring][W2 resident][diag tiles][biases][barriers]
  const uint32_t raw_base = base;
  const uint32_t a2_base = raw_base + (uint32_t)kNR * p.raw_stage;
  const uint32_t w2_base = a2_base + (uint32_t)kNA2 * kA2Stage;
  const uint32_t dg_base = w2_base + (uint32_t)p.kblocks * p.b_stage;
  const uint32_t n_diag = 9u * (uint32_t)(p.cin >> 4);
  const uint32_t f_off = (dg_base - base) + n_diag * kDiagTile;
  float* dwb_s = reinterpret_cast<float*>(gen + f_off);        // [cin]   depthwise bias (x 0.5 when pre_act)
  float* pwb_s = dwb_s + p.cin;                                // [cout]  pointwise bias (x 0.5 when act)
  const uint32_t bar_base = base + f_off + (uint32_t)(p.cin + p.cout) * 4u;
  auto rawfull = [&](int s) { return bar_base + 8u * s; };
  auto rawempty = [&](int s) { return bar_base + 8u * (4 + s); };
  auto d1full = [&](int s) { return bar_base + 8u * (8 + s); };
  auto d1empty = [&](int s) { return bar_base + 8u * (12 + s); };
  auto a2full = [&](int s) { return bar_base + 8u * (16 + s); };
  auto a2empty = [&](int s) { return bar_base + 8u * (20 + s); };
  auto d2full = [&](int s) { return bar_base + 8u * (24 + s); };
  auto d2empty = [&](int s) { return bar_base + 8u * (26 + s); };
  const uint32_t wfull = bar_base + 8u * 28;
  const uint32_t tmem_slot = bar_base + 8u * 29;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  // ---- parameters (not produced by the previous kernel): biases and the diagonal weight tiles
  {
    const float ps = p.pre_act ? 0.5f : 1.0f, as = p.act ? 0.5f : 1.0f;   // SiLU(x) = h + h*tanh(h), h = x/2
    for (int i = threadIdx.x; i < p.cin; i += kThreads) dwb_s[i] = ps * p.dwb[i];
    for (int i = threadIdx.x; i < p.cout; i += kThreads) pwb_s[i] = as * p.bias[i];
    uint4* dz = reinterpret_cast<uint4*>(gen + (dg_base - base));
    for (uint32_t i = threadIdx.x; i < n_diag * (kDiagTile / 16); i += kThreads) dz[i] = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  {
    // tile (tap, g): row n (output channel 16g+n) holds w[tap][16g+n] at k = n.  32-byte swizzle:
    // the 16-byte chunk index of a row is XORed with address bit 7 = (n >> 2) & 1.
    const int groups = p.cin >> 4;
    for (int i = threadIdx.x; i < 9 * p.cin; i += kThreads) {
      const int tap = i / p.cin, c = i - tap * p.cin;
      const int g = c >> 4, n = c & 15;
      const uint32_t chunk = (uint32_t)(n >> 3) ^ (uint32_t)((n >> 2) & 1);
      const uint32_t off = (uint32_t)(tap * groups + g) * kDiagTile + (uint32_t)n * 32u + (chunk << 4) + (uint32_t)(n & 7) * 2u;
      *reinterpret_cast<__nv_bfloat16*>(gen + (dg_base - base) + off) = p.dww[i];
    }
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmIn) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmB) : "memory");
    for (int s = 0; s < 4; ++s) {
      mbar_init(rawfull(s), 1);
      mbar_init(rawempty(s), 1);
      mbar_init(d1full(s), 1);
      mbar_init(d1empty(s), kMidWarps);
      mbar_init(a2full(s), kMidWarps);
      mbar_init(a2empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(d2full(s), 1);
      mbar_init(d2empty(s), kFinWarps);
    }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the diagonal tiles were written by the generic proxy
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t d2_col0 = (uint32_t)kND1 * 64u;
  const int KB = p.kblocks;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      mbar_expect_tx(wfull, (uint32_t)p.b_box * (uint32_t)KB);
      for (int kb = 0; kb < KB; ++kb) tma_load_2d(w2_base + kb * p.b_stage, &p.tmB, wfull, kb * 64, 0);
      int rs = 0;
      uint32_t rp = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int xt, yt, b;
        g_split(p, tile, xt, yt, b);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(rawempty(rs), rp ^ 1u);
          mbar_expect_tx(rawfull(rs), (uint32_t)p.raw_box);
          tma_load_4d(raw_base + rs * p.raw_stage, &p.tmIn, rawfull(rs), kb * 64, xt * p.tw - 1, yt * p.R - 1, b);
          if (++rs == kNR) { rs = 0; rp ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ================================
    if (elect_one()) {
      const uint32_t hi128 = p.hi128, hi32 = p.hi32, idw = p.idesc_dw, ipw = p.idesc_pw;
      const int groups = p.cin >> 4;
      const uint32_t bw8 = (uint32_t)p.bw * 8u;                        // one image row of the halo tile, 16-byte units
      mbar_wait(wfull, 0);
      tc_fence_after();
      // state of the depthwise stream (item i) and of the lagging 1x1 stream (item j)
      int rs = 0, d1 = 0, a2 = 0, acc = 0, jkb = 0;
      uint32_t rp = 0, d1p = 0, a2p = 0, accp = 0;
      // One round = the depthwise MMAs of k-block i plus the 1x1 MMAs of k-block i - kLag, behind ONE set of
      // barrier waits and one tcgen05 fence (measured with tools/mma_bench.cu: a commit + wait + fence round costs
      // ~130 cycles of issue time, an M=128 MMA 48 cycles for any N <= 32).
      auto pw_wait = [&]() {
        if (jkb == 0) mbar_wait(d2empty(acc), accp ^ 1u);
        mbar_wait(a2full(a2), a2p);
      };
      auto pw_issue = [&]() {
        const uint32_t d_tmem = tmem_base + d2_col0 + (uint32_t)acc * (uint32_t)p.cout;
        const uint32_t alo = ((a2_base + (uint32_t)a2 * kA2Stage) >> 4) | (1u << 16);
        const uint32_t blo = ((w2_base + (uint32_t)jkb * p.b_stage) >> 4) | (1u << 16);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t da = ((uint64_t)hi128 << 32) | (uint64_t)(alo + 2 * kk);
          const uint64_t db = ((uint64_t)hi128 << 32) | (uint64_t)(blo + 2 * kk);
          umma_bf16(d_tmem, da, db, ipw, (jkb | kk) != 0 ? 1u : 0u);
        }
        umma_commit(a2empty(a2));
        if (++a2 == kNA2) { a2 = 0; a2p ^= 1u; }
        if (++jkb == KB) {
          jkb = 0;
          umma_commit(d2full(acc));
          if (++acc == 2) { acc = 0; accp ^= 1u; }
        }
      };
      long long issued = 0, done_pw = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < KB; ++kb) {
          const bool with_pw = issued + 1 - done_pw > kLag;
          mbar_wait(rawfull(rs), rp);
          mbar_wait(d1empty(d1), d1p ^ 1u);
          if (with_pw) pw_wait();
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)d1 * 64u;
          const uint32_t alo = ((raw_base + (uint32_t)rs * p.raw_stage) >> 4) | (1u << 16);
          const uint32_t blo = ((dg_base + (uint32_t)(kb * 4) * kDiagTile) >> 4) | (1u << 16);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            const uint32_t at = alo + (uint32_t)(tap / 3) * bw8 + (uint32_t)(tap % 3) * 8u;
            const uint32_t bt = blo + (uint32_t)(tap * groups) * (kDiagTile >> 4);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint64_t da = ((uint64_t)hi128 << 32) | (uint64_t)(at + 2 * g);
              const uint64_t db = ((uint64_t)hi32 << 32) | (uint64_t)(bt + (uint32_t)g * (kDiagTile >> 4));
              umma_bf16(d_tmem + 16u * g, da, db, idw, tap != 0 ? 1u : 0u);
            }
          }
          umma_commit(rawempty(rs));
          umma_commit(d1full(d1));
          if (++rs == kNR) { rs = 0; rp ^= 1u; }
          if (++d1 == kND1) { d1 = 0; d1p ^= 1u; }
          ++issued;
          if (with_pw) { pw_issue(); ++done_pw; }
        }
      }
      while (done_pw < issued) {
        pw_wait();
        tc_fence_after();
        pw_issue();
        ++done_pw;
      }
    }
  } else if (warp < 2 + kMidWarps) {
    // ============================== middle stage ===============================
    // warp -> TMEM lane quarter q = warp % 4 (rows 32q..32q+31), column half = (warp - 2) / 4 of the
    // 64-column D1 slot.  A thread owns one row: 32 channels -> 4 swizzled 16-byte chunks of the A2 row.
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t row = (uint32_t)(q * 32 + lane);
    const float pre = p.pre_act ? 0.5f : 1.0f;
    const bool pre_act = p.pre_act != 0;
    int d1 = 0, a2 = 0, kb = 0;
    uint32_t d1p = 0, a2p = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (kb = 0; kb < KB; ++kb) {
        mbar_wait(d1full(d1), d1p);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(d1 * 64 + half * 32);
        uint32_t r0[16], r1[16];
        tmem_ld16(taddr, r0);
        tmem_ld16(taddr + 16, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d1empty(d1));
        if (++d1 == kND1) { d1 = 0; d1p ^= 1u; }
        const float* bp = dwb_s + kb * 64 + half * 32;
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t* src = j < 8 ? r0 : r1;
          const int jj = (j & 7) * 2;
          const float2 b2 = *reinterpret_cast<const float2*>(bp + 2 * j);
          float v0, v1;
          ffma2(v0, v1, __uint_as_float(src[jj]), __uint_as_float(src[jj + 1]), pre, pre, b2.x, b2.y);
          if (pre_act) silu2_from_half(v0, v1);
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
          packed[j] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        mbar_wait(a2empty(a2), a2p ^ 1u);
        const uint32_t ab = a2_base + (uint32_t)a2 * kA2Stage + row * 128u;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const uint32_t chunk = (uint32_t)(half * 4 + cc) ^ (row & 7u);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ab + (chunk << 4)), "r"(packed[4 * cc]), "r"(packed[4 * cc + 1]),
                       "r"(packed[4 * cc + 2]), "r"(packed[4 * cc + 3])
                       : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(a2full(a2));
        if (++a2 == kNA2) { a2 = 0; a2p ^= 1u; }
      }
    }
  } else {
    // ============================== epilogue ===================================
    const int q = warp & 3, half = (warp - 2 - kMidWarps) >> 2;
    const int nchunks = p.cout >> 4;
    const int c_half = (nchunks + 1) >> 1;
    const int cbeg = half ? c_half : 0, cend = half ? nchunks : c_half;
    const uint32_t row = (uint32_t)(q * 32 + lane);
    const uint32_t oy = __umulhi(row, p.mg_bw), ox = row - oy * (uint32_t)p.bw;
    const bool row_ok = ox < (uint32_t)p.tw && oy < (uint32_t)p.R;
    const bool act = p.act != 0;
    const float pre = act ? 0.5f : 1.0f;
    const uint32_t hw = (uint32_t)p.H * (uint32_t)p.W;
    int acc = 0;
    uint32_t accp = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int xt, yt, b;
      g_split(p, tile, xt, yt, b);
      const uint32_t w = (uint32_t)(xt * p.tw) + ox, h = (uint32_t)(yt * p.R) + oy;
      const bool valid = row_ok && w < (uint32_t)p.W && h < (uint32_t)p.H;
      const uint32_t lin = ((uint32_t)b * (uint32_t)p.H + h) * (uint32_t)p.W + w;
      __nv_bfloat16* drow = (p.dst && valid) ? p.dst + (size_t)lin * (uint32_t)p.dCtot + p.dC0 : nullptr;
      float* nrow = (p.nchw && valid) ? p.nchw + (size_t)((uint32_t)b * (uint32_t)p.nCtot + (uint32_t)p.nC0) * hw + (h * (uint32_t)p.W + w) : nullptr;
      mbar_wait(d2full(acc), accp);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + d2_col0 + (uint32_t)(acc * p.cout);
      uint32_t nxt[16];
      if (cbeg < cend) tmem_ld16(taddr + cbeg * 16, nxt);
      for (int ch = cbeg; ch < cend; ++ch) {
        const int c = ch * 16;
        float v[16];
        tmem_ld_wait();
        const float4* bp = reinterpret_cast<const float4*>(pwb_s + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 bb = bp[j];
          ffma2(v[4 * j + 0], v[4 * j + 1], __uint_as_float(nxt[4 * j + 0]), __uint_as_float(nxt[4 * j + 1]), pre, pre, bb.x, bb.y);
          ffma2(v[4 * j + 2], v[4 * j + 3], __uint_as_float(nxt[4 * j + 2]), __uint_as_float(nxt[4 * j + 3]), pre, pre, bb.z, bb.w);
        }
        if (ch + 1 < cend) {
          tmem_ld16(taddr + c + 16, nxt);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d2empty(acc));
        }
        if (act) {
#pragma unroll
          for (int j = 0; j < 16; j += 2) silu2_from_half(v[j], v[j + 1]);
        }
        if (drow) {
          if (p.st256) {
            store_bf16x16(drow + c, v);
          } else {
            store_vec<__nv_bfloat16>(drow + c, v);
            store_vec<__nv_bfloat16>(drow + c + 8, v + 8);
          }
        }
        if (nrow) {
          float* np = nrow + (size_t)c * hw;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c + j < p.nC) np[(size_t)j * hw] = v[j];
        }
      }
      if (cbeg >= cend) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(d2empty(acc));
      }
      if (++acc == 2) { acc = 0; accp ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

}  // namespace

struct DwPwMmaState {
  GParams p;
  int grid;
  size_t smem;
};

bool dwpw_mma_supported(const ly_op& op) {
  static const int enabled = getenv("LY_DWPW_MMA") ? atoi(getenv("LY_DWPW_MMA")) : 1;
  if (!enabled) return false;
  if (op.dtype != LY_BF16 || op.kind != LY_OP_DWPW) return false;
  if (op.pre_k != 3 || op.k != 1 || op.stride != 1) return false;
  if (op.src.c % 64 || op.src.c > 128 || op.src.c0 % 8 || op.src.ctot % 8) return false;
  const int cout = op.dst.ptr ? op.dst.c : (op.nchw_c + 15) / 16 * 16;
  if (cout % 16 || cout > 128) return false;
  if (op.dst.ptr && (op.dst.c0 % 8 || op.dst.ctot % 8)) return false;
  if (op.res.ptr) return false;
  return true;
}

int32_t dwpw_mma_prepare(const ly_op& op, DwPwMmaState** out) {
  LY_CHECK_ARG(dwpw_mma_supported(op), "dwpw_mma: unsupported op");
  LY_CHECK_ARG(op.src.ptr && op.w && op.bias && op.pre_w && op.pre_bias && (op.dst.ptr || op.nchw), "dwpw_mma: null pointer");
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("dwpw_mma: cuTensorMapEncodeTiled entry point not available"); return LY_E_CUDA; }
  DwPwMmaState* st = new DwPwMmaState();
  GParams& p = st->p;
  memset(&p, 0, sizeof(p));
  const int H = op.src.H, W = op.src.W, Cin = op.src.c;
  const int Cout = op.dst.ptr ? op.dst.c : (op.nchw_c + 15) / 16 * 16;
  p.H = H; p.W = W; p.B = op.B; p.cin = Cin; p.cout = Cout; p.kblocks = Cin / 64;
  p.pre_act = op.pre_act; p.act = op.act;
  // tile tw x R: (R-1)*(tw+2)+tw <= 128; most useful accumulator rows over the whole map
  {
    double best = -1.0;
    for (int tw = 1; tw <= W && tw + 2 <= 128; ++tw)
      for (int R = 1; R <= H && (R - 1) * (tw + 2) + tw <= 128; ++R) {
        const long long tiles = (long long)((W + tw - 1) / tw) * ((H + R - 1) / R);
        const double eff = (double)W * H / ((double)tiles * 128.0);
        // prefer shorter halos on ties (fewer bytes through TMA)
        const double score = eff - 1e-4 * (double)(tw + 2) * (R + 2) / (double)(tw * R);
        if (score > best) { best = score; p.tw = tw; p.R = R; }
      }
    LY_CHECK_ARG(best > 0, "dwpw_mma: no tile fits");
  }
  p.bw = p.tw + 2;
  p.tiles_x = (W + p.tw - 1) / p.tw;
  p.tiles_y = (H + p.R - 1) / p.R;
  const long long total = (long long)p.tiles_x * p.tiles_y * op.B;
  auto magic = [](uint32_t d) -> uint32_t { return d <= 1 ? 0u : (uint32_t)((1ull << 32) / d + 1); };
  p.mg_x = magic((uint32_t)p.tiles_x); p.mg_y = magic((uint32_t)p.tiles_y);
  p.mg_bw = (uint32_t)((1ull << 32) / (uint32_t)p.bw + 1);
  {
    const unsigned long long lim = 1ull << 32, tmax = (unsigned long long)total + 2ull * sm_count();
    if (tmax * p.tiles_x >= lim || tmax * p.tiles_y >= lim || (unsigned long long)op.B * H * W >= lim) {
      delete st; set_error("dwpw_mma: problem too large for 32-bit tile arithmetic"); return LY_E_ARG;
    }
  }
  p.total_tiles = (int)total;
  // the tap windows of the last accumulator rows reach 2*bw+2 rows past row 127
  const int box_rows = (p.R + 2) * p.bw, need_rows = 128 + 2 * p.bw + 2;
  p.raw_box = box_rows * 128;
  p.raw_stage = ((box_rows > need_rows ? box_rows : need_rows) * 128 + 1023) / 1024 * 1024;
  p.b_box = Cout * 128;
  p.b_stage = (p.b_box + 1023) / 1024 * 1024;
  p.tmem_cols = 512;   // 4 D1 slots x 64 + 2 accumulators x Cout (<= 128)
  const size_t need = 1024 + (size_t)kNR * p.raw_stage + (size_t)kNA2 * kA2Stage + (size_t)p.kblocks * p.b_stage +
                      (size_t)9 * (Cin / 16) * kDiagTile + (size_t)(Cin + Cout) * 4 + 8 * 32;
  if (need > 226 * 1024) { delete st; set_error("dwpw_mma: tile does not fit in shared memory"); return LY_E_ARG; }
  st->smem = need < 120 * 1024 ? 120 * 1024 : need;   // one CTA per SM (all 512 TMEM columns)

  const uint32_t sbo128 = (uint32_t)(8 * 128) >> 4, sbo32 = (uint32_t)(8 * 32) >> 4;
  p.hi128 = (sbo128 & 0x3FFFu) | (1u << 14) | (2u << 29);   // version 1 (sm_100), SWIZZLE_128B
  p.hi32 = (sbo32 & 0x3FFFu) | (1u << 14) | (6u << 29);     // SWIZZLE_32B
  p.idesc_dw = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(16 >> 3) << 17) | ((128u >> 4) << 24);
  p.idesc_pw = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cout >> 3) << 17) | ((128u >> 4) << 24);
  {
    char* base = (char*)op.src.ptr + (size_t)op.src.c0 * 2;
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)op.B};
    cuuint64_t strides[3] = {(cuuint64_t)op.src.ctot * 2, (cuuint64_t)op.src.ctot * 2 * W, (cuuint64_t)op.src.ctot * 2 * W * H};
    cuuint32_t box[4] = {64, (cuuint32_t)p.bw, (cuuint32_t)(p.R + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&p.tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete st; set_error("dwpw_mma: cuTensorMapEncodeTiled(in) failed with %d", (int)r); return LY_E_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)Cout};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)op.w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete st; set_error("dwpw_mma: cuTensorMapEncodeTiled(B) failed with %d", (int)r); return LY_E_CUDA; }
  }
  p.dww = (const __nv_bfloat16*)op.pre_w; p.dwb = op.pre_bias; p.bias = op.bias;
  p.dst = (__nv_bfloat16*)op.dst.ptr; p.dCtot = op.dst.ctot; p.dC0 = op.dst.c0;
  p.st256 = (getenv("LY_ST256") ? atoi(getenv("LY_ST256")) : 1) && op.dst.ptr && op.dst.ctot % 16 == 0 && op.dst.c0 % 16 == 0 &&
            reinterpret_cast<uintptr_t>(op.dst.ptr) % 32 == 0;
  p.nchw = op.nchw; p.nCtot = op.nchw_ctot; p.nC0 = op.nchw_c0; p.nC = op.nchw_c;
  const int sms = sm_count();
  st->grid = p.total_tiles < sms ? p.total_tiles : sms;
  static std::atomic<unsigned long long> attr_devs{0};   // per DEVICE: the attribute does not carry over to another GPU
  if (first_on_device(attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(dwpw_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { delete st; set_error("dwpw_mma: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return LY_E_CUDA; }
  }
  static const int debug = getenv("LY_TC_DEBUG") ? atoi(getenv("LY_TC_DEBUG")) : 0;
  if (debug)
    fprintf(stderr, "[dwpw_mma] %dx%d cin %d cout %d B %d: tile %d x %d (bw %d) tiles %d raw_stage %d smem %zu\n", H, W, Cin, Cout, op.B,
            p.tw, p.R, p.bw, p.total_tiles, p.raw_stage, st->smem);
  *out = st;
  return LY_OK;
}

int32_t dwpw_mma_launch(const DwPwMmaState* st, float* nchw_override, cudaStream_t s) {
  if (nchw_override) {
    GParams p = st->p;
    p.nchw = nchw_override;
    launch_k(dwpw_mma_kernel, dim3(st->grid), dim3(kThreads), st->smem, s, p);
  } else {
    launch_k(dwpw_mma_kernel, dim3(st->grid), dim3(kThreads), st->smem, s, st->p);
  }
  return post_launch("dwpw_mma");
}

void dwpw_mma_free(DwPwMmaState* st) { delete st; }

}  // namespace ly
