// Fused depthwise 3x3 (+BN +SiLU)  ->  pointwise 1x1 (+BN +SiLU / +bias), bf16 hot path.
//
// The v10Detect class branch and every CIB block are chains of  dw3x3 -> 1x1  pairs
// (leanyolo/models/yolov10/head.py:95-107, layers.py:256-264).  Run separately, the depthwise
// result makes a full HBM round trip (write C channels per pixel, read them back) between
// two bandwidth-bound kernels.  Here the depthwise conv is the PRODUCER of the GEMM's A
// operand: its output only ever exists as a swizzled K-major tile in shared memory.
//
//   warp 0      TMA: (th+2) x (tw+2) halo box of 64 input channels per k-block (zero padding =
//               hardware OOB fill) into a raw ring; 1x1 weight slab [Cout x 64] into a B ring
//   warps 10-25 depthwise producers: a warp owns a 4x2 patch of output pixels, a lane 2 of the 64
//               channels; fp32 accumulate, SiLU, bf16, written as the 128-byte-swizzled A tile
//               (row = pixel of the tile, 64 channels = one swizzle row)
//   warp 1      tcgen05.mma M=128 x N=Cout x K=64 per k-block into a double-buffered TMEM
//               accumulator
//   warps 2-9   epilogue: TMEM -> bias (+SiLU) -> bf16 NHWC slice (or the public NCHW fp32)
//
// The tile is a rectangle of up to 16 such patches of one image (16x8 on 80x80 maps, 20x6 on
// 40x40 and 20x20); rows of the M=128 MMA past tw*th are never written and are masked in
// the epilogue.
#include <cuda.h>
#include <string.h>
#include "common.cuh"
#include "tma.cuh"
#include "tc.cuh"

namespace ly {

namespace {

constexpr int kFEpWarps = 8, kFDwWarps = 16;
constexpr int kPH = 2;                 // patch = 4 wide x kPH tall output pixels per depthwise warp
constexpr int kFThreads = 64 + 32 * (kFEpWarps + kFDwWarps);
constexpr int kFMaxStages = 4;
constexpr uint32_t kFSmemBudget = 216 * 1024;
constexpr int kFAStage = 128 * 128;   // 128 rows x 64 bf16

struct FParams {
  CUtensorMap tmIn;   // (C, W, H, B), box (64, tw+2, th+2, 1), no swizzle
  CUtensorMap tmB;    // (K = Cin, N = Cout), box (64, block_n), 128-byte swizzle
  int tw, th, tile_px, npx, npy;   // tile = npx x npy patches of 4 x kPH pixels (one depthwise warp each)
  int tiles_x, tiles_y, total_tiles;
  uint32_t mg_x, mg_y;
  int H, W, B;
  int kblocks, cin;
  int block_n, tmem_cols;
  int r_stage, r_box, n_r, n_a, b_stage, b_box, n_b;
  int b_resident;        // all k-blocks of the 1x1 weights stay in shared memory (n_b == kblocks)
  uint32_t idesc, desc_hi;
  int pre_act, act;
  const __nv_bfloat16* dww;
  const float* dwb;
  const float* bias;
  __nv_bfloat16* dst; int dCtot, dC0;
  float* nchw; int nCtot, nC0, nC;
  int rev;               // walk the tiles last to first (see g_reverse)
  int st256;             // 32-byte aligned NHWC rows: one 256-bit store per 16-channel chunk
};

__device__ __forceinline__ void f_split(const FParams& p, int tile, int& xt, int& yt, int& b) {
  uint32_t t = (uint32_t)(p.rev ? p.total_tiles - 1 - tile : tile);
  uint32_t qx = p.mg_x ? __umulhi(t, p.mg_x) : t; xt = (int)(t - qx * (uint32_t)p.tiles_x); t = qx;
  uint32_t qy = p.mg_y ? __umulhi(t, p.mg_y) : t; yt = (int)(t - qy * (uint32_t)p.tiles_y); b = (int)qy;
}

__global__ void __launch_bounds__(kFThreads, 1) dwpw_kernel(const __grid_constant__ FParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t a_base = base;
  const uint32_t b_base = a_base + (uint32_t)p.n_a * kFAStage;
  const uint32_t r_base = b_base + (uint32_t)p.n_b * p.b_stage;
  const uint32_t w_off = (r_base - base) + (uint32_t)p.n_r * p.r_stage;
  float* wf = reinterpret_cast<float*>(gen + w_off);            // [9][cin] depthwise weights (x 0.5 when pre_act)
  float* dwb = wf + 9 * p.cin;                                  // [cin]
  float* pwb = dwb + p.cin;                                     // [block_n]
  const uint32_t bar_base = base + w_off + (uint32_t)(10 * p.cin + p.block_n) * 4u;
  auto rfull = [&](int s) { return bar_base + 8u * s; };
  auto rempty = [&](int s) { return bar_base + 8u * (kFMaxStages + s); };
  auto afull = [&](int s) { return bar_base + 8u * (2 * kFMaxStages + s); };
  auto aempty = [&](int s) { return bar_base + 8u * (3 * kFMaxStages + s); };
  auto bfull = [&](int s) { return bar_base + 8u * (4 * kFMaxStages + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (5 * kFMaxStages + s); };
  auto tfull = [&](int s) { return bar_base + 8u * (6 * kFMaxStages + s); };
  auto tempty = [&](int s) { return bar_base + 8u * (6 * kFMaxStages + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (6 * kFMaxStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  {
    const float ps = p.pre_act ? 0.5f : 1.0f, as = p.act ? 0.5f : 1.0f;   // SiLU(x) = h + h*tanh(h), h = x/2
    for (int i = threadIdx.x; i < 9 * p.cin; i += kFThreads) wf[i] = ps * __bfloat162float(p.dww[i]);
    for (int i = threadIdx.x; i < p.cin; i += kFThreads) dwb[i] = ps * p.dwb[i];
    for (int i = threadIdx.x; i < p.block_n; i += kFThreads) pwb[i] = as * p.bias[i];
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmIn) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmB) : "memory");
    for (int s = 0; s < kFMaxStages; ++s) {
      mbar_init(rfull(s), 1);
      mbar_init(rempty(s), kFDwWarps);
      mbar_init(afull(s), kFDwWarps);
      mbar_init(aempty(s), 1);
      mbar_init(bfull(s), 1);
      mbar_init(bempty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), kFEpWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();   // weights / biases above are parameters; activations are read from here on
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      int rs = 0, bs = 0;
      uint32_t rp = 0, bp = 0;
      if (p.b_resident) {
        mbar_expect_tx(bfull(0), (uint32_t)p.b_box * (uint32_t)p.kblocks);
        for (int kb = 0; kb < p.kblocks; ++kb) tma_load_2d(b_base + kb * p.b_stage, &p.tmB, bfull(0), kb * 64, 0);
      }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int xt, yt, b;
        f_split(p, tile, xt, yt, b);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(rempty(rs), rp ^ 1u);
          mbar_expect_tx(rfull(rs), (uint32_t)p.r_box);
          tma_load_4d(r_base + rs * p.r_stage, &p.tmIn, rfull(rs), kb * 64, xt * p.tw - 1, yt * p.th - 1, b);
          if (++rs == p.n_r) { rs = 0; rp ^= 1u; }
          if (!p.b_resident) {
            mbar_wait(bempty(bs), bp ^ 1u);
            mbar_expect_tx(bfull(bs), (uint32_t)p.b_box);
            tma_load_2d(b_base + bs * p.b_stage, &p.tmB, bfull(bs), kb * 64, 0);
            if (++bs == p.n_b) { bs = 0; bp ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ================================
    if (elect_one()) {
      const uint32_t hi = p.desc_hi, idesc = p.idesc;
      int as = 0, bs = 0, acc = 0;
      uint32_t pa = 0, pb = 0, pacc = 0;
      const bool bres = p.b_resident != 0;
      if (bres) {
        mbar_wait(bfull(0), 0);
        tc_fence_after();
      }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty(acc), pacc ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.block_n);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(afull(as), pa);
          if (!bres) mbar_wait(bfull(bs), pb);
          tc_fence_after();
          const uint32_t alo = ((a_base + (uint32_t)as * kFAStage) >> 4) | (1u << 16);
          const uint32_t blo = ((b_base + (uint32_t)(bres ? kb : bs) * p.b_stage) >> 4) | (1u << 16);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = ((uint64_t)hi << 32) | (uint64_t)(alo + 2 * kk);
            const uint64_t db = ((uint64_t)hi << 32) | (uint64_t)(blo + 2 * kk);
            umma_bf16(d_tmem, da, db, idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(aempty(as));
          if (++as == p.n_a) { as = 0; pa ^= 1u; }
          if (!bres) {
            umma_commit(bempty(bs));
            if (++bs == p.n_b) { bs = 0; pb ^= 1u; }
          }
        }
        umma_commit(tfull(acc));
        if (++acc == 2) { acc = 0; pacc ^= 1u; }
      }
    }
  } else if (warp >= 2 + kFEpWarps) {
    // ============================== depthwise producers =======================
    // One warp = one 4x2 patch of output pixels; a lane owns 2 of the 64 channels of the k-block.
    // Every shared-memory access is then a conflict-free 128-byte row (raw tile, A tile) and an
    // input pixel is loaded once per patch (24 loads for 8 outputs) instead of once per output
    // row and strip.  (Measured steps: 16-byte channel vectors x 4-pixel strips re-read weights
    // and neighbouring pixels, the L1/shared pipe ran at 95 % and the fused kernel was slower
    // than the two separate ones; 8 warps x 4x4 patches were bound by the serial instruction
    // stream of each producer warp (IPC 0.2): 16 warps x 4x2 patches; a 4-warp epilogue then became the bottleneck of the 2-k-block layers, so 8.)
    const int pw_i = warp - (2 + kFEpWarps);                 // patch index inside the tile
    const bool active = pw_i < p.npx * p.npy;
    const int pyi = pw_i / p.npx, pxi = pw_i - pyi * p.npx;
    const int iw = p.tw + 2;
    int rs = 0, as = 0;
    uint32_t rp = 0, pa = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(rfull(rs), rp);
        float acc[kPH][4][2];
        if (active) {
          const int c = kb * 64 + 2 * lane;
          float wv[9][2];
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            const float2 w2 = *reinterpret_cast<const float2*>(wf + tp * p.cin + c);
            wv[tp][0] = w2.x; wv[tp][1] = w2.y;
          }
          const float2 b2 = *reinterpret_cast<const float2*>(dwb + c);
#pragma unroll
          for (int y = 0; y < kPH; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) { acc[y][x][0] = b2.x; acc[y][x][1] = b2.y; }
          // halo pixel (X, Y) of the raw tile sits at ((Y * iw + X) * 64 + channel) * 2 bytes
          const uint8_t* rt = gen + (r_base - base) + (size_t)rs * p.r_stage + ((size_t)((kPH * pyi) * iw + 4 * pxi) * 64 + 2 * lane) * 2;
#pragma unroll
          for (int iy = 0; iy < kPH + 2; ++iy) {
            float in[6][2];
#pragma unroll
            for (int ix = 0; ix < 6; ++ix) {
              const uint32_t raw = *reinterpret_cast<const uint32_t*>(rt + (size_t)(iy * iw + ix) * 128);
              in[ix][0] = __uint_as_float(raw << 16);
              in[ix][1] = __uint_as_float(raw & 0xffff0000u);
            }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const int oy = iy - ky;
              if (oy >= 0 && oy < kPH) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                  for (int ox = 0; ox < 4; ++ox)     // the lane's two channels of one output pixel: one packed FFMA2
                    ffma2(acc[oy][ox][0], acc[oy][ox][1], in[ox + kx][0], in[ox + kx][1], wv[ky * 3 + kx][0], wv[ky * 3 + kx][1],
                          acc[oy][ox][0], acc[oy][ox][1]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(rempty(rs));   // this warp is done reading the raw stage
        if (++rs == p.n_r) { rs = 0; rp ^= 1u; }

        mbar_wait(aempty(as), pa ^ 1u);
        if (active) {
          const uint32_t ab = a_base + (uint32_t)as * kFAStage;
#pragma unroll
          for (int y = 0; y < kPH; ++y)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
              float v0 = acc[y][x][0], v1 = acc[y][x][1];
              if (p.pre_act) silu2_from_half(v0, v1);
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
              const uint32_t r = (uint32_t)((kPH * pyi + y) * p.tw + 4 * pxi + x);   // tile row = pixel index, x fastest
              const uint32_t addr = ab + r * 128u + ((((uint32_t)lane >> 2) ^ (r & 7u)) << 4) + ((uint32_t)lane & 3u) * 4u;
              asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(*reinterpret_cast<const uint32_t*>(&h2)) : "memory");
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(afull(as));
        if (++as == p.n_a) { as = 0; pa ^= 1u; }
      }
    }
  } else {
    // ============================== epilogue (2 warps per TMEM lane quarter) ==========
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int nchunks = p.block_n >> 4;
    const int c_half = (nchunks + 1) >> 1;
    const int cbeg = half ? c_half : 0, cend = half ? nchunks : c_half;
    const int row = q * 32 + lane;
    const int py = row / p.tw, px = row - py * p.tw;
    const bool row_ok = row < p.tile_px;
    const float pre = p.act ? 0.5f : 1.0f;
    int acc = 0;
    uint32_t pacc = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int xt, yt, b;
      f_split(p, tile, xt, yt, b);
      const int w = xt * p.tw + px, h = yt * p.th + py;
      const bool valid = row_ok && w < p.W && h < p.H;
      const long long lin = ((long long)b * p.H + h) * p.W + w;
      mbar_wait(tfull(acc), pacc);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.block_n);
      __nv_bfloat16* drow = (p.dst && valid) ? p.dst + lin * p.dCtot + p.dC0 : nullptr;
      float* nrow = (p.nchw && valid) ? p.nchw + ((long long)b * p.nCtot + p.nC0) * ((long long)p.H * p.W) + ((long long)h * p.W + w) : nullptr;
      uint32_t nxt[16];
      if (cbeg < cend) tmem_ld16(taddr + cbeg * 16, nxt);
      for (int ch = cbeg; ch < cend; ++ch) {
        const int c = ch * 16;
        float v[16];
        tmem_ld_wait();
        const float4* bp = reinterpret_cast<const float4*>(pwb + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 bb = bp[j];
          v[4 * j + 0] = fmaf(__uint_as_float(nxt[4 * j + 0]), pre, bb.x);
          v[4 * j + 1] = fmaf(__uint_as_float(nxt[4 * j + 1]), pre, bb.y);
          v[4 * j + 2] = fmaf(__uint_as_float(nxt[4 * j + 2]), pre, bb.z);
          v[4 * j + 3] = fmaf(__uint_as_float(nxt[4 * j + 3]), pre, bb.w);
        }
        if (ch + 1 < cend) {
          tmem_ld16(taddr + c + 16, nxt);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(acc));
        }
        if (p.act) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = silu_from_half(v[j]);
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
          float* np = nrow + (long long)c * p.H * p.W;
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (c + j < p.nC) np[(long long)j * p.H * p.W] = v[j];
        }
      }
      if (cbeg >= cend) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(acc));
      }
      if (++acc == 2) { acc = 0; pacc ^= 1u; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
  }
}

int f_pow2_ge(int v) { int p = 32; while (p < v) p <<= 1; return p; }

}  // namespace

struct DwPwState {
  FParams p;
  int grid;
  size_t smem;
  DwPwMmaState* mma = nullptr;   // set: the launch goes to dwpw_mma_kernel (depthwise stage on the tensor cores)
};

bool dwpw_supported(const ly_op& op) {
  if (op.dtype != LY_BF16 || op.kind != LY_OP_DWPW) return false;
  if (op.pre_k != 3 || op.k != 1 || op.stride != 1) return false;
  if (op.src.c % 64 || op.src.c0 % 8 || op.src.ctot % 8 || op.src.c > 1024) return false;
  const int cout = op.dst.ptr ? op.dst.c : (op.nchw_c + 15) / 16 * 16;
  if (cout % 16 || cout > 256) return false;
  if (op.dst.ptr && (op.dst.c0 % 8 || op.dst.ctot % 8)) return false;
  if (op.res.ptr) return false;
  return true;
}

int32_t dwpw_prepare(const ly_op& op, DwPwState** out) {
  LY_CHECK_ARG(dwpw_supported(op), "dwpw: unsupported op (bf16, dw 3x3 s1 -> 1x1, Cin %% 64 == 0, Cout <= 256, no shortcut)");
  LY_CHECK_ARG(op.src.ptr && op.w && op.bias && op.pre_w && op.pre_bias && (op.dst.ptr || op.nchw), "dwpw: null pointer");
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("dwpw: cuTensorMapEncodeTiled entry point not available"); return LY_E_CUDA; }
  DwPwState* st = new DwPwState();
  if (dwpw_mma_supported(op)) {
    const int32_t rc = dwpw_mma_prepare(op, &st->mma);
    if (rc != LY_OK) { delete st; return rc; }
    *out = st;
    return LY_OK;
  }
  FParams& p = st->p;
  memset(&p, 0, sizeof(p));
  const int H = op.src.H, W = op.src.W, Cin = op.src.c;
  const int Cout = op.dst.ptr ? op.dst.c : (op.nchw_c + 15) / 16 * 16;
  p.H = H; p.W = W; p.B = op.B; p.cin = Cin; p.kblocks = Cin / 64;
  p.block_n = Cout; p.tmem_cols = f_pow2_ge(2 * Cout);
  p.pre_act = op.pre_act; p.act = op.act;
  p.rev = g_reverse;
  // tile = npx x npy patches of 4 x kPH pixels, at most one patch per depthwise warp: fewest tiles
  // per image first, then the smallest halo box
  {
    long long best_tiles = 1LL << 60, best_halo = 1LL << 60;
    for (int npx = 1; npx <= kFDwWarps; ++npx)
      for (int npy = 1; npx * npy <= kFDwWarps; ++npy) {
        const int tw = 4 * npx, th = kPH * npy;
        if (tw * th > 128) continue;
        const long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th);
        const long long halo = (long long)(tw + 2) * (th + 2);
        if (tiles < best_tiles || (tiles == best_tiles && halo < best_halo)) {
          best_tiles = tiles; best_halo = halo; p.npx = npx; p.npy = npy;
        }
      }
  }
  p.tw = 4 * p.npx; p.th = kPH * p.npy;
  p.tile_px = p.tw * p.th;
  p.tiles_x = (W + p.tw - 1) / p.tw;
  p.tiles_y = (H + p.th - 1) / p.th;
  const long long total = (long long)p.tiles_x * p.tiles_y * op.B;
  auto magic = [](uint32_t d) -> uint32_t { return d <= 1 ? 0u : (uint32_t)((1ull << 32) / d + 1); };
  p.mg_x = magic((uint32_t)p.tiles_x); p.mg_y = magic((uint32_t)p.tiles_y);
  {
    const unsigned long long lim = 1ull << 32, tmax = (unsigned long long)total + 2ull * sm_count();
    if (tmax * p.tiles_x >= lim || tmax * p.tiles_y >= lim) { delete st; set_error("dwpw: problem too large for 32-bit tile arithmetic"); return LY_E_ARG; }
  }
  p.total_tiles = (int)total;

  p.r_box = (p.tw + 2) * (p.th + 2) * 128;
  p.r_stage = (p.r_box + 1023) / 1024 * 1024;
  p.b_box = Cout * 128;
  p.b_stage = (p.b_box + 1023) / 1024 * 1024;
  const uint32_t fixed = 1024 + (uint32_t)(10 * Cin + Cout) * 4 + 8 * (6 * kFMaxStages + 8);
  p.n_a = 3; p.n_b = 3; p.n_r = 3;
  if ((long long)p.kblocks * p.b_stage <= 64 * 1024) { p.b_resident = 1; p.n_b = p.kblocks; }
  auto need = [&]() { return (long long)fixed + (long long)p.n_a * kFAStage + (long long)p.n_b * p.b_stage + (long long)p.n_r * p.r_stage; };
  if (need() > kFSmemBudget && !p.b_resident) p.n_b = 2;
  if (need() > kFSmemBudget) p.n_a = 2;
  if (need() > kFSmemBudget) p.n_r = 2;
  if (need() > kFSmemBudget) { delete st; set_error("dwpw: tile does not fit in shared memory"); return LY_E_ARG; }
  while (p.n_r < kFMaxStages && need() + p.r_stage <= kFSmemBudget) ++p.n_r;
  st->smem = (size_t)need();
  if (st->smem < 120 * 1024) st->smem = 120 * 1024;   // one CTA per SM (TMEM allocations must not contend)

  const uint32_t sbo = (uint32_t)(8 * 64 * 2) >> 4;
  p.desc_hi = (sbo & 0x3FFFu) | (1u << 14) | (2u << 29);   // version 1 (sm_100), SWIZZLE_128B
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Cout >> 3) << 17) | ((128u >> 4) << 24);
  {
    char* base = (char*)op.src.ptr + (size_t)op.src.c0 * 2;
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)op.B};
    cuuint64_t strides[3] = {(cuuint64_t)op.src.ctot * 2, (cuuint64_t)op.src.ctot * 2 * W, (cuuint64_t)op.src.ctot * 2 * W * H};
    cuuint32_t box[4] = {64, (cuuint32_t)(p.tw + 2), (cuuint32_t)(p.th + 2), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&p.tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete st; set_error("dwpw: cuTensorMapEncodeTiled(in) failed with %d", (int)r); return LY_E_CUDA; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)Cin, (cuuint64_t)Cout};
    cuuint64_t strides[1] = {(cuuint64_t)Cin * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)Cout};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)op.w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete st; set_error("dwpw: cuTensorMapEncodeTiled(B) failed with %d", (int)r); return LY_E_CUDA; }
  }
  p.dww = (const __nv_bfloat16*)op.pre_w; p.dwb = op.pre_bias; p.bias = op.bias;
  p.dst = (__nv_bfloat16*)op.dst.ptr; p.dCtot = op.dst.ctot; p.dC0 = op.dst.c0;
  p.st256 = op.dst.ptr && op.dst.ctot % 16 == 0 && op.dst.c0 % 16 == 0 && reinterpret_cast<uintptr_t>(op.dst.ptr) % 32 == 0;
  p.nchw = op.nchw; p.nCtot = op.nchw_ctot; p.nC0 = op.nchw_c0; p.nC = op.nchw_c;
  const int sms = sm_count();
  st->grid = p.total_tiles < sms ? p.total_tiles : sms;
  static std::atomic<unsigned long long> attr_devs{0};   // per DEVICE: the attribute does not carry over to another GPU
  if (first_on_device(attr_devs)) {
    cudaError_t e = cudaFuncSetAttribute(dwpw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) { delete st; set_error("dwpw: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return LY_E_CUDA; }
  }
  *out = st;
  return LY_OK;
}

int32_t dwpw_launch(const DwPwState* st, float* nchw_override, cudaStream_t s) {
  if (st->mma) return dwpw_mma_launch(st->mma, nchw_override, s);
  if (nchw_override) {
    FParams p = st->p;
    p.nchw = nchw_override;
    launch_k(dwpw_kernel, dim3(st->grid), dim3(kFThreads), st->smem, s, p);
  } else {
    launch_k(dwpw_kernel, dim3(st->grid), dim3(kFThreads), st->smem, s, st->p);
  }
  return post_launch("dwpw_tc");
}

void dwpw_free(DwPwState* st) {
  if (st && st->mma) dwpw_mma_free(st->mma);
  delete st;
}

}  // namespace ly
