// BN-folded NHWC bf16 implicit-GEMM convolution on Blackwell tensor cores.
//
//   D[M = B*Ho*Wo, N = Cout] = A[M, K = k*k*Cin] * W[N, K]^T  (+bias, SiLU, +residual)
//
// * A is never materialised: for every filter tap (r,s) and channel block the producer
//   issues ONE 4-D TMA tile load (C, W, H, B) of the NHWC activation at the shifted
//   coordinate; out-of-bounds (padding) elements are zero-filled by the TMA unit and a
//   stride-2 conv is the tensor map's elementStrides = 2.  The tile of 128 output pixels
//   is a (tw x th x tb) brick chosen per layer so that feature maps tile without waste
//   (e.g. 16x8x1 @160^2, 8x8x2 @40^2, 4x4x8 @20^2); 1x1/s1 layers collapse to a flat
//   [M, C] matrix.
// * MMA: tcgen05.mma.cta_group::1.kind::f16, M=128, N=BLOCK_N (<=256), K=16 per
//   instruction, bf16 x bf16 -> fp32 accumulators in TMEM (double-buffered so the
//   epilogue of tile i overlaps the MMAs of tile i+1).  Operands are K-major in shared
//   memory with the 32/64/128-byte swizzle that matches the channel block (16/32/64).
// * concat / split / residual are epilogue addressing: the output goes to a channel slice
//   (dC0, dCtot) of the consumer's concat buffer, the input tensor map starts at the
//   producer's channel offset, the shortcut is read in the epilogue.
// * Persistent: one CTA per SM, static round-robin over (pixel-brick, n-tile) tiles.
//   Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM alloc), 2..9 = epilogue.
//
// Reference semantics: leanyolo/models/yolov10/layers.py:51-88 (Conv = conv+BN+SiLU).
#include <cuda.h>
#include <string.h>
#include "common.cuh"
#include "tma.cuh"

namespace ly {

namespace {

constexpr int kThreads = 320;   // TMA warp + MMA warp + 8 epilogue warps
constexpr int kMaxCout = 1024;   // bias vector staged in shared memory
constexpr int kMaxStages = 12;
constexpr uint32_t kSmemBudget = 200 * 1024;

struct Params {
  CUtensorMap tmA;
  CUtensorMap tmB;
  // tiling
  int tw, th, tb;
  int tiles_w, tiles_h, tiles_b, tiles_n, total_tiles;
  int Wo, Ho, Bo;          // (possibly collapsed) output extents used for tiling / masking
  int hw_real;             // Ho*Wo of the real tensor (NCHW addressing)
  int k, stride, pad;
  int kc, kc_blocks, num_kb;
  int block_n, tmem_cols;
  int stages, a_stage, b_stage;   // bytes (stage strides)
  int b_box;                      // bytes one B TMA box delivers (<= b_stage)
  int b_resident;
  uint32_t idesc, desc_hi;
  int act;
  __nv_bfloat16* dst; int dCtot, dC0;
  const __nv_bfloat16* res; int rCtot, rC0;
  const float* bias;
  float* nchw; int nCtot, nC0, nC;
  int cin_pad;
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// SiLU(x) = x*sigmoid(x) = h + h*tanh(h), h = x/2: one MUFU op (tanh.approx, rel. err 2^-11)
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.b32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------------------- kernel
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A stages][B stages or resident B][barriers]
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + (uint32_t)p.stages * p.a_stage;
  const uint32_t b_bytes_total = p.b_resident ? (uint32_t)p.num_kb * p.b_stage : (uint32_t)p.stages * p.b_stage;
  const uint32_t bar_base = b_base + b_bytes_total;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
  const uint32_t bres_bar = bar_base + 8u * (2 * kMaxStages + 4);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ __align__(16) float s_bias[kMaxCout];
  for (int i = threadIdx.x; i < p.tiles_n * p.block_n; i += kThreads) s_bias[i] = p.bias[i];

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 8);
    }
    mbar_init(bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      if (p.b_resident) {
        mbar_expect_tx(bres_bar, (uint32_t)p.num_kb * p.b_box);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const int tap = kb / p.kc_blocks, cb = kb - tap * p.kc_blocks;
          tma_load_2d(b_base + kb * p.b_stage, &p.tmB, bres_bar, tap * p.cin_pad + cb * p.kc, 0);
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx = p.a_stage + (p.b_resident ? 0 : p.b_box);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int t = tile;
        const int nt = t % p.tiles_n; t /= p.tiles_n;
        const int wt = t % p.tiles_w; t /= p.tiles_w;
        const int ht = t % p.tiles_h;
        const int bt = t / p.tiles_h;
        const int w0 = wt * p.tw * p.stride - p.pad, h0 = ht * p.th * p.stride - p.pad, b0 = bt * p.tb;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          const int tap = kb / p.kc_blocks, cb = kb - tap * p.kc_blocks;
          const int r = tap / p.k, s = tap - r * p.k;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), tx);
          tma_load_4d(a_base + stage * p.a_stage, &p.tmA, full_bar(stage), cb * p.kc, w0 + s, h0 + r, b0);
          if (!p.b_resident)
            tma_load_2d(b_base + stage * p.b_stage, &p.tmB, full_bar(stage), tap * p.cin_pad + cb * p.kc, nt * p.block_n);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ================================
    if (elect_one()) {
      if (p.b_resident) {
        mbar_wait(bres_bar, 0);
        tc_fence_after();
      }
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const int ksteps = p.kc / 16;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.block_n);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t a_addr = a_base + stage * p.a_stage;
          const uint32_t b_addr = p.b_resident ? b_base + kb * p.b_stage : b_base + stage * p.b_stage;
          const uint64_t hi = (uint64_t)p.desc_hi << 32;
#pragma unroll 4
          for (int kk = 0; kk < ksteps; ++kk) {
            const uint64_t da = hi | (uint64_t)(((a_addr + kk * 32) >> 4) & 0x3FFFu) | (1ull << 16);
            const uint64_t db = hi | (uint64_t)(((b_addr + kk * 32) >> 4) & 0x3FFFu) | (1ull << 16);
            umma_bf16(d_tmem, da, db, p.idesc, (kb | kk) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tfull_bar(as));
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ============================== epilogue (8 warps) ========================
    // warp -> (TMEM lane quarter q = warp % 4, column half); a thread owns one pixel row and
    // walks its columns in 16-wide chunks, the TMEM load of chunk i+1 in flight while chunk i
    // is activated and stored.  The accumulator stage is released as soon as the last chunk
    // sits in registers, so the MMA warp can start tile i+2 while this tile is still stored.
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int nchunks = p.block_n >> 4;
    const int c_half = (nchunks + 1) >> 1;
    const int cbeg = half ? c_half : 0, cend = half ? nchunks : c_half;
    const int row = q * 32 + lane;
    const int dw = row % p.tw;
    const int dh = (row / p.tw) % p.th;
    const int db = row / (p.tw * p.th);
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int t = tile;
      const int nt = t % p.tiles_n; t /= p.tiles_n;
      const int wt = t % p.tiles_w; t /= p.tiles_w;
      const int ht = t % p.tiles_h;
      const int bt = t / p.tiles_h;
      const int w = wt * p.tw + dw, h = ht * p.th + dh, b = bt * p.tb + db;
      const bool valid = w < p.Wo && h < p.Ho && b < p.Bo;
      const long long lin = ((long long)b * p.Ho + h) * p.Wo + w;
      const int n0 = nt * p.block_n;
      mbar_wait(tfull_bar(as), aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.block_n);
      __nv_bfloat16* drow = p.dst ? p.dst + lin * p.dCtot + p.dC0 + n0 : nullptr;
      const __nv_bfloat16* rrow = p.res ? p.res + lin * p.rCtot + p.rC0 + n0 : nullptr;
      long long nchw_b = 0, nchw_rem = 0;
      if (p.nchw) { nchw_b = lin / p.hw_real; nchw_rem = lin - nchw_b * p.hw_real; }
      uint32_t nxt[16];
      if (cbeg < cend) tmem_ld16(taddr + cbeg * 16, nxt);
      for (int ch = cbeg; ch < cend; ++ch) {
        const int c = ch * 16;
        uint32_t r[16];
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = nxt[j];
        if (ch + 1 < cend) {
          tmem_ld16(taddr + c + 16, nxt);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(as));
        }
        if (valid) {
          float v[16];
          const float4* bp = reinterpret_cast<const float4*>(s_bias + n0 + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 bb = bp[j];
            v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + bb.x;
            v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb.y;
            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb.z;
            v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb.w;
          }
          if (p.act) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = silu_tanh(v[j]);
          }
          if (rrow) {
            float rv[16];
            load_vec<__nv_bfloat16>(rrow + c, rv);
            load_vec<__nv_bfloat16>(rrow + c + 8, rv + 8);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += rv[j];
          }
          if (drow) {
            store_vec<__nv_bfloat16>(drow + c, v);
            store_vec<__nv_bfloat16>(drow + c + 8, v + 8);
          }
          if (p.nchw) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = n0 + c + j;
              if (n < p.nC) p.nchw[(nchw_b * p.nCtot + p.nC0 + n) * (long long)p.hw_real + nchw_rem] = v[j];
            }
          }
        }
      }
      if (cbeg >= cend) {   // this warp has no columns (narrow N): still release the stage
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(as));
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
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
  return true;
}

int32_t conv_tc_prepare(const ly_op& op, ConvTcState** out) {
  LY_CHECK_ARG(conv_tc_supported(op), "conv_tc: unsupported op (bf16, k in {1,3}, stride in {1,2}, 16-channel granularity)");
  LY_CHECK_ARG(op.src.ptr && op.w && op.bias && (op.dst.ptr || op.nchw), "conv_tc: null pointer");
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
  p.hw_real = Ho * Wo;
  p.kc = Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 16);
  p.kc_blocks = Cin / p.kc;
  p.num_kb = op.k * op.k * p.kc_blocks;
  // N tile: the largest multiple of 16 that divides Cout and is <= 256
  int bn = 16;
  for (int c = 16; c <= 256 && c <= Cout; c += 16)
    if (Cout % c == 0) bn = c;
  p.block_n = bn;
  p.tiles_n = Cout / bn;
  p.tmem_cols = pow2_ge(2 * bn);

  // pixel brick
  const bool flat = (op.k == 1 && op.stride == 1);
  int dimW, dimH, dimB;
  if (flat) { dimW = op.B * Ho * Wo; dimH = 1; dimB = 1; } else { dimW = Wo; dimH = Ho; dimB = op.B; }
  int best_tw = 128, best_th = 1, best_tb = 1;
  {
    double best_cost = 1e30;
    for (int tw = 128; tw >= 1; tw >>= 1)
      for (int th = 128 / tw; th >= 1; th >>= 1) {
        const int tb = 128 / (tw * th);
        if (tw * op.stride > 256 || th * op.stride > 256 || tb > 256) continue;
        const double cover = (double)((dimW + tw - 1) / tw * tw) * ((dimH + th - 1) / th * th) * ((dimB + tb - 1) / tb * tb);
        // prefer wide bricks (longer contiguous runs) on ties, penalise batch-spanning bricks slightly
        const double cost = cover * (1.0 + 1e-3 * (tb > 1) + 1e-4 * (128 / tw));
        if (cost < best_cost) { best_cost = cost; best_tw = tw; best_th = th; best_tb = tb; }
      }
  }
  p.tw = best_tw; p.th = best_th; p.tb = best_tb;
  p.Wo = dimW; p.Ho = dimH; p.Bo = dimB;
  p.tiles_w = (dimW + p.tw - 1) / p.tw;
  p.tiles_h = (dimH + p.th - 1) / p.th;
  p.tiles_b = (dimB + p.tb - 1) / p.tb;
  const long long total = (long long)p.tiles_w * p.tiles_h * p.tiles_b * p.tiles_n;
  if (total > 0x7FFFFFFF) { delete st; set_error("conv_tc: too many tiles"); return LY_E_ARG; }
  p.total_tiles = (int)total;

  // shared-memory pipeline
  p.a_stage = 128 * p.kc * 2;
  p.b_box = bn * p.kc * 2;
  p.b_stage = (p.b_box + 1023) / 1024 * 1024;
  const long long b_all = (long long)p.num_kb * p.b_stage;
  static int resident_ok = -1;
  if (resident_ok < 0) { const char* e = getenv("LY_TC_B_RESIDENT"); resident_ok = e ? atoi(e) : 1; }
  p.b_resident = (resident_ok && p.tiles_n == 1 && b_all <= 96 * 1024) ? 1 : 0;
  const uint32_t bar_bytes = 8 * (2 * kMaxStages + 8);
  long long avail = (long long)kSmemBudget - bar_bytes - 1024 - (p.b_resident ? b_all : 0);
  int stages = (int)(avail / (p.a_stage + (p.b_resident ? 0 : p.b_stage)));
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) { delete st; set_error("conv_tc: tile does not fit in shared memory"); return LY_E_ARG; }
  p.stages = stages;
  st->smem = 1024 + (size_t)stages * p.a_stage + (p.b_resident ? (size_t)b_all : (size_t)stages * p.b_stage) + bar_bytes;
  if (st->smem < 120 * 1024) st->smem = 120 * 1024;  // force one CTA per SM (TMEM allocations must not contend)

  // descriptors
  const int swz = p.kc == 64 ? 2 : (p.kc == 32 ? 4 : 6);       // UMMA LayoutType: SW128 / SW64 / SW32
  const uint32_t sbo = (uint32_t)(8 * p.kc * 2) >> 4;          // 8-row group stride, 16-byte units
  p.desc_hi = (sbo & 0x3FFFu) | (1u << 14) /*version = 1 (sm_100)*/ | ((uint32_t)swz << 29);
  p.idesc = (1u << 4) /*D=f32*/ | (1u << 7) /*A=bf16*/ | (1u << 10) /*B=bf16*/ | ((uint32_t)(bn >> 3) << 17) | ((128u >> 4) << 24);

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
    box[0] = p.kc; box[1] = p.tw * op.stride; box[2] = p.th * op.stride; box[3] = p.tb;
    estr[0] = 1; estr[1] = op.stride; estr[2] = op.stride; estr[3] = 1;
    // a brick may not exceed the tensor extent in box units the driver rejects; clamp is not needed: OOB is legal
    CUresult r = encode(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        tswz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
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

  p.dst = (__nv_bfloat16*)op.dst.ptr; p.dCtot = op.dst.ctot; p.dC0 = op.dst.c0;
  p.res = (const __nv_bfloat16*)op.res.ptr; p.rCtot = op.res.ctot; p.rC0 = op.res.c0;
  p.bias = op.bias;
  p.nchw = op.nchw; p.nCtot = op.nchw_ctot; p.nC0 = op.nchw_c0; p.nC = op.nchw_c;

  const int sms = sm_count();
  st->grid = p.total_tiles < sms ? p.total_tiles : sms;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024 - (int)sizeof(float) * kMaxCout - 1024 /* static smem: s_bias */);
    if (e != cudaSuccess) { delete st; set_error("conv_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return LY_E_CUDA; }
    attr_set = true;
  }
  *out = st;
  return LY_OK;
}

int32_t conv_tc_launch(const ConvTcState* st, float* nchw_override, cudaStream_t s) {
  if (nchw_override) {
    Params p = st->p;
    p.nchw = nchw_override;
    conv_tc_kernel<<<st->grid, kThreads, st->smem, s>>>(p);
  } else {
    conv_tc_kernel<<<st->grid, kThreads, st->smem, s>>>(st->p);
  }
  return post_launch("conv_tc");
}

void conv_tc_free(ConvTcState* st) { delete st; }

}  // namespace ly
