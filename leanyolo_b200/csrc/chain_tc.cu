// LY_OP_CHAIN: a chain of dense Conv+BN(+SiLU) stages fused per spatial tile on the tensor cores.
//
// A whole C2f block (cv1 1x1 -> Bottleneck 3x3, 3x3 + shortcut -> cv2 1x1 over the concat;
// leanyolo/models/yolov10/layers.py:91-173) or the 3x3 -> 1x1 tail of a box-regression stack
// (head.py:86-92) runs as ONE launch: only the block's input and its output touch HBM.  Layer by
// layer those blocks are bound by the HBM round trips of their thin intermediates (C2f @160^2 of
// yolov10s: 1.4 ms for 0.26 ms of compulsory traffic), not by the tensor pipe.
//
// Geometry.  A tile is TW x TH output pixels; with h = number of 3x3 stages on the longest path its
// FRAME is (TH + 2h) rows of PW = TW + 2h pixels.  Every tensor of the tile is a shared-memory REGION
// [frame pixel][<= 64 channels] bf16, K-major with the 128/64/32-byte swizzle of its row width, frame
// pixels in row-major order.  A region that has gone through s 3x3 stages is stored SHIFTED: its row r
// holds frame pixel r + s*(PW+1).  With that convention
//     3x3 stage: output row r reads input row r + ky*PW + kx for tap (ky, kx)
//     1x1 stage: output row r reads input row r + (s_out - s_in)*(PW+1)
// so every (tap, K-block) of every stage is the SAME region viewed a few rows further down: an M = 128
// MMA takes 128 consecutive rows through an UMMA descriptor that simply starts at that row (the swizzle
// is a function of the absolute shared-memory address; conv_tc.cu's band mode relies on the same fact).
// Rows whose frame pixel lies outside the image are written as zeros (= the next conv's zero padding);
// rows that wrap around the frame edge hold garbage that no valid output ever reads.
//
// Roles (one CTA per SM, persistent over tiles):
//   warp 0      TMA: all stage weights once (resident), the input frame of every tile (OOB = zero fill)
//   warp 1      tcgen05.mma issuer: walks the (stage, M-tile) item list; an item waits for the M-tiles
//               of its producer stages that cover its rows (mbarriers written by the epilogue warps)
//   warps 2-17  epilogue: 4 groups x 4 warps (one per TMEM lane quarter); group g takes items i = g mod 4:
//               TMEM -> +bias, SiLU, (+shortcut from its region) -> bf16 -> swizzled region rows, or the
//               global NHWC slice / public NCHW fp32 tensor for the last stage
// Accumulators: 512 TMEM columns = 512/slot_w slots used round-robin by the items.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "common.cuh"
#include "tma.cuh"
#include "tc.cuh"

namespace ly {

namespace {

constexpr int kEpiWarps = 16;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxSt = LY_CHAIN_MAX_STAGES, kMaxBlk = LY_CHAIN_MAX_BLOCKS, kMaxReg = LY_CHAIN_MAX_REGIONS;
constexpr int kMaxMt = 8;          // M tiles per stage and tile
constexpr int kMaxItems = kMaxSt * kMaxMt;
constexpr int kMaxSlots = 8;
constexpr int kMaxEntries = 1024;  // MMAs per item list (descriptor table in shared memory)
constexpr int kNumBars = 3 + 2 * kMaxSlots + kMaxSt * kMaxMt + 1;   // + the TMEM base slot
constexpr uint32_t kSmemMax = 227 * 1024 - 1024;                    // dynamic shared memory incl. the 1 KB alignment slack

struct StageP {
  int k, act, cout, n_src, ksteps, n_mt;
  int src_region[kMaxBlk], src_a16[kMaxBlk], src_rows[kMaxBlk], src_prod[kMaxBlk];
  int dst_region, dst_c0, dst_shift;
  int res_region, res_c0, res_rows, res_prod;
  int reads_halo;                 // rows past the M tile a tap window reaches: 2*PW + 2 (3x3) or 0
  uint32_t w_off, w_tile;         // bytes from the smem base / bytes per (tap, block) weight tile
  uint32_t bias_off;              // floats into the bias block
  uint32_t idesc, b_hi;
};

struct Params {
  CUtensorMap tmX;
  CUtensorMap tmW[kMaxSt];
  StageP st[kMaxSt];
  const float* bias_g[kMaxSt];
  int n_stages, n_in, n_regions;
  uint32_t region_off[kMaxReg];
  int region_rowb[kMaxReg];
  uint32_t region_hi[kMaxReg];
  int TW, TH, PW, FH, halo;
  int tiles_x, tiles_y, total_tiles;
  uint32_t mg_pw;
  int H, W, B;
  int n_items, early;
  // item list of one iteration: (stage, M tile), flags bit0 = belongs to the NEXT tile of this CTA (an "early" stage
  // hoisted into the tail of the previous tile), bit1 = reads the input frame, bit2 = last reader of the input frame;
  // wstage/wmt: before issuing, also wait for M tiles 0..wmt of stage wstage of the CURRENT tile (0xFF: none)
  unsigned char item_stage[kMaxItems], item_mt[kMaxItems], item_flags[kMaxItems], item_wstage[kMaxItems], item_wmt[kMaxItems];
  unsigned short item_nmma[kMaxItems], item_first[kMaxItems];   // MMAs of the item / its first entry in the descriptor table
  unsigned char item_dep[kMaxItems][8];                         // per producer stage: last M tile the item needs (0xFF: none)
  uint32_t tab_off, hdr_off;                                    // byte offsets of the descriptor table / item headers
  int n_entries;
  int slot_w, n_slots, n_groups, wpg;   // epilogue: n_groups groups of wpg warps (wpg = 4: one warp per TMEM lane quarter; 8: + a column split)
  uint32_t ehdr_off;                    // per-item epilogue constants (3 x uint4), built in the prologue
  uint32_t x_bytes, w_bytes, bias_off_b, bar_off;
  __nv_bfloat16* dst; int dCtot, dC0, st256;
  float* nchw; int nCtot, nC0, nC;
  int prof;     // LY_CHAIN_PROF=1: CTA 0 prints where each role spent its cycles
};

__device__ __forceinline__ uint32_t lds32x4(uint32_t addr, uint32_t* v) {
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
  return 0;
}
__device__ __forceinline__ void sts32x4(uint32_t addr, const uint32_t* v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}

// The MMAs of one source block of one item: all taps x k-steps as straight-line code.  A descriptor is (constant high
// word | 14-bit address field): moving to the next tap / k-step is an integer add on the low word.  (A generic
// loop with run-time bounds spent ~130 cycles of dependent integer work per MMA in the single issuing thread.)
template <int TAPS, int KS>
__device__ __forceinline__ void issue_block(uint32_t d_tmem, uint32_t a_hi, uint32_t a0, uint32_t rowb16, uint32_t pwr16, uint32_t b_hi,
                                            uint32_t b0, uint32_t bstep, uint32_t idesc, uint32_t acc0) {
#pragma unroll
  for (int tap = 0; tap < TAPS; ++tap) {
    const uint32_t alo = a0 + (uint32_t)(tap / 3) * pwr16 + (uint32_t)(tap % 3) * rowb16;
    const uint32_t blo = b0 + (uint32_t)tap * bstep;
#pragma unroll
    for (int kk = 0; kk < KS; ++kk)
      umma_bf16(d_tmem, ((uint64_t)a_hi << 32) | (uint64_t)(alo + 2 * kk), ((uint64_t)b_hi << 32) | (uint64_t)(blo + 2 * kk), idesc,
                (tap | kk) != 0 ? 1u : acc0);
  }
}

__global__ void __launch_bounds__(kThreads, 1) chain_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* s_bias = reinterpret_cast<float*>(gen + p.bias_off_b);
  const uint32_t bar = base + p.bar_off;
  const uint32_t b_xfull = bar, b_xempty = bar + 8u, b_wfull = bar + 16u;
  auto tfull = [&](int s) { return bar + 8u * (3 + s); };
  auto tempty = [&](int s) { return bar + 8u * (3 + kMaxSlots + s); };
  auto done = [&](int st, int mt) { return bar + 8u * (3 + 2 * kMaxSlots + st * kMaxMt + mt); };
  const uint32_t tmem_slot = bar + 8u * (kNumBars - 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  for (int s = 0; s < p.n_stages; ++s) {
    const float sc = p.st[s].act ? 0.5f : 1.0f;     // SiLU(x) = h + h*tanh(h), h = x/2
    for (int i = threadIdx.x; i < p.st[s].cout; i += kThreads) s_bias[p.st[s].bias_off + i] = sc * p.bias_g[s][i];
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmX) : "memory");
    for (int s = 0; s < p.n_stages; ++s) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmW[s]) : "memory");
    mbar_init(b_xfull, 1);
    mbar_init(b_xempty, 1);
    mbar_init(b_wfull, 1);
    for (int s = 0; s < kMaxSlots; ++s) {
      mbar_init(tfull(s), 1);
      mbar_init(tempty(s), (uint32_t)p.wpg);
    }
    for (int s = 0; s < kMaxSt; ++s)
      for (int m = 0; m < kMaxMt; ++m) mbar_init(done(s, m), (uint32_t)p.wpg);
    fence_barrier_init();
    // the weights are parameters (never written by a kernel): request them before waiting for the producer
    mbar_expect_tx(b_wfull, p.w_bytes);
    for (int s = 0; s < p.n_stages; ++s) {
      const StageP& S = p.st[s];
      const int kc = S.ksteps * 16, ctot = kc * S.n_src;
      for (int tap = 0; tap < S.k * S.k; ++tap)
        for (int b = 0; b < S.n_src; ++b)
          tma_load_2d(base + S.w_off + (uint32_t)(tap * S.n_src + b) * S.w_tile, &p.tmW[s], b_wfull, tap * ctot + b * kc, 0);
    }
  }
  // Descriptor table: the (A, B) shared-memory descriptors of every MMA of the item list.  They do not depend on the
  // tile, so the single issuing thread only streams 16-byte entries instead of chasing parameter-bank loads
  // (measured: ~400 cycles of dependent constant loads per item and source block in the generic loop).
  uint4* tab = reinterpret_cast<uint4*>(gen + p.tab_off);
  uint4* hdr = reinterpret_cast<uint4*>(gen + p.hdr_off);
  uint4* ehdr = reinterpret_cast<uint4*>(gen + p.ehdr_off);
  for (int i = 0; i < p.n_items; ++i) {
    const StageP& S = p.st[p.item_stage[i]];
    const int mt = p.item_mt[i], taps = S.k * S.k, per_blk = taps * S.ksteps, n = p.item_nmma[i];
    for (int idx = threadIdx.x; idx < n; idx += kThreads) {
      const int b = idx / per_blk, rem = idx - b * per_blk, tap = rem / S.ksteps, kk = rem - tap * S.ksteps;
      const int R = S.src_region[b];
      const uint32_t rowb16 = (uint32_t)p.region_rowb[R] >> 4;
      const uint32_t trow = S.k == 3 ? (uint32_t)((tap / 3) * p.PW + (tap % 3)) : 0u;
      const uint32_t alo = (((base + p.region_off[R]) >> 4) + ((uint32_t)(mt * 128 + S.src_rows[b]) + trow) * rowb16 + (uint32_t)S.src_a16[b] + 2u * kk) | (1u << 16);
      const uint32_t blo = (((base + S.w_off + (uint32_t)(tap * S.n_src + b) * S.w_tile) >> 4) + 2u * kk) | (1u << 16);
      tab[p.item_first[i] + idx] = make_uint4(alo, p.region_hi[R], blo, S.b_hi);
    }
    if (threadIdx.x == 0) {
      const unsigned char* d = p.item_dep[i];
      hdr[2 * i] = make_uint4((uint32_t)p.item_stage[i] | ((uint32_t)mt << 8) | ((uint32_t)p.item_flags[i] << 16), (uint32_t)n | ((uint32_t)p.item_first[i] << 16),
                              S.idesc, 0u);
      hdr[2 * i + 1] = make_uint4((uint32_t)d[0] | ((uint32_t)d[1] << 8) | ((uint32_t)d[2] << 16) | ((uint32_t)d[3] << 24),
                                  (uint32_t)d[4] | ((uint32_t)d[5] << 8) | ((uint32_t)d[6] << 16) | ((uint32_t)d[7] << 24), 0u, 0u);
      // epilogue constants of the item: everything the epilogue warps would otherwise chase through the parameter bank
      const bool to_smem = S.dst_region >= 0, has_res = S.res_region >= 0;
      const uint32_t drb = to_smem ? (uint32_t)p.region_rowb[S.dst_region] : 0u, rrb = has_res ? (uint32_t)p.region_rowb[S.res_region] : 0u;
      int res_need = 0;
      if (has_res) res_need = min(p.st[S.res_prod].n_mt - 1, (mt * 128 + 127 + S.res_rows) >> 7);
      ehdr[3 * i] = make_uint4((uint32_t)p.item_flags[i] | ((uint32_t)mt << 8) | ((uint32_t)p.item_stage[i] << 16) | ((to_smem ? 1u : 0u) << 24) |
                                   ((has_res ? 1u : 0u) << 25) | ((S.act ? 1u : 0u) << 26),
                               to_smem ? base + p.region_off[S.dst_region] + (uint32_t)(mt * 128) * drb : 0u, drb | ((uint32_t)S.dst_c0 << 16),
                               (uint32_t)(S.dst_shift * (p.PW + 1) + mt * 128));
      ehdr[3 * i + 1] = make_uint4(has_res ? base + p.region_off[S.res_region] + (uint32_t)(mt * 128 + S.res_rows) * rrb : 0u, rrb | ((uint32_t)S.res_c0 << 16),
                                   (uint32_t)(mt * 128 + S.res_rows), (uint32_t)(has_res ? S.res_prod : 0) | ((uint32_t)res_need << 8));
      ehdr[3 * i + 2] = make_uint4((uint32_t)S.cout, S.bias_off, 0u, 0u);
    }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int tiles_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (elect_one()) {
      uint32_t itp = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int b = tile / tiles_img, t2 = tile - b * tiles_img;
        const int yt = t2 / p.tiles_x, xt = t2 - yt * p.tiles_x;
        mbar_wait(b_xempty, itp ^ 1u);
        mbar_expect_tx(b_xfull, p.x_bytes);
        for (int i = 0; i < p.n_in; ++i)
          tma_load_4d(base + p.region_off[i], &p.tmX, b_xfull, 64 * i, xt * p.TW - p.halo, yt * p.TH - p.halo, b);
        itp ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ================================
    if (elect_one()) {
      mbar_wait(b_wfull, 0);
      tc_fence_after();
      uint32_t gi = 0;
      long long t_x = 0, t_done = 0, t_slot = 0, t_issue = 0, t_start = clock64();
      const int n_tiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
      int x_tile = -1;                       // local tile whose input frame has been waited for
      for (int it = p.early ? -1 : 0; it < n_tiles; ++it) {
        int waited[kMaxSt];
#pragma unroll
        for (int s = 0; s < kMaxSt; ++s) waited[s] = 0;
        // wait until stage `prod` of local tile `it` has written its rows up to last_row
        auto need_rows = [&](int prod, int last_row) {
          if (prod < 0 || it < 0) return;
          int need = last_row >> 7;
          if (need > p.st[prod].n_mt - 1) need = p.st[prod].n_mt - 1;
          while (waited[prod] <= need) {
            mbar_wait(done(prod, waited[prod]), (uint32_t)it & 1u);
            ++waited[prod];
          }
        };
        for (int i = 0; i < p.n_items; ++i) {
          const uint4 h0 = hdr[2 * i], h1 = hdr[2 * i + 1];
          const int fl = (int)((h0.x >> 16) & 0xFFu);
          const int t = it + (fl & 1);
          if (t < 0 || t >= n_tiles) continue;
          long long t1 = clock64();
          if ((fl & 2) && x_tile != t) {
            mbar_wait(b_xfull, (uint32_t)t & 1u);
            x_tile = t;
            t_x += clock64() - t1;
            t1 = clock64();
          }
          if (it >= 0) {
#pragma unroll
            for (int s = 0; s < kMaxSt; ++s) {
              const uint32_t need = ((s < 4 ? h1.x : h1.y) >> (8 * (s & 3))) & 0xFFu;
              if (need != 0xFFu)
                while (waited[s] <= (int)need) {
                  mbar_wait(done(s, waited[s]), (uint32_t)it & 1u);
                  ++waited[s];
                }
            }
          }
          tc_fence_after();
          long long t2 = clock64();
          t_done += t2 - t1;
          const uint32_t slot = gi % (uint32_t)p.n_slots, use = gi / (uint32_t)p.n_slots;
          ++gi;
          mbar_wait(tempty(slot), (use & 1u) ^ 1u);
          tc_fence_after();
          long long t3 = clock64();
          t_slot += t3 - t2;
          const uint32_t d_tmem = tmem_base + slot * (uint32_t)p.slot_w;
          const uint32_t idesc = h0.z;
          const int n = (int)(h0.y & 0xFFFFu);
          const uint4* e = tab + (h0.y >> 16);
          int m = 0;
          for (; m + 4 <= n; m += 4) {
            const uint4 d0 = e[m], d1 = e[m + 1], d2 = e[m + 2], d3 = e[m + 3];
            umma_bf16(d_tmem, ((uint64_t)d0.y << 32) | d0.x, ((uint64_t)d0.w << 32) | d0.z, idesc, m != 0 ? 1u : 0u);
            umma_bf16(d_tmem, ((uint64_t)d1.y << 32) | d1.x, ((uint64_t)d1.w << 32) | d1.z, idesc, 1u);
            umma_bf16(d_tmem, ((uint64_t)d2.y << 32) | d2.x, ((uint64_t)d2.w << 32) | d2.z, idesc, 1u);
            umma_bf16(d_tmem, ((uint64_t)d3.y << 32) | d3.x, ((uint64_t)d3.w << 32) | d3.z, idesc, 1u);
          }
          for (; m < n; ++m) {
            const uint4 d0 = e[m];
            umma_bf16(d_tmem, ((uint64_t)d0.y << 32) | d0.x, ((uint64_t)d0.w << 32) | d0.z, idesc, m != 0 ? 1u : 0u);
          }
          umma_commit(tfull(slot));
          if (fl & 4) umma_commit(b_xempty);    // the input frame may be overwritten by the next tile's
          t_issue += clock64() - t3;
        }
        // every region row written for tile `it` has been consumed (or at least produced) before later tiles
        // overwrite it: wait for the M tiles nobody asked for, which also keeps the barrier phases in step
        for (int s = 0; s < p.n_stages; ++s)
          if (p.st[s].dst_region >= 0) need_rows(s, p.st[s].n_mt * 128 - 1);
      }
      if (p.prof && blockIdx.x == 0)
        printf("[chain prof] mma: tiles %d total %lld  wait_x %lld wait_done %lld wait_slot %lld issue %lld (cycles)\n", n_tiles,
               clock64() - t_start, t_x, t_done, t_slot, t_issue);
    }
  } else {
    // ============================== epilogue (n_groups groups of wpg warps) ==============
    const int q = warp & 3;
    const int ew = warp - 2;                                   // 0..15
    const int group = ew / p.wpg, sub = (ew % p.wpg) >> 2;      // sub: column share of the warp inside its group (wpg = 8: 0 / 1)
    const int nsub = p.wpg >> 2;
    const uint32_t row_in = (uint32_t)(q * 32 + lane);
    uint32_t gi = 0;
    long long e_wait = 0, e_work = 0, e_start = clock64();
    const int n_tiles = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t PW = (uint32_t)p.PW;
    for (int it = p.early ? -1 : 0; it < n_tiles; ++it) {
      for (int i = 0; i < p.n_items; ++i) {
        const uint4 e0 = ehdr[3 * i];
        const int t = it + (int)(e0.x & 1u);
        if (t < 0 || t >= n_tiles) continue;
        const uint32_t slot = gi % (uint32_t)p.n_slots, use = gi / (uint32_t)p.n_slots;
        // items go to the groups by their GLOBAL sequence number: n_slots is a multiple of the group count, so every
        // use of an accumulator slot is drained by the same group and its parity waits see every phase in turn
        const bool mine_item = (int)(gi % (uint32_t)p.n_groups) == group;
        ++gi;
        if (!mine_item) continue;
        const uint4 e1 = ehdr[3 * i + 1], e2 = ehdr[3 * i + 2];
        const uint32_t itp = (uint32_t)t & 1u;
        const int tile = (int)blockIdx.x + t * (int)gridDim.x;
        const int b = tile / tiles_img, t2 = tile - b * tiles_img;
        const int yt = t2 / p.tiles_x, xt = t2 - yt * p.tiles_x;
        const int y00 = yt * p.TH - p.halo, x00 = xt * p.TW - p.halo;
        const int s = (int)((e0.x >> 16) & 0xFFu), mt = (int)((e0.x >> 8) & 0xFFu);
        const bool to_smem = (e0.x >> 24) & 1u, has_res = (e0.x >> 25) & 1u, act = (e0.x >> 26) & 1u;
        const uint32_t qi = e0.w + row_in;                       // frame pixel index of this thread's row
        const uint32_t fy = __umulhi(qi, p.mg_pw), fx = qi - fy * PW;
        const int gy = y00 + (int)fy, gx = x00 + (int)fx;
        const bool in_img = (unsigned)gy < (unsigned)p.H && (unsigned)gx < (unsigned)p.W;
        // shared-memory destination / shortcut rows
        uint32_t d_row = 0, d_sw = 0, r_row = 0, r_sw = 0;
        const uint32_t drb = e0.z & 0xFFFFu, dc0 = e0.z >> 16, rrb = e1.y & 0xFFFFu, rc0 = e1.y >> 16;
        if (to_smem) {
          const uint32_t off = ((uint32_t)mt * 128u + row_in) * drb;
          d_row = e0.y + row_in * drb;
          d_sw = (off >> 7) & ((drb >> 4) - 1u);
        }
        if (has_res) {
          const uint32_t rr = e1.z + row_in;
          r_row = e1.x + row_in * rrb;
          r_sw = ((rr * rrb) >> 7) & ((rrb >> 4) - 1u);
          // the shortcut region was written by epilogue warps of an earlier stage: make sure those rows are there
          const int rp = (int)(e1.w & 0xFFu), need = (int)(e1.w >> 8);
          for (int m = 0; m <= need; ++m) mbar_wait(done(rp, m), itp);
        }
        // global destination (last stage): only the tile's own pixels
        __nv_bfloat16* drow = nullptr;
        float* nrow = nullptr;
        if (!to_smem) {
          const bool mine = in_img && fy >= (uint32_t)p.halo && fy < (uint32_t)(p.halo + p.TH) && fx >= (uint32_t)p.halo &&
                            fx < (uint32_t)(p.halo + p.TW);
          if (mine) {
            const uint32_t lin = ((uint32_t)b * (uint32_t)p.H + (uint32_t)gy) * (uint32_t)p.W + (uint32_t)gx;
            if (p.nchw) nrow = p.nchw + ((size_t)b * p.nCtot + p.nC0) * ((size_t)p.H * p.W) + (size_t)gy * p.W + gx;
            else drow = p.dst + (size_t)lin * (uint32_t)p.dCtot + p.dC0;
          }
        }
        const float pre = act ? 0.5f : 1.0f;
        const float* bias = s_bias + e2.y;
        const int nch = (int)e2.x >> 4;
        // column share of this warp: chunks [c_lo, c_hi)
        const int per = (nch + nsub - 1) / nsub;
        const int c_lo = min(sub * per, nch), c_hi = min(c_lo + per, nch);
        const long long w0 = clock64();
        mbar_wait(tfull(slot), use & 1u);
        const long long w1 = clock64();
        e_wait += w1 - w0;
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + slot * (uint32_t)p.slot_w;
        uint32_t nxt[16];
        if (c_lo < c_hi) tmem_ld16(taddr + 16 * c_lo, nxt);
        else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(slot));
        }
        for (int ch = c_lo; ch < c_hi; ++ch) {
          const int c = ch * 16;
          float v[16];
          tmem_ld_wait();
          const float4* bp = reinterpret_cast<const float4*>(bias + c);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 bb = bp[j];
            ffma2(v[4 * j + 0], v[4 * j + 1], __uint_as_float(nxt[4 * j + 0]), __uint_as_float(nxt[4 * j + 1]), pre, pre, bb.x, bb.y);
            ffma2(v[4 * j + 2], v[4 * j + 3], __uint_as_float(nxt[4 * j + 2]), __uint_as_float(nxt[4 * j + 3]), pre, pre, bb.z, bb.w);
          }
          if (ch + 1 < c_hi) {
            tmem_ld16(taddr + c + 16, nxt);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty(slot));
          }
          if (act) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) silu2_from_half(v[j], v[j + 1]);
          }
          if (has_res) {
            uint32_t rw[8];
            const uint32_t j0 = (rc0 + (uint32_t)c) >> 3;
            lds32x4(r_row + ((j0 ^ r_sw) << 4), rw);
            lds32x4(r_row + (((j0 + 1u) ^ r_sw) << 4), rw + 4);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[2 * j] += __uint_as_float(rw[j] << 16);
              v[2 * j + 1] += __uint_as_float(rw[j] & 0xffff0000u);
            }
          }
          if (to_smem) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
              w[j] = in_img ? *reinterpret_cast<const uint32_t*>(&h) : 0u;
            }
            const uint32_t j0 = (dc0 + (uint32_t)c) >> 3;
            sts32x4(d_row + ((j0 ^ d_sw) << 4), w);
            sts32x4(d_row + (((j0 + 1u) ^ d_sw) << 4), w + 4);
          } else if (drow) {
            if (p.st256) {
              store_bf16x16(drow + c, v);
            } else {
              store_vec<__nv_bfloat16>(drow + c, v);
              store_vec<__nv_bfloat16>(drow + c + 8, v + 8);
            }
          } else if (nrow) {
            float* np = nrow + (size_t)c * ((size_t)p.H * p.W);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c + j < p.nC) np[(size_t)j * ((size_t)p.H * p.W)] = v[j];
          }
        }
        if (to_smem) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
          __syncwarp();
          if (lane == 0) mbar_arrive(done(s, mt));
        }
        e_work += clock64() - w1;
      }
    }
    if (p.prof && blockIdx.x == 0 && lane == 0 && q == 0)
      printf("[chain prof] epilogue group %d: total %lld wait_tfull %lld work %lld\n", group, clock64() - e_start, e_wait, e_work);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

int env_i(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

double mma_cycles(int n) {   // measured issue cost of an M=128, K=16 tcgen05.mma (tools/mma_bench.cu)
  if (n <= 32) return 48.0;
  if (n <= 64) return 52.0;
  if (n <= 128) return 52.0 + (n - 64) * (12.4 / 64.0);
  return 64.4 + (n - 128) * (63.3 / 128.0);
}

struct Layout {
  int TW, TH, PW, FH;
  int n_mt[kMaxSt];
  int rows[kMaxReg];
  uint32_t region_off[kMaxReg];
  uint32_t w_off[kMaxSt], w_tile[kMaxSt];
  uint32_t bias_off_b, bar_off, tab_off, hdr_off, ehdr_off, total;
  int items;
  double cost;
};

}  // namespace

struct ChainState {
  B2bState* b2b = nullptr;   // set: the chain is a 3x3 -> 1x1 tail handled by conv_b2b.cu (the fields below are unused)
  Params p;
  ly_chain chain;   // own copy (weights / bias pointers are device pointers that stay valid)
  int grid;
  size_t smem;
};

bool chain_tc_supported(const ly_op& op) {
  return op.kind == LY_OP_CHAIN && op.dtype == LY_BF16 && op.chain != nullptr;
}

int32_t chain_tc_prepare(const ly_op& op, ChainState** out) {
  LY_CHECK_ARG(chain_tc_supported(op), "chain: needs a bf16 op with a chain description");
  if (conv_b2b_supported(op)) {
    B2bState* b = nullptr;
    if (conv_b2b_prepare(op, &b) == LY_OK) {     // (does not fit in shared memory: fall through to the generic kernel)
      ChainState* st = new ChainState();
      st->b2b = b;
      *out = st;
      return LY_OK;
    }
  }
  const ly_chain& ch = *op.chain;
  LY_CHECK_ARG(ch.n_stages >= 1 && ch.n_stages <= kMaxSt && ch.n_regions >= 1 && ch.n_regions <= kMaxReg && ch.n_in >= 1 &&
                   ch.n_in <= ch.n_regions, "chain: bad stage / region counts");
  LY_CHECK_ARG(op.src.ptr && (op.dst.ptr || op.nchw) && !(op.dst.ptr && op.nchw), "chain: needs a source and exactly one destination");
  LY_CHECK_ARG(op.src.c0 % 8 == 0 && op.src.ctot % 8 == 0, "chain: source slice must be 16-byte aligned");
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("chain: cuTensorMapEncodeTiled entry point not available"); return LY_E_CUDA; }
  const int H = op.src.H, W = op.src.W;

  // ---- static analysis of the stage program: shifts, producers, validity
  int shift[kMaxReg], prod[kMaxReg], rowb[kMaxReg];
  int cin_total = 0;
  for (int r = 0; r < ch.n_regions; ++r) {
    const int c = ch.region_c[r];
    LY_CHECK_ARG(c == 16 || c == 32 || c == 64, "chain: region %d has %d channels (16, 32 or 64)", r, c);
    rowb[r] = 2 * c; shift[r] = 0; prod[r] = -1;
    if (r < ch.n_in) {
      LY_CHECK_ARG(c == ch.region_c[0] && (ch.n_in == 1 || c == 64), "chain: input regions must be 64 channels wide (or a single narrower one)");
      cin_total += c;
    }
  }
  LY_CHECK_ARG(cin_total == op.src.c, "chain: the input regions hold %d channels, the source view %d", cin_total, op.src.c);

  ChainState* st = new ChainState();
  st->chain = ch;
  Params& p = st->p;
  memset(&p, 0, sizeof(p));
  p.n_stages = ch.n_stages; p.n_in = ch.n_in; p.n_regions = ch.n_regions;
  p.H = H; p.W = W; p.B = op.B;
  bool written[kMaxReg] = {false};
  int x_last_stage = 0, bias_total = 0, max_cout = 16;
  int so[kMaxSt], src_sh[kMaxSt][kMaxBlk], res_sh[kMaxSt];
  auto fail = [&](const char* msg, int s) { set_error("chain: stage %d: %s", s, msg); delete st; return LY_E_ARG; };
  for (int s = 0; s < ch.n_stages; ++s) {
    const ly_chain_stage& cs = ch.st[s];
    StageP& S = p.st[s];
    if (!(cs.k == 1 || cs.k == 3)) return fail("k must be 1 or 3", s);
    if (cs.cout < 16 || cs.cout > 256 || cs.cout % 16) return fail("cout must be a multiple of 16 in 16..256", s);
    if (cs.n_src < 1 || cs.n_src > kMaxBlk || !cs.w || !cs.bias) return fail("bad sources / null weights", s);
    const int kc = cs.src[0].c;
    if (!(kc == 16 || kc == 32 || kc == 64)) return fail("source blocks must be 16, 32 or 64 channels wide", s);
    int smax = 0;
    for (int b = 0; b < cs.n_src; ++b) {
      const ly_chain_blk& blk = cs.src[b];
      if (blk.region < 0 || blk.region >= ch.n_regions || blk.c != kc || blk.c0 % 16 || blk.c0 + blk.c > ch.region_c[blk.region])
        return fail("bad source block", s);
      if (blk.region >= ch.n_in && !written[blk.region]) return fail("source region is read before it is written", s);
      smax = std::max(smax, shift[blk.region]);
      if (blk.region < ch.n_in) x_last_stage = s;
    }
    so[s] = smax + (cs.k == 3 ? 1 : 0);
    S.k = cs.k; S.act = cs.act; S.cout = cs.cout; S.n_src = cs.n_src; S.ksteps = kc / 16;
    for (int b = 0; b < cs.n_src; ++b) {
      const ly_chain_blk& blk = cs.src[b];
      S.src_region[b] = blk.region; S.src_a16[b] = blk.c0 * 2 / 16; S.src_prod[b] = prod[blk.region];
      src_sh[s][b] = (cs.k == 3 ? so[s] - 1 : so[s]) - shift[blk.region];
    }
    S.res_region = -1; S.res_prod = -1; res_sh[s] = 0;
    if (cs.res.region >= 0) {
      const ly_chain_blk& blk = cs.res;
      if (blk.region >= ch.n_regions || blk.c != cs.cout || blk.c0 % 16 || blk.c0 + blk.c > ch.region_c[blk.region]) return fail("bad shortcut block", s);
      if (blk.region < ch.n_in || !written[blk.region]) return fail("the shortcut must be a region written by an earlier stage", s);
      if (shift[blk.region] > so[s]) return fail("shortcut region is deeper than the stage output", s);
      S.res_region = blk.region; S.res_c0 = blk.c0; S.res_prod = prod[blk.region];
      res_sh[s] = so[s] - shift[blk.region];
    }
    S.dst_region = -1; S.dst_shift = so[s];
    if (cs.dst.region >= 0) {
      const ly_chain_blk& blk = cs.dst;
      if (s == ch.n_stages - 1) return fail("the last stage writes the global destination", s);
      if (blk.region < ch.n_in || blk.region >= ch.n_regions || blk.c != cs.cout || blk.c0 % 16 || blk.c0 + blk.c > ch.region_c[blk.region])
        return fail("bad destination block", s);
      for (int b = 0; b < cs.n_src; ++b)
        if (cs.src[b].region == blk.region && !(cs.k == 3 && cs.n_src == 1 && cs.src[b].c0 == blk.c0 && cs.src[b].c == blk.c))
          return fail("in-place output is only safe for a single-source 3x3 stage over the same channels", s);
      if (cs.res.region == blk.region) return fail("destination aliases the shortcut", s);
      S.dst_region = blk.region; S.dst_c0 = blk.c0;
      shift[blk.region] = so[s]; prod[blk.region] = s; written[blk.region] = true;
    } else if (s != ch.n_stages - 1) {
      return fail("only the last stage may write the global destination", s);
    }
    S.bias_off = (uint32_t)bias_total;
    bias_total += cs.cout;
    max_cout = std::max(max_cout, cs.cout);
    p.bias_g[s] = cs.bias;
    const int swz = kc == 64 ? 2 : (kc == 32 ? 4 : 6);
    S.b_hi = (((uint32_t)(8 * kc * 2) >> 4) & 0x3FFFu) | (1u << 14) | ((uint32_t)swz << 29);
    S.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(cs.cout >> 3) << 17) | ((128u >> 4) << 24);
  }
  const int halo = so[ch.n_stages - 1];
  const int cout_final = ch.st[ch.n_stages - 1].cout;
  if (op.dst.ptr) {
    if (op.dst.c != cout_final || op.dst.c0 % 8 || op.dst.ctot % 8 || op.dst.H != H || op.dst.W != W) { delete st; set_error("chain: destination view does not match the last stage"); return LY_E_ARG; }
  } else if (op.nchw_c > cout_final || op.nchw_c < 1) { delete st; set_error("chain: nchw_c does not match the last stage"); return LY_E_ARG; }
  p.halo = halo;
  p.slot_w = max_cout <= 64 ? 64 : (max_cout <= 128 ? 128 : 256);
  p.n_slots = std::min(kMaxSlots, 512 / p.slot_w);
  p.wpg = env_i("LY_CHAIN_WPG", 8) == 4 ? 4 : 8;
  p.n_groups = std::min(kEpiWarps / p.wpg, p.n_slots);
  for (int r = 0; r < ch.n_regions; ++r) {
    p.region_rowb[r] = rowb[r];
    const int swz = rowb[r] == 128 ? 2 : (rowb[r] == 64 ? 4 : 6);
    p.region_hi[r] = (((uint32_t)(8 * rowb[r]) >> 4) & 0x3FFFu) | (1u << 14) | ((uint32_t)swz << 29);
  }

  // ---- tile size: minimise modelled MMA issue cycles per image under the shared-memory budget
  auto layout = [&](int TW, int TH, Layout& L) -> bool {
    L.TW = TW; L.TH = TH; L.PW = TW + 2 * halo; L.FH = TH + 2 * halo;
    if (L.PW > 256 || L.FH > 256) return false;
    int wr[kMaxReg], rd[kMaxReg];
    for (int r = 0; r < ch.n_regions; ++r) { wr[r] = r < ch.n_in ? L.FH * L.PW : 0; rd[r] = 0; }
    L.items = 0; L.cost = 0;
    for (int s = 0; s < ch.n_stages; ++s) {
      const ly_chain_stage& cs = ch.st[s];
      const int rows = (L.FH - 2 * so[s]) * L.PW - 2 * so[s];
      if (rows <= 0) return false;
      const int mt = (rows + 127) / 128;
      if (mt > kMaxMt) return false;
      L.n_mt[s] = mt; L.items += mt;
      const int halo_rows = cs.k == 3 ? 2 * L.PW + 2 : 0;
      for (int b = 0; b < cs.n_src; ++b) rd[cs.src[b].region] = std::max(rd[cs.src[b].region], mt * 128 + src_sh[s][b] * (L.PW + 1) + halo_rows);
      if (cs.res.region >= 0) rd[cs.res.region] = std::max(rd[cs.res.region], mt * 128 + res_sh[s] * (L.PW + 1));
      if (cs.dst.region >= 0) wr[cs.dst.region] = std::max(wr[cs.dst.region], mt * 128);
      L.cost += (double)mt * cs.k * cs.k * cs.n_src * (cs.src[0].c / 16) * mma_cycles(cs.cout);
    }
    if (L.items > kMaxItems) return false;
    uint32_t off = 0;
    for (int r = 0; r < ch.n_regions; ++r) {
      L.rows[r] = std::max(wr[r], rd[r]);
      L.region_off[r] = off;
      off += ((uint32_t)L.rows[r] * (uint32_t)rowb[r] + 1023u) / 1024u * 1024u;
    }
    for (int s = 0; s < ch.n_stages; ++s) {
      const ly_chain_stage& cs = ch.st[s];
      L.w_tile[s] = ((uint32_t)cs.cout * cs.src[0].c * 2 + 1023u) / 1024u * 1024u;
      L.w_off[s] = off;
      off += L.w_tile[s] * (uint32_t)(cs.k * cs.k * cs.n_src);
    }
    L.bias_off_b = off; off += ((uint32_t)bias_total * 4 + 15u) / 16u * 16u;
    L.bar_off = off; off += (8u * kNumBars + 15u) / 16u * 16u;
    int entries = 0;
    for (int s = 0; s < ch.n_stages; ++s) entries += L.n_mt[s] * ch.st[s].k * ch.st[s].k * ch.st[s].n_src * (ch.st[s].src[0].c / 16);
    if (entries > kMaxEntries) return false;
    L.tab_off = off; off += 16u * (uint32_t)entries;
    L.hdr_off = off; off += 32u * (uint32_t)L.items;
    L.ehdr_off = off; off += 48u * (uint32_t)L.items;
    L.total = off + 1024u;
    if (L.total > kSmemMax) return false;
    const long long tiles = (long long)((W + TW - 1) / TW) * ((H + TH - 1) / TH);
    L.cost = (L.cost + 1200.0) * (double)tiles;
    return true;
  };
  Layout best; best.cost = 1e300; bool found = false;
  const int force_tw = env_i("LY_CHAIN_TW", 0), force_th = env_i("LY_CHAIN_TH", 0);
  for (int TW = 2; TW <= std::min(W + 1, 128); TW += 2)
    for (int TH = 1; TH <= std::min(H, 128); ++TH) {
      if (force_tw && (TW != force_tw || TH != force_th)) continue;
      Layout L;
      if (layout(TW, TH, L) && L.cost < best.cost) { best = L; found = true; }
    }
  if (!found) { delete st; set_error("chain: no tile fits in shared memory"); return LY_E_ARG; }
  const Layout& L = best;
  p.TW = L.TW; p.TH = L.TH; p.PW = L.PW; p.FH = L.FH;
  p.tiles_x = (W + L.TW - 1) / L.TW; p.tiles_y = (H + L.TH - 1) / L.TH;
  const long long total = (long long)p.tiles_x * p.tiles_y * op.B;
  if (total > 0x7FFFFFFF) { delete st; set_error("chain: too many tiles"); return LY_E_ARG; }
  p.total_tiles = (int)total;
  p.mg_pw = (uint32_t)((1ull << 32) / (uint32_t)L.PW + 1);
  for (int s = 0; s < ch.n_stages; ++s) {
    StageP& S = p.st[s];
    S.n_mt = L.n_mt[s];
    S.reads_halo = S.k == 3 ? 2 * L.PW + 2 : 0;
    for (int b = 0; b < S.n_src; ++b) S.src_rows[b] = src_sh[s][b] * (L.PW + 1);
    S.res_rows = res_sh[s] * (L.PW + 1);
    S.w_off = L.w_off[s]; S.w_tile = L.w_tile[s];
    p.w_bytes += (uint32_t)(S.k * S.k * S.n_src) * (uint32_t)(S.cout * S.ksteps * 16 * 2);
  }
  // ---- item schedule.  A stage that reads only the input frame ("early") does not depend on anything of its own
  // tile, so its M tiles are hoisted into the tail of the PREVIOUS tile's item list: each one right after the last
  // item of that tile whose MMAs read the rows it is going to overwrite.  The stage boundary bubble (issuer waiting
  // for the epilogue of the first stage) then overlaps the previous tile's last stage.
  bool is_early[kMaxSt];
  bool reads_in[kMaxSt];
  int n_early = 0;
  for (int s = 0; s < ch.n_stages; ++s) {
    reads_in[s] = false; is_early[s] = ch.st[s].dst.region >= 0 && ch.st[s].res.region < 0;
    for (int b = 0; b < ch.st[s].n_src; ++b) {
      if (ch.st[s].src[b].region < ch.n_in) reads_in[s] = true; else is_early[s] = false;
    }
    n_early += is_early[s];
  }
  bool early_mode = env_i("LY_CHAIN_EARLY", 1) != 0 && n_early == 1 && is_early[0];
  for (int s = 1; s < ch.n_stages; ++s)
    if (reads_in[s]) early_mode = false;      // a later stage still needs the input frame of its own tile
  struct Item { int s, m, fl, ws, wm; };
  Item items[kMaxItems];
  int n_it = 0;
  for (int s = early_mode ? 1 : 0; s < ch.n_stages; ++s)
    for (int m = 0; m < p.st[s].n_mt; ++m) items[n_it++] = Item{s, m, reads_in[s] ? 2 : 0, 0xFF, 0};
  if (early_mode) {
    const StageP& E = p.st[0];
    int prev_pos = -1, res_readers = 0;
    for (int c = 1; c < ch.n_stages; ++c) res_readers += (p.st[c].res_region == E.dst_region && p.st[c].res_prod == 0);
    if (res_readers > 1) early_mode = false;
    for (int j = 0; early_mode && j < E.n_mt; ++j) {
      const int lo = j * 128, hi = j * 128 + 127;
      int pos = prev_pos, ws = 0xFF, wm = 0;
      for (int q = 0; q < n_it; ++q) {
        if (items[q].fl & 1) continue;
        const StageP& C = p.st[items[q].s];
        for (int b = 0; b < C.n_src; ++b) {
          if (C.src_region[b] != E.dst_region || C.src_prod[b] != 0) continue;
          const int rlo = items[q].m * 128 + C.src_rows[b], rhi = rlo + 127 + C.reads_halo;
          if (rlo <= hi && rhi >= lo) pos = std::max(pos, q);
        }
        if (C.res_region == E.dst_region && C.res_prod == 0) {
          const int rlo = items[q].m * 128 + C.res_rows, rhi = rlo + 127;
          if (rlo <= hi && rhi >= lo) { ws = items[q].s; wm = std::max(wm, items[q].m); }
        }
      }
      // insert after position `pos`
      for (int q = n_it; q > pos + 1; --q) items[q] = items[q - 1];
      items[pos + 1] = Item{0, j, 1 | 2, ws, wm};
      ++n_it;
      prev_pos = pos + 1;
    }
  }
  if (!early_mode) {   // (re)build the plain stage-major list
    n_it = 0;
    for (int s = 0; s < ch.n_stages; ++s)
      for (int m = 0; m < p.st[s].n_mt; ++m) items[n_it++] = Item{s, m, reads_in[s] ? 2 : 0, 0xFF, 0};
  }
  int last_x = -1;
  for (int q = 0; q < n_it; ++q)
    if (items[q].fl & 2) last_x = q;
  items[last_x].fl |= 4;
  p.n_items = n_it; p.early = early_mode ? 1 : 0;
  p.tab_off = L.tab_off; p.hdr_off = L.hdr_off; p.ehdr_off = L.ehdr_off; p.n_entries = 0;
  for (int q = 0; q < n_it; ++q) {
    const StageP& S = p.st[items[q].s];
    p.item_stage[q] = (unsigned char)items[q].s; p.item_mt[q] = (unsigned char)items[q].m; p.item_flags[q] = (unsigned char)items[q].fl;
    p.item_wstage[q] = (unsigned char)items[q].ws; p.item_wmt[q] = (unsigned char)items[q].wm;
    p.item_nmma[q] = (unsigned short)(S.k * S.k * S.n_src * S.ksteps);
    p.item_first[q] = (unsigned short)p.n_entries;
    p.n_entries += p.item_nmma[q];
    // producer M tiles this item has to wait for (tile `it` of the issuer's loop): sources and shortcut of a late
    // item; for an early item the shortcut readers of the region it overwrites
    unsigned char* dep = p.item_dep[q];
    for (int s = 0; s < 8; ++s) dep[s] = 0xFF;
    auto need = [&](int prod, int last_row) {
      if (prod < 0) return;
      const int m = std::min(p.st[prod].n_mt - 1, last_row >> 7);
      dep[prod] = dep[prod] == 0xFF ? (unsigned char)m : (unsigned char)std::max<int>(dep[prod], m);
    };
    const int last = items[q].m * 128 + 127;
    if (!(items[q].fl & 1)) {
      for (int b = 0; b < S.n_src; ++b) need(S.src_prod[b], last + S.reads_halo + S.src_rows[b]);
      need(S.res_prod, last + S.res_rows);
    }
    if (items[q].ws != 0xFF) need(items[q].ws, items[q].wm * 128 + 127);
  }
  for (int r = 0; r < ch.n_regions; ++r) p.region_off[r] = L.region_off[r];
  p.bias_off_b = L.bias_off_b; p.bar_off = L.bar_off;
  const int kc_in = ch.region_c[0];
  p.x_bytes = (uint32_t)ch.n_in * (uint32_t)(L.FH * L.PW * kc_in * 2);
  st->smem = std::max<size_t>(L.total, 120 * 1024);   // one CTA per SM (the TMEM allocation takes all 512 columns)

  // ---- tensor maps
  {
    const CUtensorMapSwizzle tswz = kc_in == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc_in == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const CUtensorMapL2promotion promo = kc_in == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                        : (kc_in == 32 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_NONE);
    char* gbase = (char*)op.src.ptr + (size_t)op.src.c0 * 2;
    cuuint64_t dims[4] = {(cuuint64_t)op.src.c, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)op.B};
    cuuint64_t strides[3] = {(cuuint64_t)op.src.ctot * 2, (cuuint64_t)op.src.ctot * 2 * W, (cuuint64_t)op.src.ctot * 2 * W * H};
    cuuint32_t box[4] = {(cuuint32_t)kc_in, (cuuint32_t)L.PW, (cuuint32_t)L.FH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&p.tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, gbase, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, tswz,
                        promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete st; set_error("chain: cuTensorMapEncodeTiled(input) failed with %d", (int)r); return LY_E_CUDA; }
  }
  for (int s = 0; s < ch.n_stages; ++s) {
    const ly_chain_stage& cs = ch.st[s];
    const int kc = cs.src[0].c;
    const CUtensorMapSwizzle tswz = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    cuuint64_t dims[2] = {(cuuint64_t)cs.k * cs.k * kc * cs.n_src, (cuuint64_t)cs.cout};
    cuuint64_t strides[1] = {dims[0] * 2};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)cs.cout};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&p.tmW[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)cs.w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        tswz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { delete st; set_error("chain: cuTensorMapEncodeTiled(weights %d) failed with %d", s, (int)r); return LY_E_CUDA; }
  }
  p.dst = (__nv_bfloat16*)op.dst.ptr; p.dCtot = op.dst.ctot; p.dC0 = op.dst.c0;
  p.st256 = op.dst.ptr && op.dst.ctot % 16 == 0 && op.dst.c0 % 16 == 0 && reinterpret_cast<uintptr_t>(op.dst.ptr) % 32 == 0;
  p.nchw = op.nchw; p.nCtot = op.nchw_ctot; p.nC0 = op.nchw_c0; p.nC = op.nchw_c;
  st->grid = std::min(p.total_tiles, sm_count());
  p.prof = env_i("LY_CHAIN_PROF", 0);
  if (env_i("LY_CHAIN_DEBUG", 0))
    fprintf(stderr, "[chain] %dx%d B %d stages %d halo %d: tile %dx%d frame %dx%d tiles %d items %d smem %zu slots %d x %d\n", H, W, op.B,
            ch.n_stages, halo, p.TW, p.TH, p.PW, p.FH, p.total_tiles, p.n_items, st->smem, p.n_slots, p.slot_w);
  if (env_i("LY_CHAIN_DEBUG", 0) > 1) {
    fprintf(stderr, "[chain] items (early %d):", p.early);
    for (int q = 0; q < p.n_items; ++q) fprintf(stderr, " %s%d.%d", (p.item_flags[q] & 1) ? "+" : "", p.item_stage[q], p.item_mt[q]);
    fprintf(stderr, "\n");
  }
  cudaError_t e = cudaFuncSetAttribute(chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
  if (e != cudaSuccess) { delete st; set_error("chain: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return LY_E_CUDA; }
  *out = st;
  return LY_OK;
}

int32_t chain_tc_launch(const ChainState* st, float* nchw_override, cudaStream_t s) {
  if (st->b2b) return conv_b2b_launch(st->b2b, nchw_override, s);
  if (nchw_override) {
    Params p = st->p;
    p.nchw = nchw_override;
    launch_k(chain_tc_kernel, dim3(st->grid), dim3(kThreads), st->smem, s, p);
  } else {
    launch_k(chain_tc_kernel, dim3(st->grid), dim3(kThreads), st->smem, s, st->p);
  }
  return post_launch("chain_tc");
}

void chain_tc_free(ChainState* st) {
  if (st && st->b2b) conv_b2b_free(st->b2b);
  delete st;
}

}  // namespace ly
