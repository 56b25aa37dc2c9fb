// GPU-resident detection tail.
//
//  dfl_kernel    DFL softmax-expectation + anchor decode + sigmoid, one thread per anchor.
//                (postprocess.py:217-235 / 103-131, utils/tal.py:10-46)  -- HBM-bound: reads
//                (4*reg_max+nc)*4 B per anchor once, coalesced along the anchor axis.
//  topk_kernel   two-stage top-k (postprocess.py:243-261), one CTA per image: 64-bit
//                radix-select + bitonic sort in shared memory.  Keys are
//                (score bits << 32 | ~index): all distinct, so the canonical tie rule
//                (score desc, index asc) is exact and the result is deterministic.
//  nms_kernel    greedy IoU NMS (box_ops.py:49-78), one CTA per image: candidate compaction,
//                bitonic sort by (score desc, index asc), then chunks of 512 candidates:
//                (A) suppression by the already-kept set, (B) 512x512 IoU bitmask,
//                (C) warp-ballot sequential scan.  Stops at max_keep survivors -- identical
//                to the reference's run-to-exhaustion + [:max_det] (later boxes never affect
//                earlier decisions).  IoU uses the reference's exact fp32 operation order
//                with explicitly rounded intrinsics (no FMA contraction) so keep-sets are
//                bit-exact against the CPU oracle on identical inputs.
#include "common.cuh"

namespace ly {

namespace {

constexpr int TOPK_MAX = 1024;   // max_det supported by the per-CTA sort buffers
constexpr int NMS_CH = 512;      // candidates per NMS chunk
constexpr int NMS_PRESEL = 2048; // candidates selected and sorted up front when more pass the threshold
constexpr int NT = 1024;
constexpr int NT_TOPK = 512;     // top-k: 3 CTAs per SM so that a batch of 256 images is one wave

struct Levels {
  const float* p[4];
  int H[4], W[4], stride[4], off[4];  // off = first global anchor index of the level
  int n, B, nc, reg_max, A, direct, clamp_h, clamp_w;
};

__device__ __forceinline__ float sigmoid_precise(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ int level_of(const Levels& lv, int g) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (i < lv.n && g >= lv.off[i]) l = i;
  return l;
}

// ------------------------------------------------------------------------------------ DFL
// LABEL = false (top-k path): only the best class SCORE is needed, and sigmoid is monotonic,
// so it is sigmoid(max logit): 1 sigmoid per anchor instead of nc.  The NMS path needs the
// reference's argmax over the sigmoid VALUES (first maximum wins, ties included): LABEL = true.
// RM = 16: the standard reg_max with the 16 bin loads of a side issued together (fully unrolled);
// RM = 0: any reg_max, rolled loops.
// xyxy box of anchor (level l, cell a) of image b: DFL softmax-expectation + anchor decode (postprocess.py:217-235,
// utils/tal.py:10-46), or the legacy direct-offset layout (postprocess.py:70-92).  p = &pred[b][0][a].
template <int RM>
__device__ __forceinline__ float4 decode_box(const Levels& lv, int l, int a, const float* p) {
  const int HW = lv.H[l] * lv.W[l];
  const int y = a / lv.W[l], x = a - y * lv.W[l];
  const float s = (float)lv.stride[l];
  float x1, y1, x2, y2;
  if (lv.direct) {
    // postprocess.py:70-92: sigmoid centre offsets on the integer grid, exp sizes
    const float cx = (sigmoid_precise(p[0]) + (float)x) * s;
    const float cy = (sigmoid_precise(p[(long long)HW]) + (float)y) * s;
    const float bw = expf(p[2LL * HW]) * s, bh = expf(p[3LL * HW]) * s;
    x1 = cx - bw / 2; y1 = cy - bh / 2; x2 = cx + bw / 2; y2 = cy + bh / 2;
    if (lv.clamp_w > 0) {
      x1 = fminf(fmaxf(x1, 0.f), (float)lv.clamp_w); x2 = fminf(fmaxf(x2, 0.f), (float)lv.clamp_w);
      y1 = fminf(fmaxf(y1, 0.f), (float)lv.clamp_h); y2 = fminf(fmaxf(y2, 0.f), (float)lv.clamp_h);
    }
  } else {
    float d[4];
#pragma unroll
    for (int side = 0; side < 4; ++side) {
      const float* q = p + (long long)side * lv.reg_max * HW;
      float mx = -INFINITY, den = 0.f, num = 0.f;
      if (RM > 0) {
        float t[RM > 0 ? RM : 1];
#pragma unroll
        for (int i = 0; i < RM; ++i) t[i] = __ldg(q + (long long)i * HW);
#pragma unroll
        for (int i = 0; i < RM; ++i) mx = fmaxf(mx, t[i]);
#pragma unroll
        for (int i = 0; i < RM; ++i) {
          const float e = expf(t[i] - mx);
          den += e;
          num += e * (float)i;
        }
      } else {
        for (int i = 0; i < lv.reg_max; ++i) mx = fmaxf(mx, q[(long long)i * HW]);
        for (int i = 0; i < lv.reg_max; ++i) {
          const float e = expf(q[(long long)i * HW] - mx);
          den += e;
          num += e * (float)i;
        }
      }
      d[side] = num / den;
    }
    const float ax = (float)x + 0.5f, ay = (float)y + 0.5f;
    x1 = (ax - d[0]) * s; y1 = (ay - d[1]) * s; x2 = (ax + d[2]) * s; y2 = (ay + d[3]) * s;
  }
  return make_float4(x1, y1, x2, y2);
}

template <bool LABEL, int RM>
__global__ void __launch_bounds__(256)
dfl_kernel(Levels lv, float* __restrict__ boxes, float* __restrict__ best, int* __restrict__ label) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (g >= lv.A) return;
  const int l = level_of(lv, g);
  const int a = g - lv.off[l];
  const int HW = lv.H[l] * lv.W[l];
  const int creg = lv.direct ? 4 : 4 * lv.reg_max;
  const float* p = lv.p[l] + (long long)b * (creg + lv.nc) * HW + a;
  const float4 bx = decode_box<RM>(lv, l, a, p);
  const float x1 = bx.x, y1 = bx.y, x2 = bx.z, y2 = bx.w;
  const float* c = p + (long long)creg * HW;
  float bs = -1.f;
  int bl = 0;
  if (LABEL) {
    for (int i = 0; i < lv.nc; ++i) {
      const float v = sigmoid_precise(c[(long long)i * HW]);
      if (v > bs) { bs = v; bl = i; }   // first maximum wins, like torch.max
    }
  } else {
    float mx = -INFINITY;
#pragma unroll 8
    for (int i = 0; i < lv.nc; ++i) mx = fmaxf(mx, c[(long long)i * HW]);
    bs = sigmoid_precise(mx);
  }
  const long long o = (long long)b * lv.A + g;
  reinterpret_cast<float4*>(boxes)[o] = make_float4(x1, y1, x2, y2);
  best[o] = bs;
  if (LABEL) label[o] = bl;
}

// Top-k path: only the best class score of every anchor is needed up front (sigmoid is monotonic: sigmoid(max
// logit)); the box of an anchor is decoded later, for the k selected anchors only (topk_kernel).  Reads the class
// logits once (nc * 4 B per anchor) instead of the whole head tensor.
__global__ void __launch_bounds__(256) best_kernel(Levels lv, float* __restrict__ best) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (g >= lv.A) return;
  const int l = level_of(lv, g);
  const int a = g - lv.off[l];
  const int HW = lv.H[l] * lv.W[l];
  const int creg = 4 * lv.reg_max;
  const float* c = lv.p[l] + ((long long)b * (creg + lv.nc) + creg) * HW + a;
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  int i = 0;
  for (; i + 4 <= lv.nc; i += 4) {     // four independent loads in flight per step
    m0 = fmaxf(m0, __ldg(c + (long long)i * HW));
    m1 = fmaxf(m1, __ldg(c + (long long)(i + 1) * HW));
    m2 = fmaxf(m2, __ldg(c + (long long)(i + 2) * HW));
    m3 = fmaxf(m3, __ldg(c + (long long)(i + 3) * HW));
  }
  for (; i < lv.nc; ++i) m0 = fmaxf(m0, __ldg(c + (long long)i * HW));
  best[(long long)b * lv.A + g] = sigmoid_precise(fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)));
}

// ------------------------------------------------------------------- block-wide primitives
__device__ __forceinline__ unsigned orderable(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// k-th largest of n DISTINCT 64-bit keys (k in [1,n]); key(i) may be called many times.
// Keys are (orderable score << 32 | ~index).  Radix select, 8 bits per pass from the top; four keys per thread and
// step are fetched before any of them is binned (the loads are independent: the passes were bound by the latency of one
// dependent global load per step, ncu: topk_kernel 320 us at 0.4 IPC).  After the four SCORE passes the k-th key's
// score is known; unless equal scores straddle the cut (the bucket holds more keys than are still wanted) every key
// of that score qualifies and the four index passes are skipped: the returned threshold is (score << 32).
template <typename KeyFn>
__device__ unsigned long long block_kth_largest(KeyFn key, int n, int k, unsigned* hist, unsigned* bcast) {
  unsigned long long prefix = 0, mask = 0;
  for (int pass = 7; pass >= 0; --pass) {
    const int shift = pass * 8;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 4 * blockDim.x) {
      unsigned long long kk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * blockDim.x + threadIdx.x;
        kk[j] = i < n ? key(i) : 0ull;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + j * blockDim.x + threadIdx.x;
        unsigned digit = 0xFFFFFFFFu;  // sentinel: not a candidate
        if (i < n && (kk[j] & mask) == prefix) digit = (unsigned)(kk[j] >> shift) & 255u;
        // warp-aggregated histogram: one shared atomic per distinct digit per warp
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, digit);
        if (digit != 0xFFFFFFFFu && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[digit], __popc(peers));
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int cum = 0, d = 255;
      for (; d > 0; --d) {
        if (cum + (int)hist[d] >= k) break;
        cum += hist[d];
      }
      bcast[0] = d;
      bcast[1] = k - cum;
      bcast[2] = hist[d];
    }
    __syncthreads();
    prefix |= (unsigned long long)bcast[0] << shift;
    mask |= 0xFFull << shift;
    k = bcast[1];
    const bool whole_bucket = (int)bcast[2] == k;
    __syncthreads();
    if (pass == 4 && whole_bucket) return prefix;      // no tie across the cut: every key of this score is wanted
  }
  return prefix;
}

// in-place bitonic sort, descending, n a power of two, buffer in shared or global memory
__device__ void block_bitonic_desc(unsigned long long* a, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long x = a[lo], y = a[hi];
        if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ----------------------------------------------------------------------------------- top-k
template <int RM>
__global__ void __launch_bounds__(NT_TOPK, 3)
topk_kernel(Levels lv, int k, const float* __restrict__ best, float* s2,
            float* __restrict__ out, int* __restrict__ out_anchor, int* __restrict__ out_cls, const float* __restrict__ lb_meta) {
  __shared__ unsigned long long sortbuf[TOPK_MAX];
  __shared__ int anchors[TOPK_MAX];
  __shared__ unsigned hist[256];
  __shared__ unsigned bcast[3];
  __shared__ int cnt;
  const int b = blockIdx.x;
  const int A = lv.A, nc = lv.nc;
  const int P = next_pow2(k);
  const float* bestb = best + (long long)b * A;

  // ---- stage 1: top-k anchors by best class score
  auto key1 = [&](int i) -> unsigned long long {
    return ((unsigned long long)orderable(bestb[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
  };
  unsigned long long thr = block_kth_largest(key1, A, k, hist, bcast);
  if (threadIdx.x == 0) cnt = 0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) sortbuf[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < A; i += blockDim.x) {
    const unsigned long long kk = key1(i);
    if (kk >= thr) sortbuf[atomicAdd(&cnt, 1)] = kk;
  }
  __syncthreads();
  block_bitonic_desc(sortbuf, P);
  for (int r = threadIdx.x; r < k; r += blockDim.x) anchors[r] = (int)(0xFFFFFFFFu - (unsigned)(sortbuf[r] & 0xFFFFFFFFull));
  __syncthreads();

  // ---- stage 2: scores of the k selected anchors x nc classes
  float* s2b = s2 + (long long)b * k * nc;
  const int creg = 4 * lv.reg_max;
  for (int i = threadIdx.x; i < k * nc; i += blockDim.x) {
    const int r = i / nc, c = i - r * nc;
    const int g = anchors[r];
    const int l = level_of(lv, g);
    const int HW = lv.H[l] * lv.W[l];
    s2b[i] = sigmoid_precise(lv.p[l][((long long)b * (creg + nc) + creg + c) * HW + (g - lv.off[l])]);
  }
  __syncthreads();
  auto key2 = [&](int i) -> unsigned long long {
    return ((unsigned long long)orderable(s2b[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
  };
  thr = block_kth_largest(key2, k * nc, k, hist, bcast);
  if (threadIdx.x == 0) cnt = 0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) sortbuf[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < k * nc; i += blockDim.x) {
    const unsigned long long kk = key2(i);
    if (kk >= thr) sortbuf[atomicAdd(&cnt, 1)] = kk;
  }
  __syncthreads();
  block_bitonic_desc(sortbuf, P);

  for (int r = threadIdx.x; r < k; r += blockDim.x) {
    const int flat = (int)(0xFFFFFFFFu - (unsigned)(sortbuf[r] & 0xFFFFFFFFull));
    const int rel = flat / nc, c = flat - rel * nc;
    const int g = anchors[rel];
    const int l = level_of(lv, g);
    const float4 bx = decode_box<RM>(lv, l, g - lv.off[l], lv.p[l] + (long long)b * (creg + nc) * (lv.H[l] * lv.W[l]) + (g - lv.off[l]));
    float* o = out + ((long long)b * k + r) * 6;
    if (lb_meta) {
      // unletterbox_coords (utils/box_ops.py:96-124) fused into the decode epilogue: back to the source image's own
      // pixel coordinates, fp32 op for op like the reference ((x - pad) / gain, clamp to the image)
      const float gw = lb_meta[6 * b + 0], gh = lb_meta[6 * b + 1], px = lb_meta[6 * b + 2], py = lb_meta[6 * b + 3];
      const float H = lb_meta[6 * b + 4], W = lb_meta[6 * b + 5];
      o[0] = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.x, px), gw), 0.f), W);
      o[1] = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.y, py), gh), 0.f), H);
      o[2] = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.z, px), gw), 0.f), W);
      o[3] = fminf(fmaxf(__fdiv_rn(__fsub_rn(bx.w, py), gh), 0.f), H);
    } else {
      o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
    }
    o[4] = s2b[flat];
    o[5] = (float)c;
    if (out_anchor) out_anchor[(long long)b * k + r] = g;
    if (out_cls) out_cls[(long long)b * k + r] = c;
  }
}

// ------------------------------------------------------------------------------------- NMS
// IoU > thr, evaluated exactly like box_ops.py:31-46 in fp32 (each op individually rounded)
__device__ __forceinline__ float box_area_rn(const float4& a) {
  return __fmul_rn(fmaxf(__fsub_rn(a.z, a.x), 0.f), fmaxf(__fsub_rn(a.w, a.y), 0.f));
}
__device__ __forceinline__ bool iou_gt(const float4& a, float area_a, const float4& b, float area_b, float thr) {
  const float w = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.f);
  const float h = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.f);
  const float inter = __fmul_rn(w, h);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  const float iou = __fdiv_rn(inter, __fadd_rn(uni, 1e-9f));
  return iou > thr;
}

// torchvision.ops.nms (CPU kernel nms_kernel_impl, what export.py:182 calls): areas are NOT clamped, no epsilon
__device__ __forceinline__ float box_area_tv(const float4& a) { return __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y)); }
__device__ __forceinline__ bool iou_gt_tv(const float4& a, float area_a, const float4& b, float area_b, float thr) {
  const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
  const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return ovr > thr;     // thr = the largest float <= the double threshold (the reference compares in double)
}
template <bool TV>
__device__ __forceinline__ bool iou_gt_m(const float4& a, float area_a, const float4& b, float area_b, float thr) {
  return TV ? iou_gt_tv(a, area_a, b, area_b, thr) : iou_gt(a, area_a, b, area_b, thr);
}

struct NmsSmem {
  float4 kbox[TOPK_MAX];
  float karea[TOPK_MAX];
  int klabel[TOPK_MAX];
  float4 cbox[NMS_CH];
  float carea[NMS_CH];
  int clabel[NMS_CH];
  int cidx[NMS_CH];
  unsigned alive[NMS_CH / 32];
  unsigned mask[NMS_CH][NMS_CH / 32];
  int n_cand, n_kept, done;
};

// boxes [B,N,4], scores [B,N], labels [B,N] or null.  A candidate is valid when
// (n_valid ? i < n_valid[b] : true) && (use_conf ? score > conf : true).
// sort_g: per-image global scratch of npad u64 (used when the keys do not fit in smem).
template <bool TV>
__global__ void __launch_bounds__(NT)
nms_kernel(const float* __restrict__ boxes, const float* __restrict__ scores, const int* __restrict__ labels,
           const int* __restrict__ n_valid, int N, int npad, int use_conf, float conf, float thr, int max_keep,
           int classwise, int sort_in_smem, unsigned long long* sort_g, int* __restrict__ keep,
           int* __restrict__ keep_count, float* __restrict__ out_rows) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NmsSmem& S = *reinterpret_cast<NmsSmem*>(smem_raw);
  unsigned long long* sort_s = reinterpret_cast<unsigned long long*>(smem_raw + ((sizeof(NmsSmem) + 15) / 16) * 16);
  const int b = blockIdx.x;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (long long)b * N;
  const float* sc = scores + (long long)b * N;
  const int* lb = labels ? labels + (long long)b * N : nullptr;
  const int nv = n_valid ? min(n_valid[b], N) : N;
  unsigned long long* sb = sort_in_smem ? sort_s : sort_g + (long long)b * npad;

  // Candidate keys, score-descending.  The greedy scan stops at max_keep survivors, so it almost never looks past the
  // first few thousand candidates: when more than NMS_PRESEL pass the threshold (config 3: all 8400 anchors) only the
  // NMS_PRESEL best are selected (radix select on the keys computed on the fly) and sorted; if they run out before
  // max_keep boxes survive, the kernel falls back to the full sort and starts over (identical result either way).
  // ncu on config 3 before this: 1.10 ms per launch, DRAM 0.4 %, L1 66 %: the 16384-key shared-memory bitonic sort.
  auto key_of = [&](int i) -> unsigned long long {
    const float s = sc[i];
    if (use_conf && !(s > conf)) return 0ull;      // below any real key (a real key has a non-zero index part or score part)
    return ((unsigned long long)orderable(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
  };
  __shared__ unsigned hist_s[256];
  __shared__ unsigned bcast_s[3];
  int* keepb = keep + (long long)b * max_keep;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int WORDS = NMS_CH / 32;
  for (int attempt = 0; attempt < 2; ++attempt) {
  if (threadIdx.x == 0) { S.n_cand = 0; S.n_kept = 0; S.done = 0; }
  __syncthreads();
  int total_valid = 0;
  {
    int local = 0;
    for (int i = threadIdx.x; i < nv; i += blockDim.x) local += key_of(i) != 0ull;
    local = __reduce_add_sync(0xFFFFFFFFu, local);
    if (lane == 0 && local) atomicAdd(&S.n_cand, local);
    __syncthreads();
    total_valid = S.n_cand;
    __syncthreads();
    if (threadIdx.x == 0) S.n_cand = 0;
    __syncthreads();
  }
  const bool presel = attempt == 0 && total_valid > NMS_PRESEL;
  unsigned long long thr_key = 1ull;               // keep every valid key
  if (presel) thr_key = block_kth_largest(key_of, nv, NMS_PRESEL, hist_s, bcast_s);
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    const unsigned long long kk = key_of(i);
    if (kk >= thr_key && kk != 0ull) sb[atomicAdd(&S.n_cand, 1)] = kk;
  }
  __syncthreads();
  const int n = S.n_cand;
  const int P = next_pow2(n > 1 ? n : 1);
  for (int i = n + threadIdx.x; i < P; i += blockDim.x) sb[i] = 0;
  __syncthreads();
  block_bitonic_desc(sb, P);


  for (int c0 = 0; c0 < n; c0 += NMS_CH) {
    const int m = min(NMS_CH, n - c0);
    for (int i = threadIdx.x; i < NMS_CH; i += blockDim.x) {
      if (i < m) {
        const int idx = (int)(0xFFFFFFFFu - (unsigned)(sb[c0 + i] & 0xFFFFFFFFull));
        const float4 v = bx[idx];
        S.cidx[i] = idx;
        S.cbox[i] = v;
        S.carea[i] = TV ? box_area_tv(v) : box_area_rn(v);
        S.clabel[i] = lb ? lb[idx] : 0;
      }
    }
    for (int i = threadIdx.x; i < WORDS; i += blockDim.x) {
      const int lo = i * 32;
      S.alive[i] = m >= lo + 32 ? 0xFFFFFFFFu : (m > lo ? ((1u << (m - lo)) - 1u) : 0u);
    }
    __syncthreads();
    // ---- (A) suppression by the kept set: one warp per candidate, lanes stride the kept list
    const int nk = S.n_kept;
    if (nk > 0) {
      for (int i = warp; i < m; i += NT / 32) {
        const float4 v = S.cbox[i];
        const float av = S.carea[i];
        const int lv = S.clabel[i];
        bool hit = false;
        for (int j = lane; j < nk && !hit; j += 32)
          hit = (!classwise || S.klabel[j] == lv) && iou_gt_m<TV>(S.kbox[j], S.karea[j], v, av, thr);
        if (__any_sync(0xFFFFFFFFu, hit) && lane == 0) atomicAnd(&S.alive[i >> 5], ~(1u << (i & 31)));
      }
    }
    // ---- (B) intra-chunk bitmask: bit j of mask[i] set iff j > i and i suppresses j
    for (int t = threadIdx.x; t < m * WORDS; t += blockDim.x) {
      const int i = t / WORDS, wj = t - i * WORDS;
      unsigned bits = 0;
      if (wj * 32 + 31 > i) {
        const float4 v = S.cbox[i];
        const float av = S.carea[i];
        const int lv = S.clabel[i];
        const int jend = min(32, m - wj * 32);
        for (int jj = 0; jj < jend; ++jj) {
          const int j = wj * 32 + jj;
          if (j > i && (!classwise || S.clabel[j] == lv) && iou_gt_m<TV>(v, av, S.cbox[j], S.carea[j], thr)) bits |= 1u << jj;
        }
      }
      S.mask[i][wj] = bits;
    }
    __syncthreads();
    // ---- (C) sequential scan by warp 0: lane w owns word w of the "removed" set
    if (warp == 0) {
      unsigned removed = lane < WORDS ? ~S.alive[lane] : 0xFFFFFFFFu;
      int kept = S.n_kept;
      for (int wi = 0; wi < (m + 31) / 32 && kept < max_keep; ++wi) {
        // candidates of word wi that are still alive form the work list; resolve in order
        unsigned cur = __shfl_sync(0xFFFFFFFFu, removed, wi);
        for (int bit = 0; bit < 32 && kept < max_keep; ++bit) {
          const int i = wi * 32 + bit;
          if (i >= m) break;
          if (cur & (1u << bit)) continue;
          // keep i
          if (lane == 0) {
            S.kbox[kept] = S.cbox[i];
            S.karea[kept] = S.carea[i];
            S.klabel[kept] = S.clabel[i];
            keepb[kept] = S.cidx[i];
          }
          ++kept;
          if (lane < WORDS) removed |= S.mask[i][lane];
          cur = __shfl_sync(0xFFFFFFFFu, removed, wi);
        }
      }
      if (lane == 0) {
        S.n_kept = kept;
        if (kept >= max_keep) S.done = 1;
      }
    }
    __syncthreads();
    if (S.done) break;
  }
  __syncthreads();
  // the preselected candidates ran out before max_keep boxes survived and there are more: full sort, start over
  if (!(presel && S.n_kept < max_keep)) break;
  __syncthreads();
  }
  __syncthreads();
  const int kept = S.n_kept;
  if (threadIdx.x == 0) keep_count[b] = kept;
  for (int r = kept + threadIdx.x; r < max_keep; r += blockDim.x) keepb[r] = -1;
  if (out_rows) {
    for (int r = threadIdx.x; r < max_keep; r += blockDim.x) {
      float* o = out_rows + ((long long)b * max_keep + r) * 6;
      if (r < kept) {
        const int idx = keepb[r];
        const float4 v = S.kbox[r];
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
        o[4] = sc[idx];
        o[5] = (float)S.klabel[r];
      } else {
        o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.f;
      }
    }
  }
}

// ------------------------------------------------------------------------ export-style outputs
// export.py:126-144 (nms = False): top-k anchors by best class score with everything below `conf` masked to -1,
// argmax class, boxes clamped to the image, num_dets = #(score >= conf).  One CTA per image.
__global__ void __launch_bounds__(NT_TOPK, 3)
export_topk_kernel(int A, int k, float conf, float img_w, float img_h, const float* __restrict__ boxes, const float* __restrict__ best,
                   const int* __restrict__ label, float* __restrict__ out, int* __restrict__ num) {
  __shared__ unsigned long long sortbuf[TOPK_MAX];
  __shared__ unsigned hist[256];
  __shared__ unsigned bcast[3];
  __shared__ int cnt, nvalid;
  const int b = blockIdx.x;
  const int P = next_pow2(k);
  const float* bestb = best + (long long)b * A;
  auto key = [&](int i) -> unsigned long long {
    const float v = bestb[i];
    return ((unsigned long long)orderable(v >= conf ? v : -1.0f) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
  };
  const unsigned long long thr = block_kth_largest(key, A, k, hist, bcast);
  if (threadIdx.x == 0) { cnt = 0; nvalid = 0; }
  for (int i = threadIdx.x; i < P; i += blockDim.x) sortbuf[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < A; i += blockDim.x) {
    const unsigned long long kk = key(i);
    if (kk >= thr) sortbuf[atomicAdd(&cnt, 1)] = kk;
  }
  __syncthreads();
  block_bitonic_desc(sortbuf, P);
  for (int r = threadIdx.x; r < k; r += blockDim.x) {
    const int g = (int)(0xFFFFFFFFu - (unsigned)(sortbuf[r] & 0xFFFFFFFFull));
    const float4 bx = reinterpret_cast<const float4*>(boxes)[(long long)b * A + g];
    const float sc = fmaxf(bestb[g], 0.f);
    float* o = out + ((long long)b * k + r) * 6;
    o[0] = fminf(fmaxf(bx.x, 0.f), img_w); o[1] = fminf(fmaxf(bx.y, 0.f), img_h);
    o[2] = fminf(fmaxf(bx.z, 0.f), img_w); o[3] = fminf(fmaxf(bx.w, 0.f), img_h);
    o[4] = sc;
    o[5] = (float)label[(long long)b * A + g];
    if (sc >= conf) atomicAdd(&nvalid, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) num[b] = nvalid;
}

// export.py:148-176: the k2 best (anchor, class) pairs of an image, score-descending with the canonical tie rule
// (score desc, flat index anchor*nc + class asc), and their boxes offset per (image, class) group in fp32.
// The pairs of the flat top-k all belong to the k1 = min(k2, A) best anchors by best class score (each of those
// anchors contributes a pair that precedes any pair of an anchor outside the set), so the selection is two-stage
// like topk_kernel: anchors first, then k1*nc pair scores.
__global__ void __launch_bounds__(NT_TOPK, 3)
export_pairs_kernel(Levels lv, int k1, int k2, int img0, float group_off, const float* __restrict__ boxes,
                    const float* __restrict__ best, float* s2, float* __restrict__ cand_box, float* __restrict__ cand_off,
                    float* __restrict__ cand_score, int* __restrict__ cand_cls) {
  __shared__ unsigned long long sortbuf[TOPK_MAX];
  __shared__ int anchors[TOPK_MAX];
  __shared__ unsigned hist[256];
  __shared__ unsigned bcast[3];
  __shared__ int cnt;
  const int b = blockIdx.x;
  const int A = lv.A, nc = lv.nc;
  const float* bestb = best + (long long)b * A;
  auto key1 = [&](int i) -> unsigned long long {
    return ((unsigned long long)orderable(bestb[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
  };
  unsigned long long thr = block_kth_largest(key1, A, k1, hist, bcast);
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < A; i += blockDim.x)
    if (key1(i) >= thr) anchors[atomicAdd(&cnt, 1)] = i;     // any order: stage 2 keys carry the flat index
  __syncthreads();
  float* s2b = s2 + (long long)b * k1 * nc;
  const int creg = 4 * lv.reg_max;
  for (int i = threadIdx.x; i < k1 * nc; i += blockDim.x) {
    const int r = i / nc, c = i - r * nc;
    const int g = anchors[r];
    const int l = level_of(lv, g);
    const int HW = lv.H[l] * lv.W[l];
    s2b[i] = sigmoid_precise(lv.p[l][((long long)b * (creg + nc) + creg + c) * HW + (g - lv.off[l])]);
  }
  __syncthreads();
  auto key2 = [&](int i) -> unsigned long long {
    const int r = i / nc, c = i - r * nc;
    return ((unsigned long long)orderable(s2b[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)(anchors[r] * nc + c));
  };
  thr = block_kth_largest(key2, k1 * nc, k2, hist, bcast);
  const int P = next_pow2(k2);
  if (threadIdx.x == 0) cnt = 0;
  for (int i = threadIdx.x; i < P; i += blockDim.x) sortbuf[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < k1 * nc; i += blockDim.x) {
    const unsigned long long kk = key2(i);
    if (kk >= thr) sortbuf[atomicAdd(&cnt, 1)] = kk;
  }
  __syncthreads();
  block_bitonic_desc(sortbuf, P);
  for (int r = threadIdx.x; r < k2; r += blockDim.x) {
    const unsigned flat = 0xFFFFFFFFu - (unsigned)(sortbuf[r] & 0xFFFFFFFFull);
    const int g = (int)(flat / (unsigned)nc), c = (int)(flat - (unsigned)g * (unsigned)nc);
    const float4 bx = reinterpret_cast<const float4*>(boxes)[(long long)b * A + g];
    const long long o = (long long)b * k2 + r;
    // export.py:165-170: group_id = img * C + cls; off = group_id * (10 * max(H, W)); every step rounded to fp32
    const float gid = __fadd_rn(__fmul_rn((float)(img0 + b), (float)nc), (float)c);
    const float off = __fmul_rn(gid, group_off);
    reinterpret_cast<float4*>(cand_box)[o] = bx;
    reinterpret_cast<float4*>(cand_off)[o] = make_float4(__fadd_rn(bx.x, off), __fadd_rn(bx.y, off), __fadd_rn(bx.z, off), __fadd_rn(bx.w, off));
    // the score is recomputed from the key's bits (exactly the value that was sorted)
    const unsigned u = (unsigned)(sortbuf[r] >> 32);
    cand_score[o] = __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
    cand_cls[o] = c;
  }
}

// export.py:184-198: the first max_det survivors (already score-descending), zeroed below `conf`
__global__ void export_finish_kernel(int k2, int kmax, float conf, float img_w, float img_h, const float* __restrict__ cand_box,
                                     const float* __restrict__ cand_score, const int* __restrict__ cand_cls,
                                     const int* __restrict__ keep, const int* __restrict__ keep_count, float* __restrict__ out,
                                     int* __restrict__ num) {
  const int b = blockIdx.x;
  __shared__ int nvalid;
  if (threadIdx.x == 0) nvalid = 0;
  __syncthreads();
  const int kept = keep_count[b];
  for (int r = threadIdx.x; r < kmax; r += blockDim.x) {
    float* o = out + ((long long)b * kmax + r) * 6;
    bool ok = false;
    if (r < kept) {
      const int idx = keep[(long long)b * kmax + r];
      const float sc = cand_score[(long long)b * k2 + idx];
      if (sc >= conf) {
        const float4 bx = reinterpret_cast<const float4*>(cand_box)[(long long)b * k2 + idx];
        o[0] = fminf(fmaxf(bx.x, 0.f), img_w); o[1] = fminf(fmaxf(bx.y, 0.f), img_h);
        o[2] = fminf(fmaxf(bx.z, 0.f), img_w); o[3] = fminf(fmaxf(bx.w, 0.f), img_h);
        o[4] = sc;
        o[5] = (float)cand_cls[(long long)b * k2 + idx];
        ok = true;
        atomicAdd(&nvalid, 1);
      }
    }
    if (!ok) o[0] = o[1] = o[2] = o[3] = o[4] = o[5] = 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) num[b] = nvalid;
}

int32_t make_levels(const ly_levels* in, Levels& lv) {
  LY_CHECK_ARG(in && in->n_levels >= 1 && in->n_levels <= 4, "decode: n_levels must be 1..4");
  LY_CHECK_ARG(in->B >= 1 && in->nc >= 1 && in->reg_max >= 1, "decode: bad B/nc/reg_max");
  lv.n = in->n_levels; lv.B = in->B; lv.nc = in->nc; lv.reg_max = in->reg_max;
  lv.direct = in->direct; lv.clamp_h = in->clamp_h; lv.clamp_w = in->clamp_w;
  int off = 0;
  for (int i = 0; i < 4; ++i) {
    if (i < lv.n) {
      LY_CHECK_ARG(in->preds[i] && in->H[i] > 0 && in->W[i] > 0, "decode: bad level %d", i);
      lv.p[i] = in->preds[i]; lv.H[i] = in->H[i]; lv.W[i] = in->W[i]; lv.stride[i] = in->stride[i];
      lv.off[i] = off;
      off += in->H[i] * in->W[i];
    } else {
      lv.p[i] = nullptr; lv.H[i] = lv.W[i] = lv.stride[i] = 0; lv.off[i] = 0x7FFFFFFF;
    }
  }
  lv.A = off;
  return LY_OK;
}

inline long long rup256(long long v) { return (v + 255) / 256 * 256; }
inline int pow2_ge(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct Scratch {
  float* boxes; float* best; int* label; float* s2; unsigned long long* sortbuf; int* keep; int* count;
  long long total;
};

Scratch carve(void* base, int B, int A, int nc, int max_det) {
  Scratch s;
  long long o = 0;
  auto take = [&](long long bytes) { long long r = o; o += rup256(bytes); return (char*)base + r; };
  s.boxes = (float*)take((long long)B * A * 16);
  s.best = (float*)take((long long)B * A * 4);
  s.label = (int*)take((long long)B * A * 4);
  s.s2 = (float*)take((long long)B * (max_det < A ? max_det : A) * nc * 4);
  s.sortbuf = (unsigned long long*)take((long long)B * pow2_ge(A) * 8);
  s.keep = (int*)take((long long)B * max_det * 4);
  s.count = (int*)take((long long)B * 4);
  s.total = o;
  return s;
}

constexpr int kSmemSortMaxKeys = 16384;  // 128 KB of keys next to ~65 KB of NMS state

int32_t run_nms(const float* boxes, const float* scores, const int* labels, const int* n_valid, int B, int N,
                int use_conf, float conf, float thr, int max_keep, int classwise, unsigned long long* sort_g,
                int* keep, int* keep_count, float* out_rows, cudaStream_t st, bool tv = false) {
  LY_CHECK_ARG(max_keep >= 1 && max_keep <= TOPK_MAX, "nms: max_keep must be in 1..%d", TOPK_MAX);
  const int npad = pow2_ge(N);
  const int in_smem = npad <= kSmemSortMaxKeys;
  const size_t smem = ((sizeof(NmsSmem) + 15) / 16) * 16 + (in_smem ? (size_t)npad * 8 : 0);
  // (the attribute is per device: set it on every call, it is cheap)
  LY_CUDA(cudaFuncSetAttribute(nms_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  LY_CUDA(cudaFuncSetAttribute(nms_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  if (tv)
    nms_kernel<true><<<B, NT, smem, st>>>(boxes, scores, labels, n_valid, N, npad, use_conf, conf, thr, max_keep, classwise,
                                          in_smem, sort_g, keep, keep_count, out_rows);
  else
    nms_kernel<false><<<B, NT, smem, st>>>(boxes, scores, labels, n_valid, N, npad, use_conf, conf, thr, max_keep, classwise,
                                           in_smem, sort_g, keep, keep_count, out_rows);
  return post_launch("nms");
}

}  // namespace

}  // namespace ly

using namespace ly;

extern "C" int64_t ly_decode_scratch_bytes(const ly_levels* in, int32_t max_det) {
  Levels lv;
  if (make_levels(in, lv) != LY_OK || max_det < 1) return -1;
  return carve(nullptr, lv.B, lv.A, lv.nc, max_det).total;
}

static int32_t decode_topk_impl(const ly_levels* in, int32_t max_det, const float* lb_meta, float* out, int32_t* out_anchor,
                                int32_t* out_cls, void* scratch, int64_t scratch_bytes, void* stream);

extern "C" int32_t ly_decode_topk(const ly_levels* in, int32_t max_det, float* out, int32_t* out_anchor, int32_t* out_cls,
                                  void* scratch, int64_t scratch_bytes, void* stream) {
  return decode_topk_impl(in, max_det, nullptr, out, out_anchor, out_cls, scratch, scratch_bytes, stream);
}

extern "C" int32_t ly_decode_topk_lb(const ly_levels* in, int32_t max_det, const float* lb_meta, float* out, int32_t* out_anchor,
                                     int32_t* out_cls, void* scratch, int64_t scratch_bytes, void* stream) {
  LY_CHECK_ARG(lb_meta != nullptr, "decode_topk_lb: null letterbox meta");
  return decode_topk_impl(in, max_det, lb_meta, out, out_anchor, out_cls, scratch, scratch_bytes, stream);
}

static int32_t decode_topk_impl(const ly_levels* in, int32_t max_det, const float* lb_meta, float* out, int32_t* out_anchor,
                                int32_t* out_cls, void* scratch, int64_t scratch_bytes, void* stream) {
  Levels lv;
  int32_t rc = make_levels(in, lv);
  if (rc != LY_OK) return rc;
  LY_CHECK_ARG(!lv.direct, "decode_topk: DFL layout only");
  LY_CHECK_ARG(max_det >= 1 && max_det <= TOPK_MAX, "decode_topk: max_det must be in 1..%d", TOPK_MAX);
  LY_CHECK_ARG(out && scratch, "decode_topk: null pointer");
  Scratch s = carve(scratch, lv.B, lv.A, lv.nc, max_det);
  LY_CHECK_ARG(scratch_bytes >= s.total, "decode_topk: scratch too small (%lld < %lld)", (long long)scratch_bytes, s.total);
  cudaStream_t st = (cudaStream_t)stream;
  const int k = max_det < lv.A ? max_det : lv.A;
  dim3 g1((lv.A + 255) / 256, lv.B);
  best_kernel<<<g1, 256, 0, st>>>(lv, s.best);
  rc = post_launch("best_score");
  if (rc != LY_OK) return rc;
  if (lv.reg_max == 16) topk_kernel<16><<<lv.B, NT_TOPK, 0, st>>>(lv, k, s.best, s.s2, out, out_anchor, out_cls, lb_meta);
  else topk_kernel<0><<<lv.B, NT_TOPK, 0, st>>>(lv, k, s.best, s.s2, out, out_anchor, out_cls, lb_meta);
  return post_launch("topk");
}

extern "C" int32_t ly_decode_nms(const ly_levels* in, float conf_thresh, float iou_thresh, int32_t max_det, int32_t classwise,
                                 float* out, int32_t* out_count, int32_t* out_anchor, void* scratch, int64_t scratch_bytes,
                                 void* stream) {
  Levels lv;
  int32_t rc = make_levels(in, lv);
  if (rc != LY_OK) return rc;
  LY_CHECK_ARG(out && out_count && scratch, "decode_nms: null pointer");
  LY_CHECK_ARG(max_det >= 1 && max_det <= TOPK_MAX, "decode_nms: max_det must be in 1..%d", TOPK_MAX);
  Scratch s = carve(scratch, lv.B, lv.A, lv.nc, max_det);
  LY_CHECK_ARG(scratch_bytes >= s.total, "decode_nms: scratch too small (%lld < %lld)", (long long)scratch_bytes, s.total);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 g1((lv.A + 255) / 256, lv.B);
  if (lv.reg_max == 16) dfl_kernel<true, 16><<<g1, 256, 0, st>>>(lv, s.boxes, s.best, s.label);
  else dfl_kernel<true, 0><<<g1, 256, 0, st>>>(lv, s.boxes, s.best, s.label);
  rc = post_launch("dfl_decode");
  if (rc != LY_OK) return rc;
  return run_nms(s.boxes, s.best, s.label, nullptr, lv.B, lv.A, 1, conf_thresh, iou_thresh, max_det, classwise, s.sortbuf,
                 out_anchor ? out_anchor : s.keep, out_count, out, st);
}

extern "C" int64_t ly_nms_scratch_bytes(int32_t B, int32_t N) {
  if (B < 1 || N < 1) return -1;
  return rup256((long long)B * pow2_ge(N) * 8);
}

extern "C" int32_t ly_nms(const float* boxes, const float* scores, const int32_t* labels, const int32_t* n_valid, int32_t B,
                          int32_t N, float iou_thresh, int32_t max_keep, int32_t classwise, int32_t* keep,
                          int32_t* keep_count, void* scratch, int64_t scratch_bytes, void* stream) {
  LY_CHECK_ARG(boxes && scores && keep && keep_count && scratch, "nms: null pointer");
  LY_CHECK_ARG(B >= 1 && N >= 1, "nms: B and N must be >= 1");
  LY_CHECK_ARG(!classwise || labels, "nms: classwise needs labels");
  LY_CHECK_ARG(scratch_bytes >= ly_nms_scratch_bytes(B, N), "nms: scratch too small");
  return run_nms(boxes, scores, labels, n_valid, B, N, 0, 0.f, iou_thresh, max_keep, classwise,
                 (unsigned long long*)scratch, keep, keep_count, nullptr, (cudaStream_t)stream);
}

// ---- export-style fixed-shape outputs --------------------------------------
namespace ly { namespace {
struct ExportScratch {
  Scratch d; float* cand_box; float* cand_off; float* cand_score; int* cand_cls; long long total;
};
ExportScratch carve_export(void* base, int B, int A, int nc, int max_det, int pre_topk) {
  ExportScratch e;
  const int k1 = pre_topk < A ? pre_topk : A;
  e.d = carve(base, B, A, nc, k1 > max_det ? k1 : max_det);
  long long o = e.d.total;
  auto take = [&](long long bytes) { long long r = o; o += rup256(bytes); return (char*)base + r; };
  e.cand_box = (float*)take((long long)B * pre_topk * 16);
  e.cand_off = (float*)take((long long)B * pre_topk * 16);
  e.cand_score = (float*)take((long long)B * pre_topk * 4);
  e.cand_cls = (int*)take((long long)B * pre_topk * 4);
  e.total = o;
  return e;
}
} }  // namespace ly::

extern "C" int64_t ly_decode_export_scratch_bytes(const ly_levels* in, int32_t max_det, int32_t pre_topk) {
  Levels lv;
  if (make_levels(in, lv) != LY_OK || max_det < 1 || pre_topk < 1) return -1;
  return carve_export(nullptr, lv.B, lv.A, lv.nc, max_det, pre_topk).total;
}

extern "C" int32_t ly_decode_export(const ly_levels* in, float conf, int32_t max_det, int32_t nms, double iou_thresh, int32_t pre_topk,
                                    int32_t img_h, int32_t img_w, int32_t img0, float* out, int32_t* num_dets, void* scratch,
                                    int64_t scratch_bytes, void* stream) {
  Levels lv;
  int32_t rc = make_levels(in, lv);
  if (rc != LY_OK) return rc;
  LY_CHECK_ARG(!lv.direct, "decode_export: DFL layout only");
  LY_CHECK_ARG(out && num_dets && scratch, "decode_export: null pointer");
  LY_CHECK_ARG(max_det >= 1 && max_det <= TOPK_MAX && pre_topk >= 1 && pre_topk <= TOPK_MAX, "decode_export: max_det and pre_topk must be in 1..%d", TOPK_MAX);
  LY_CHECK_ARG(img_h > 0 && img_w > 0 && img0 >= 0, "decode_export: bad image size / index");
  ExportScratch e = carve_export(scratch, lv.B, lv.A, lv.nc, max_det, pre_topk);
  LY_CHECK_ARG(scratch_bytes >= e.total, "decode_export: scratch too small (%lld < %lld)", (long long)scratch_bytes, e.total);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 g1((lv.A + 255) / 256, lv.B);
  if (lv.reg_max == 16) dfl_kernel<true, 16><<<g1, 256, 0, st>>>(lv, e.d.boxes, e.d.best, e.d.label);
  else dfl_kernel<true, 0><<<g1, 256, 0, st>>>(lv, e.d.boxes, e.d.best, e.d.label);
  rc = post_launch("dfl_decode");
  if (rc != LY_OK) return rc;
  if (!nms) {
    const int k = max_det < lv.A ? max_det : lv.A;
    export_topk_kernel<<<lv.B, NT_TOPK, 0, st>>>(lv.A, k, conf, (float)img_w, (float)img_h, e.d.boxes, e.d.best, e.d.label, out, num_dets);
    return post_launch("export_topk");
  }
  const long long pairs = (long long)lv.A * lv.nc;
  const int k2 = pairs < pre_topk ? (int)pairs : pre_topk;
  const int k1 = k2 < lv.A ? k2 : lv.A;
  const int kmax = max_det < k2 ? max_det : k2;
  const float group_off = (float)((img_h > img_w ? img_h : img_w) * 10.0);
  export_pairs_kernel<<<lv.B, NT_TOPK, 0, st>>>(lv, k1, k2, img0, group_off, e.d.boxes, e.d.best, e.d.s2, e.cand_box, e.cand_off,
                                                e.cand_score, e.cand_cls);
  rc = post_launch("export_pairs");
  if (rc != LY_OK) return rc;
  // torchvision compares the fp32 overlap with the DOUBLE threshold: ovr > thr_d  <=>  ovr > (largest float <= thr_d)
  float thr_f = (float)iou_thresh;
  if ((double)thr_f > iou_thresh) thr_f = nextafterf(thr_f, -INFINITY);
  rc = run_nms(e.cand_off, e.cand_score, nullptr, nullptr, lv.B, k2, 0, 0.f, thr_f, kmax, 0, e.d.sortbuf, e.d.keep, e.d.count, nullptr, st, true);
  if (rc != LY_OK) return rc;
  export_finish_kernel<<<lv.B, 256, 0, st>>>(k2, kmax, conf, (float)img_w, (float)img_h, e.cand_box, e.cand_score, e.cand_cls, e.d.keep,
                                             e.d.count, out, num_dets);
  return post_launch("export_finish");
}
