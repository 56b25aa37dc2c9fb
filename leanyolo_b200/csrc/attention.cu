// PSA attention core (leanyolo/models/yolov10/layers.py:369-378), fused flash-style:
//   out[n, h*hd + d] = sum_m softmax_m( (q_n . k_m) * scale ) * v_m[d]
// The N x N score matrix never reaches HBM (the reference materialises 4 x 400^2 fp32 per
// image).  Small-N (400 tokens @640^2, 1600 @1280^2), so this is a latency/bandwidth
// kernel: one thread per query keeps q, the running max/sum and the output row in
// registers; K/V blocks are staged in shared memory and read as warp-wide broadcasts.
//
// Channel layout of the qkv buffer (re-ordered at weight-pack time, modules.py):
//   [ q: nh x kdp | k: nh x kdp | v: nh x hd ]   (kdp = key_dim padded to 8, pad = zeros)
#include "common.cuh"
#include "tma.cuh"

namespace ly {

namespace {

constexpr int QB = 128;  // queries (= threads) per CTA
constexpr int KB = 32;   // keys per shared-memory block

template <typename T, int KDP, int HD>
__global__ void __launch_bounds__(QB)
attn_kernel(const T* __restrict__ qkv, int N, int sCtot, int sC0, int nh, T* __restrict__ out, int dCtot, int dC0,
            float scale) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float Ks[KB][KDP];
  __shared__ __align__(16) float Vs[KB][HD];
  constexpr int V = Elem<T>::kVec;
  const int h = blockIdx.y, b = blockIdx.z;
  const int n = blockIdx.x * QB + threadIdx.x;
  const bool active = n < N;
  const T* base = qkv + (long long)b * N * sCtot + sC0;
  const int qoff = h * KDP, koff = nh * KDP + h * KDP, voff = 2 * nh * KDP + h * HD;

  float q[KDP];
  if (active) {
#pragma unroll
    for (int d = 0; d < KDP; d += V) {
      float t[V];
      load_vec<T>(base + (long long)n * sCtot + qoff + d, t);
#pragma unroll
      for (int j = 0; j < V; ++j) q[d + j] = t[j] * scale;
    }
  } else {
#pragma unroll
    for (int d = 0; d < KDP; ++d) q[d] = 0.f;
  }
  float o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) o[d] = 0.f;
  float mx = -INFINITY, l = 0.f;

  for (int m0 = 0; m0 < N; m0 += KB) {
    __syncthreads();
    // stage K/V block (zero rows past N are masked below)
    for (int i = threadIdx.x; i < KB * (KDP / V); i += QB) {
      const int r = i / (KDP / V), d = (i - r * (KDP / V)) * V;
      float t[V];
      if (m0 + r < N) load_vec<T>(base + (long long)(m0 + r) * sCtot + koff + d, t);
      else {
#pragma unroll
        for (int j = 0; j < V; ++j) t[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < V; ++j) Ks[r][d + j] = t[j];
    }
    for (int i = threadIdx.x; i < KB * (HD / V); i += QB) {
      const int r = i / (HD / V), d = (i - r * (HD / V)) * V;
      float t[V];
      if (m0 + r < N) load_vec<T>(base + (long long)(m0 + r) * sCtot + voff + d, t);
      else {
#pragma unroll
        for (int j = 0; j < V; ++j) t[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < V; ++j) Vs[r][d + j] = t[j];
    }
    __syncthreads();
    const int nk = min(KB, N - m0);
    float s[KB];
    float bm = -INFINITY;
#pragma unroll
    for (int r = 0; r < KB; ++r) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < KDP; d += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(&Ks[r][d]);
        acc = fmaf(q[d], kv.x, acc);
        acc = fmaf(q[d + 1], kv.y, acc);
        acc = fmaf(q[d + 2], kv.z, acc);
        acc = fmaf(q[d + 3], kv.w, acc);
      }
      s[r] = r < nk ? acc : -INFINITY;
      bm = fmaxf(bm, s[r]);
    }
    const float nm = fmaxf(mx, bm);
    const float corr = sizeof(T) == 4 ? expf(mx - nm) : __expf(mx - nm);  // mx=-inf first time -> 0
    l *= corr;
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] *= corr;
    mx = nm;
#pragma unroll
    for (int r = 0; r < KB; ++r) {
      const float p = sizeof(T) == 4 ? expf(s[r] - mx) : __expf(s[r] - mx);  // masked keys: exp(-inf) = 0
      l += p;
#pragma unroll
      for (int d = 0; d < HD; d += 4) {
        const float4 vv = *reinterpret_cast<const float4*>(&Vs[r][d]);
        o[d] = fmaf(p, vv.x, o[d]);
        o[d + 1] = fmaf(p, vv.y, o[d + 1]);
        o[d + 2] = fmaf(p, vv.z, o[d + 2]);
        o[d + 3] = fmaf(p, vv.w, o[d + 3]);
      }
    }
  }
  if (!active) return;
  const float inv = 1.0f / l;
  T* op = out + ((long long)b * N + n) * dCtot + dC0 + h * HD;
#pragma unroll
  for (int d = 0; d < HD; d += V) {
    float t[V];
#pragma unroll
    for (int j = 0; j < V; ++j) t[j] = o[d + j] * inv;
    store_vec<T>(op + d, t);
  }
}

// ---------------------------------------------------------------------------------------
// bf16 hot path: flash-style attention on the warp-level tensor-core path.  The op is 0.5 %
// of the model's FLOPs on a 20x20 map (N = 400 tokens, 4 heads): far too small for a
// tcgen05/TMEM pipeline to amortise, so each warp owns 16 query rows and runs
// mma.sync.m16n8k16 (bf16 x bf16 -> fp32) for both S = Q K^T and O = P V, with the online
// softmax kept in registers between them (S accumulator fragments are re-packed in place as
// the A operand of the second MMA).  K/V blocks of 80 keys are double-buffered in shared
// memory with cp.async (zero-filled past N and in the padded key columns); K fragments come
// from ldmatrix, V^T fragments from ldmatrix.trans.  5 warps per CTA = 80 queries: N = 400 tiles with no waste.
// ---------------------------------------------------------------------------------------
constexpr int AT_WARPS = 5;
constexpr int AT_KB = 80;    // keys per shared-memory block (400 = 5 x 80 and 1600 = 20 x 80 tokens: no masked tail at 640^2 / 1280^2)

__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;   // src-size 0 => 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

template <int KDP, int HD>
__global__ void __launch_bounds__(AT_WARPS * 32, 3)
attn_mma_kernel(const __nv_bfloat16* __restrict__ qkv, int N, int sCtot, int sC0, int nh, __nv_bfloat16* __restrict__ out,
                int dCtot, int dC0, float scale_log2e) {
  constexpr int KD16 = (KDP + 15) / 16 * 16;
  constexpr int KS = KD16 / 16;                    // k-steps of Q K^T
  constexpr int NT = HD / 8;                       // n-tiles of the output
  constexpr int KPITCH = KD16 * 2 + 16;            // bytes; (pitch / 4) mod 8 == 4 => conflict-free fragment loads
  constexpr int VPITCH = HD * 2 + (((HD * 2 / 16) & 1) ? 0 : 16);   // odd number of 16-byte units per row
  constexpr int KCH = KD16 * 2 / 16, VCH = HD * 2 / 16;             // 16-byte chunks per row
  constexpr int STAGE = AT_KB * (KPITCH + VPITCH);
  static_assert(HD % 8 == 0 && KDP % 8 == 0, "head dims must be multiples of 8");
  __shared__ __align__(16) uint8_t sm[2 * STAGE];
  pdl_trigger();
  pdl_wait();

  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const __nv_bfloat16* base = qkv + (long long)b * N * sCtot + sC0;
  const int qoff = h * KDP, koff = nh * KDP + h * KDP, voff = 2 * nh * KDP + h * HD;
  const uint32_t sm_u = smem_u32(sm);

  // Loader mapping, fixed for the kernel's life: thread -> (row lr0 of a pass, 16-byte chunk lc of the K | V row); one pass
  // covers RP key rows, so a block costs PASSES cp.async + pointer bumps per thread.  (ncu on the round-1 loop, which
  // re-derived (row, chunk) from a flat index: 65 instructions per chunk = 40 % of the kernel's instruction stream.)
  constexpr int TOT = KCH + VCH, RP = (AT_WARPS * 32) / TOT, PASSES = (AT_KB + RP - 1) / RP;
  const int lc = threadIdx.x % TOT, lr0 = threadIdx.x / TOT;
  const bool l_on = lr0 < RP;
  const bool l_isk = lc < KCH;
  const bool l_zero = l_isk && lc * 8 >= KDP;                       // padded key columns: zero fill
  const __nv_bfloat16* l_row0 = base + (l_isk ? koff + (l_zero ? 0 : lc * 8) : voff + (lc - KCH) * 8);
  const uint32_t l_dst0 = l_isk ? (uint32_t)(lr0 * KPITCH + lc * 16) : (uint32_t)(AT_KB * KPITCH + lr0 * VPITCH + (lc - KCH) * 16);
  const uint32_t l_dstep = (uint32_t)RP * (l_isk ? (uint32_t)KPITCH : (uint32_t)VPITCH);
  const long long l_sstep = (long long)RP * sCtot;

  auto load_block = [&](int blk, int stage) {
    const int m0 = blk * AT_KB;
    if (l_on) {
      const __nv_bfloat16* src = l_row0 + (long long)(m0 + lr0) * sCtot;
      uint32_t dst = sm_u + (uint32_t)(stage * STAGE) + l_dst0;
      if (m0 + AT_KB <= N) {            // full block: no row checks
#pragma unroll
        for (int u = 0; u < PASSES; ++u) {
          if ((u + 1) * RP <= AT_KB || lr0 + u * RP < AT_KB) cp_async16(dst, src, !l_zero);
          src += l_sstep; dst += l_dstep;
        }
      } else {                          // last block of a map whose token count is not a multiple of AT_KB: rows past N are zeros
#pragma unroll
        for (int u = 0; u < PASSES; ++u) {
          const int r = lr0 + u * RP;
          if ((u + 1) * RP <= AT_KB || r < AT_KB) {
            const bool rok = m0 + r < N;
            cp_async16(dst, rok ? src : l_row0, rok && !l_zero);
          }
          src += l_sstep; dst += l_dstep;
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  const int nblk = (N + AT_KB - 1) / AT_KB;
  load_block(0, 0);

  // Q fragments (A operand), straight from global memory: rows g / g+8 of this warp's 16 queries
  const int n0 = (blockIdx.x * AT_WARPS + warp) * 16;
  uint32_t qa[KS][4];
  {
    const int r0 = n0 + g, r1 = n0 + g + 8;
    const __nv_bfloat16* q0 = base + (long long)(r0 < N ? r0 : 0) * sCtot + qoff;
    const __nv_bfloat16* q1 = base + (long long)(r1 < N ? r1 : 0) * sCtot + qoff;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int c0 = ks * 16 + 2 * t, c1 = c0 + 8;
      qa[ks][0] = (r0 < N && c0 < KDP) ? *reinterpret_cast<const uint32_t*>(q0 + c0) : 0u;
      qa[ks][1] = (r1 < N && c0 < KDP) ? *reinterpret_cast<const uint32_t*>(q1 + c0) : 0u;
      qa[ks][2] = (r0 < N && c1 < KDP) ? *reinterpret_cast<const uint32_t*>(q0 + c1) : 0u;
      qa[ks][3] = (r1 < N && c1 < KDP) ? *reinterpret_cast<const uint32_t*>(q1 + c1) : 0u;
    }
  }
  float o[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m0r = -INFINITY, m1r = -INFINITY, l0 = 0.f, l1 = 0.f;   // running max (log2 domain) / partial sums, rows g and g+8

  for (int blk = 0; blk < nblk; ++blk) {
    if (blk + 1 < nblk) {
      load_block(blk + 1, (blk + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const uint32_t vs_u = sm_u + (blk & 1) * STAGE + AT_KB * KPITCH;

    float s[AT_KB / 8][4];
    {
      // K fragments through ldmatrix: matrix i of an x4 = rows j*8 .. j*8+7, bytes 16*i .. 16*i+15 of the key rows
      // (= b0 / b1 of k-step i/2); the padded pitch keeps the eight 16-byte rows of a matrix on distinct banks
      const uint32_t krow = sm_u + (uint32_t)((blk & 1) * STAGE) + (uint32_t)(lane & 7) * KPITCH;
#pragma unroll
      for (int j = 0; j < AT_KB / 8; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        const uint32_t kj = krow + (uint32_t)(j * 8) * KPITCH;
#pragma unroll
        for (int ks = 0; ks + 1 < KS; ks += 2) {
          uint32_t b0, b1, b2, b3;
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                       : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                       : "r"(kj + (uint32_t)(ks * 32 + (lane >> 3) * 16)));
          mma_bf16_16816(s[j], qa[ks], b0, b1);
          mma_bf16_16816(s[j], qa[ks + 1], b2, b3);
        }
        if (KS & 1) {
          uint32_t b0, b1;
          asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];"
                       : "=r"(b0), "=r"(b1)
                       : "r"(kj + (uint32_t)((KS - 1) * 32 + ((lane >> 3) & 1) * 16)));
          mma_bf16_16816(s[j], qa[KS - 1], b0, b1);
        }
      }
    }
    // mask keys past N (last block only), block row max of the RAW scores (scale > 0), then one FFMA per score into
    // the log2 domain: p = 2^(s * scale - max * scale)
    if ((blk + 1) * AT_KB > N) {
      const int kbase = blk * AT_KB + 2 * t;
#pragma unroll
      for (int j = 0; j < AT_KB / 8; ++j) {
        const int key = kbase + j * 8;
        if (key >= N) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
        if (key + 1 >= N) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      }
    }
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < AT_KB / 8; ++j) {
      bm0 = fmaxf(bm0, fmaxf(s[j][0], s[j][1]));
      bm1 = fmaxf(bm1, fmaxf(s[j][2], s[j][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    const float nm0 = fmaxf(m0r, bm0 * scale_log2e), nm1 = fmaxf(m1r, bm1 * scale_log2e);   // finite: every block holds >= 1 valid key
    if (__any_sync(0xffffffffu, nm0 != m0r || nm1 != m1r)) {     // the running max moved for some row of this warp: rescale
      const float c0 = ex2(m0r - nm0), c1 = ex2(m1r - nm1);      // first block: ex2(-inf) = 0
      l0 *= c0; l1 *= c1;
#pragma unroll
      for (int i = 0; i < NT; ++i) { o[i][0] *= c0; o[i][1] *= c0; o[i][2] *= c1; o[i][3] *= c1; }
      m0r = nm0; m1r = nm1;
    }
#pragma unroll
    for (int j = 0; j < AT_KB / 8; ++j) {
      s[j][0] = ex2(fmaf(s[j][0], scale_log2e, -nm0)); s[j][1] = ex2(fmaf(s[j][1], scale_log2e, -nm0));
      s[j][2] = ex2(fmaf(s[j][2], scale_log2e, -nm1)); s[j][3] = ex2(fmaf(s[j][3], scale_log2e, -nm1));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    // O += P V : P (bf16) re-packed from the S fragments; V^T fragments through ldmatrix.trans
#pragma unroll
    for (int kk = 0; kk < AT_KB / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      const uint32_t vrow = vs_u + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * VPITCH;
#pragma unroll
      for (int dt = 0; dt + 1 < NT; dt += 2) {
        uint32_t b0, b1, b2, b3;
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                     : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3)
                     : "r"(vrow + (dt * 8 + (lane >> 4) * 8) * 2));
        mma_bf16_16816(o[dt], pa, b0, b1);
        mma_bf16_16816(o[dt + 1], pa, b2, b3);
      }
      if (NT & 1) {
        uint32_t b0, b1;
        asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
                     : "=r"(b0), "=r"(b1)
                     : "r"(vrow + (NT - 1) * 16));
        mma_bf16_16816(o[NT - 1], pa, b0, b1);
      }
    }
    __syncthreads();   // everyone is done with this stage before it is refilled
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  const int r0 = n0 + g, r1 = n0 + g + 8;
  __nv_bfloat16* o0 = out + ((long long)b * N + r0) * dCtot + dC0 + h * HD + 2 * t;
  __nv_bfloat16* o1 = out + ((long long)b * N + r1) * dCtot + dC0 + h * HD + 2 * t;
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    if (r0 < N) *reinterpret_cast<uint32_t*>(o0 + i * 8) = pack_bf16(o[i][0] * i0, o[i][1] * i0);
    if (r1 < N) *reinterpret_cast<uint32_t*>(o1 + i * 8) = pack_bf16(o[i][2] * i1, o[i][3] * i1);
  }
}

template <int KDP, int HD>
int32_t run_mma(const ly_op& op, cudaStream_t s) {
  const int N = op.src.H * op.src.W;
  dim3 grid((N + AT_WARPS * 16 - 1) / (AT_WARPS * 16), op.nh, op.B);
  launch_k(attn_mma_kernel<KDP, HD>, grid, dim3(AT_WARPS * 32), 0, s, (const __nv_bfloat16*)op.src.ptr, N, op.src.ctot, op.src.c0, op.nh,
                                                           (__nv_bfloat16*)op.dst.ptr, op.dst.ctot, op.dst.c0,
                                                           op.scale * 1.4426950408889634f);
  return post_launch("psa_attention_mma");
}

template <typename T, int KDP, int HD>
int32_t run(const ly_op& op, cudaStream_t s) {
  const int N = op.src.H * op.src.W;
  dim3 grid((N + QB - 1) / QB, op.nh, op.B);
  launch_k(attn_kernel<T, KDP, HD>, grid, dim3(QB), 0, s, (const T*)op.src.ptr, N, op.src.ctot, op.src.c0, op.nh, (T*)op.dst.ptr,
                                              op.dst.ctot, op.dst.c0, op.scale);
  return post_launch("psa_attention");
}

}  // namespace

int32_t launch_attn(const ly_op& op, cudaStream_t s) {
  LY_CHECK_ARG(op.src.ptr && op.dst.ptr, "attention: null pointer");
  LY_CHECK_ARG(op.src.c >= 2 * op.nh * op.kdp + op.nh * op.hd && op.dst.c >= op.nh * op.hd, "attention: channel mismatch");
  LY_CHECK_ARG(op.src.c0 % 8 == 0 && op.dst.c0 % 8 == 0 && op.src.ctot % 8 == 0 && op.dst.ctot % 8 == 0,
               "attention: views must be 16-byte aligned");
  const bool f32 = op.dtype == LY_F32;
  const bool simt = f32 || op.impl == LY_IMPL_SIMT;
  if (op.kdp == 32 && op.hd == 64)
    return f32 ? run<float, 32, 64>(op, s) : (simt ? run<__nv_bfloat16, 32, 64>(op, s) : run_mma<32, 64>(op, s));
  if (op.kdp == 40 && op.hd == 72)
    return f32 ? run<float, 40, 72>(op, s) : (simt ? run<__nv_bfloat16, 40, 72>(op, s) : run_mma<40, 72>(op, s));
  if (op.kdp == 16 && op.hd == 32)
    return f32 ? run<float, 16, 32>(op, s) : (simt ? run<__nv_bfloat16, 16, 32>(op, s) : run_mma<16, 32>(op, s));
  LY_CHECK_ARG(false, "attention: unsupported (key_dim_pad=%d, head_dim=%d); built: (32,64) (40,72) (16,32)", op.kdp, op.hd);
}

}  // namespace ly
