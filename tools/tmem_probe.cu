// Probe: (1) tcgen05.ld throughput with 4 / 8 / 16 warps, (2) semantics and cost of tcgen05.shift.down.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe.bin tools/tmem_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr) : "memory");
}

// out: [128 lanes][128 cols] after the experiment; info[0..]: timings
__global__ void __launch_bounds__(512, 1) probe(int nshift, uint32_t shift_col, int ld_warps, int ld_iters, uint32_t* out, long long* info) {
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int q = warp & 3;
  // ---- fill columns [0, 128): value = lane * 1000 + col
  if (warp < 4) {
    for (int c0 = 0; c0 < 128; c0 += 16) {
      uint32_t r[16];
      for (int j = 0; j < 16; ++j) r[j] = (uint32_t)((q * 32 + lane) * 1000 + c0 + j);
      tmem_st16(tmem + ((uint32_t)(q * 32) << 16) + c0, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- shift
  if (threadIdx.x == 0 && nshift > 0) {
    const long long t0 = clock64();
    for (int i = 0; i < nshift; ++i)
      asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(tmem + shift_col) : "memory");
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    info[0] = t1 - t0; info[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- read back
  if (warp < 4) {
    for (int c0 = 0; c0 < 128; c0 += 16) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) out[(q * 32 + lane) * 128 + c0 + j] = r[j];
    }
  }
  __syncthreads();
  // ---- LDTM throughput: ld_warps warps, each ld_iters x (4 loads + wait)
  long long t0 = clock64();
  uint32_t acc = 0;
  if (warp < ld_warps) {
    for (int it = 0; it < ld_iters; ++it) {
      uint32_t r0[16], r1[16], r2[16], r3[16];
      const uint32_t base = tmem + ((uint32_t)(q * 32) << 16) + ((it * 64) & 255);
      tmem_ld16(base, r0); tmem_ld16(base + 16, r1); tmem_ld16(base + 32, r2); tmem_ld16(base + 48, r3);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += r0[0] ^ r1[3] ^ r2[7] ^ r3[15];
    }
  }
  long long t1 = clock64();
  if (acc == 0x12345u) out[0] = acc;
  __syncthreads();
  if (threadIdx.x == 0) info[2] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  uint32_t* d; long long* info;
  cudaMalloc(&d, 128 * 128 * 4); cudaMalloc(&info, 64);
  static uint32_t h[128 * 128];
  long long hi[8];
  struct { int n; uint32_t col; } cases[] = {{0, 0}, {1, 0}, {2, 0}, {1, 8}, {1, 40}, {1, (32u << 16)}, {1, (32u << 16) + 16}};
  for (auto c : cases) {
    cudaMemset(info, 0, 64);
    probe<<<1, 512>>>(c.n, c.col, 4, 16, d, info);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(hi, info, 64, cudaMemcpyDeviceToHost);
    int changed = 0, cmin = 999, cmax = -1, lmin = 999, lmax = -1;
    long long dl = 0; int ndl = 0;
    for (int l = 0; l < 128; ++l)
      for (int cc = 0; cc < 128; ++cc) {
        const uint32_t v = h[l * 128 + cc], exp = (uint32_t)(l * 1000 + cc);
        if (v != exp) {
          ++changed; if (cc < cmin) cmin = cc; if (cc > cmax) cmax = cc; if (l < lmin) lmin = l; if (l > lmax) lmax = l;
          if (v % 1000 == (uint32_t)cc) { dl += (long long)l - (long long)(v / 1000); ++ndl; }
        }
      }
    printf("shift n=%d taddr_off=0x%x: changed %d cells, cols [%d,%d], lanes [%d,%d], mean lane delta %.2f over %d; issue %lld cyc, issue+complete %lld cyc\n",
           c.n, c.col, changed, cmin, cmax, lmin, lmax, ndl ? (double)dl / ndl : 0.0, ndl, hi[0], hi[1]);
    if (c.n == 1 && c.col == 0) {
      printf("  lane 0 cols 0..3: %u %u %u %u | lane 1: %u %u | lane 31: %u lane 32: %u lane 33: %u | lane 127: %u\n", h[0], h[1], h[2], h[3], h[128], h[129], h[31 * 128], h[32 * 128], h[33 * 128], h[127 * 128]);
    }
  }
  for (int w : {4, 8, 16}) {
    probe<<<1, 512>>>(0, 0, w, 2000, d, info);
    cudaDeviceSynchronize();
    cudaMemcpy(hi, info, 64, cudaMemcpyDeviceToHost);
    const double bytes = (double)w * 2000 * 4 * 16 * 32 * 4;
    printf("LDTM x16: %2d warps: %lld cycles for %.0f bytes = %.1f B/cycle/SM\n", w, hi[2], bytes, bytes / hi[2]);
  }
  return 0;
}
