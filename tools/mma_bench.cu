// Micro-benchmark: issue cost / throughput of tcgen05.mma (M=128, bf16, cta_group::1, SS mode)
// as a function of N and of what else the CTA is doing.  One CTA per SM; one elected thread
// issues L MMAs, commits to an mbarrier and waits; reports cycles per MMA.
//   MODE bit 3 (8) : every 4 MMAs run the real kernel's k-block protocol (commit + try_wait + fence)
//   MODE bit 5 (32): 8 other warps hammer TMEM loads + tanh + global stores (an "epilogue" load)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench.bin tools/mma_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(320, 1) bench(int N, int L, long long* out, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar, bar2, bar3;
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 60000;" ::"r"(smem_u32(&bar2)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t hi = ((uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29))) << 32;   // SBO=1024, v1, SW128
      const uint32_t a0 = base, b0 = base + 64 * 1024;
      const long long t0 = clock64();
      for (int kb = 0; kb < L / 4; ++kb) {
        const uint32_t aa = a0 + (kb % 3) * 16384, bb = b0 + (kb % 3) * 32768;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t da = hi | (uint64_t)(((aa + kk * 32) >> 4) & 0x3FFF) | (1ull << 16);
          const uint64_t db = hi | (uint64_t)(((bb + kk * 32) >> 4) & 0x3FFF) | (1ull << 16);
          umma(tmem, da, db, idesc, (kb | kk) ? 1u : 0u);
        }
        if (MODE & 8) {
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
          uint32_t dn = 0;
          while (!dn)   // bar3 never completes a phase: waiting for parity 1 succeeds immediately
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 1;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(dn) : "r"(smem_u32(&bar3)) : "memory");
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
      }
      const long long t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
      stop = 1;
    }
  } else if (warp >= 2 && (MODE & 32)) {
    const int q = warp & 3, lane = threadIdx.x & 31;
    float acc = 0.f;
    float* o = sink + ((size_t)blockIdx.x * 320 + threadIdx.x) * 64;
    int it = 0;
    while (!stop) {
      uint32_t r[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + 256 + (it & 7) * 16, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float h = 0.5f * __uint_as_float(r[j]) + (float)lane, t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
        v[j] = fmaf(h, t, h);
        acc += v[j];
      }
      *reinterpret_cast<float4*>(o + (it & 3) * 16) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(o + (it & 3) * 16 + 4) = make_float4(v[4], v[5], v[6], v[7]);
      ++it;
    }
    if (acc == 123.456f) o[0] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int MODE>
void run(int N, int L, long long* d, float* sink) {
  cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  bench<MODE><<<148, 320, 180 * 1024>>>(N, L, d, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("N %3d mode %2d  issue %7.1f cyc/MMA  total %7.1f cyc/MMA   ideal %4d\n", N, MODE, (double)h[0] / L, (double)h[1] / L, 128 * N / 256);
}

int main() {
  long long* d;
  float* sink;
  cudaMalloc(&d, 16);
  cudaMalloc(&sink, (size_t)148 * 320 * 64 * 4);
  const int L = 512;
  for (int N : {8, 16, 32, 64, 128, 256}) {
    run<0>(N, L, d, sink);
    run<8>(N, L, d, sink);
    run<32>(N, L, d, sink);
    run<40>(N, L, d, sink);
  }
  return 0;
}
