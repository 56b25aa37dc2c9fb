// Probe: tcgen05.ld.32x32b with a lane offset inside / across the warp's TMEM lane quarter.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%4], {%0,%1,%2,%3};" ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(taddr) : "memory");
}
__global__ void __launch_bounds__(128, 1) probe(int off, uint32_t* out) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  uint32_t r[4];
  for (int j = 0; j < 4; ++j) r[j] = (uint32_t)((warp * 32 + lane) * 1000 + j);
  tmem_st4(tmem + ((uint32_t)(warp * 32) << 16), r);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  tmem_ld4(tmem + ((uint32_t)(warp * 32 + off) << 16), r);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 4; ++j) out[(warp * 32 + lane) * 4 + j] = r[j];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 128 * 4 * 4);
  static uint32_t h[512];
  for (int off : {0, 1, 2, 16}) {
    cudaMemset(d, 0xff, 2048);
    probe<<<1, 128>>>(off, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("off %d: error %s\n", off, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, d, 2048, cudaMemcpyDeviceToHost);
    printf("off %d: thread 0 -> %u, 1 -> %u, 29 -> %u, 30 -> %u, 31 -> %u, 32 -> %u, 63 -> %u, 126 -> %u, 127 -> %u (col1 of t0: %u)\n", off, h[0], h[4], h[29 * 4], h[30 * 4],
           h[31 * 4], h[32 * 4], h[63 * 4], h[126 * 4], h[127 * 4], h[1]);
  }
  return 0;
}
