// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) issue throughput on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_bench.bin tools/ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  asm volatile("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%0, %1};\n\t"
      "fma.rn.f32x2 rc, ra, rb, rc;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "+f"(d.x), "+f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}
template <int MODE> __global__ void k(float* out, int iters, float s) {
  float2 acc[8];
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
  float2 a = make_float2(s, s * 0.5f), b = make_float2(1.0f - s, 0.25f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, a.x, b.x); acc[i].y = fmaf(acc[i].y, a.y, b.y); }
      else { float2 t = acc[i]; acc[i] = b; ffma2(acc[i], t, a); }
    }
  }
  float r = 0;
  for (int i = 0; i < 8; ++i) r += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 0.999f); else k<1><<<148 * 8, 256>>>(d, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fl = 2.0 * 16 * (double)iters * 148 * 8 * 256;
      if (rep) printf("%s: %.3f ms  %.1f TFLOP/s fp32\n", mode ? "FFMA2" : "FFMA ", ms, fl / ms / 1e9);
    }
  }
  return 0;
}
