// Microbenchmark: FP32 FMA throughput per SM with scalar FFMA vs packed FFMA2
// (fma.rn.f32x2, sm_100+).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2_bench.cu -o ffma2_bench
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
template <bool PACKED>
__global__ void k(float2* out, float2 g, int iters) {
  float2 acc[8];
  float2 a = make_float2(threadIdx.x * 1e-3f, 1.f);
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = make_float2(j, -j);
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (PACKED) acc[j] = ffma2(a, g, acc[j]);
      else { acc[j].x = fmaf(a.x, g.x, acc[j].x); acc[j].y = fmaf(a.y, g.y, acc[j].y); }
    }
  }
  float2 s = make_float2(0, 0);
#pragma unroll
  for (int j = 0; j < 8; j++) { s.x += acc[j].x; s.y += acc[j].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float2* out; cudaMalloc(&out, sms * 8 * 256 * sizeof(float2));
  const int iters = 20000;
  for (int packed = 0; packed < 2; packed++) for (int rep = 0; rep < 2; rep++) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    if (packed) k<true><<<sms * 8, 256>>>(out, make_float2(1.0001f, 0.9999f), iters);
    else k<false><<<sms * 8, 256>>>(out, make_float2(1.0001f, 0.9999f), iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double fmas = (double)sms * 8 * 256 * iters * 16;
    printf("%s: %.3f ms, %.2f TFMA/s, %.1f FMA/clk/SM at nominal %d MHz\n", packed ? "FFMA2" : "FFMA ", ms,
           fmas / ms * 1e-9, fmas / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  }
  return 0;
}
