// Microbenchmark: does one register-indexed constant load (LDC.64) per four
// packed FMAs (the instruction mix of the tile kernels' gate loop) throttle
// the FP32 pipe?  Variants: constants via LDC (indexed kernel parameter),
// via shared memory (LDS.64 broadcast), or hoisted into registers.
#include <cuda_runtime.h>
#include <cstdio>
struct P { float2 c[512]; };
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, const __grid_constant__ P p, int iters, int stride) {
  __shared__ float2 sc[512];
  for (int i = threadIdx.x; i < 512; i += 256) sc[i] = p.c[i];
  __syncthreads();
  float2 acc[4], a[4];
  for (int j = 0; j < 4; j++) { acc[j] = make_float2(j, -j); a[j] = make_float2(threadIdx.x * 1e-3f + j, 1.f); }
  float2 r[32];
  if (MODE == 2) for (int j = 0; j < 32; j++) r[j] = p.c[(j * stride) & 511];
  for (int it = 0; it < iters; it++) {
    const int base = (it * stride) & 255;   // runtime, warp-uniform
#pragma unroll
    for (int e = 0; e < 32; e++) {
      float2 g;
      if (MODE == 0) g = p.c[base + e];
      else if (MODE == 1) g = sc[base + e];
      else g = r[e];
#pragma unroll
      for (int j = 0; j < 4; j++) acc[j] = ffma2(g, a[j], acc[j]);
    }
  }
  float2 s = make_float2(0, 0);
  for (int j = 0; j < 4; j++) { s.x += acc[j].x; s.y += acc[j].y; }
  out[blockIdx.x * 256 + threadIdx.x] = s;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float2* out; cudaMalloc(&out, sms * 8 * 256 * sizeof(float2));
  P p; for (int i = 0; i < 512; i++) p.c[i] = make_float2(1.0f + 1e-6f * i, 1.0f - 1e-6f * i);
  const int iters = 4000;
  const char* names[3] = {"LDC.64 indexed param", "LDS.64 broadcast    ", "registers           "};
  for (int occ = 2; occ <= 8; occ *= 2)
  for (int mode = 0; mode < 3; mode++) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a);
      if (mode == 0) k<0><<<sms * occ, 256>>>(out, p, iters, 1);
      else if (mode == 1) k<1><<<sms * occ, 256>>>(out, p, iters, 1);
      else k<2><<<sms * occ, 256>>>(out, p, iters, 1);
      cudaEventRecord(b); cudaEventSynchronize(b);
    }
    float ms; cudaEventElapsedTime(&ms, a, b);
    double fmas = (double)sms * occ * 256 * iters * 32 * 4 * 2;
    printf("CTAs/SM %d  %s: %.3f ms  %.1f FMA/clk/SM (at 1965 MHz)\n", occ, names[mode], ms, fmas / (ms * 1e-3) / sms / 1.965e9);
  }
  return 0;
}
