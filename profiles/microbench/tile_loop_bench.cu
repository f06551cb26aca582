// Microbenchmark: the inner loop of the pair-lane reverse tile kernel (k_tile_bwd_soa) on a
// shared-memory-resident tile, with its ingredients switched on one at a time, to find which
// of them keeps the FP32 pipe below its 128 FMA/clk/SM:
//   MODE 0  math only (operands stay in registers)              -> FFMA2 issue ceiling of this mix
//   MODE 1  + 8 LDS.128 / 8 STS.128 per item                    -> shared-memory traffic
//   MODE 2  + one __syncthreads per gate (4 items per thread)   -> barrier
//   MODE 3  + lane fold and warp reduce-scatter per gate        -> the gradient flush
// Cycles come from clock64 (per CTA, averaged) so the result does not depend on the clock the
// power cap allows; wall time gives the effective MHz.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tile_loop_bench.cu -o tile_loop_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
struct V4 { float2 re, im; };
struct Mat { float re[16], im[16]; };
struct Params { int ngates, hv, lv; Mat inv[8], tr[8]; };
__device__ __forceinline__ float2 bc2(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ V4 ld4(const float4* p) { float4 t = *p; V4 v; v.re = make_float2(t.x, t.y); v.im = make_float2(t.z, t.w); return v; }
__device__ __forceinline__ void st4(float4* p, const V4& v) { *p = make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }

__device__ __forceinline__ void mv(const Mat& G, const V4 (&a)[4], V4 (&o)[4]) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    o[r].re = fmul2(bc2(G.re[4 * r]), a[0].re);
    o[r].im = fmul2(bc2(G.re[4 * r]), a[0].im);
    o[r].re = ffma2(bc2(-G.im[4 * r]), a[0].im, o[r].re);
    o[r].im = ffma2(bc2(G.im[4 * r]), a[0].re, o[r].im);
#pragma unroll
    for (int c = 1; c < 4; c++) {
      o[r].re = ffma2(bc2(G.re[4 * r + c]), a[c].re, o[r].re);
      o[r].im = ffma2(bc2(G.re[4 * r + c]), a[c].im, o[r].im);
      o[r].re = ffma2(bc2(-G.im[4 * r + c]), a[c].im, o[r].re);
      o[r].im = ffma2(bc2(G.im[4 * r + c]), a[c].re, o[r].im);
    }
  }
}
__device__ __forceinline__ void outer(const V4 (&b)[4], const V4 (&a)[4], float2 (&are)[16], float2 (&aim)[16]) {
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const float2 nbi = neg2(b[p].im);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      are[4 * p + q] = ffma2(b[p].re, a[q].re, are[4 * p + q]);
      are[4 * p + q] = ffma2(nbi, a[q].im, are[4 * p + q]);
      aim[4 * p + q] = ffma2(b[p].re, a[q].im, aim[4 * p + q]);
      aim[4 * p + q] = ffma2(b[p].im, a[q].re, aim[4 * p + q]);
    }
  }
}
__device__ __forceinline__ uint32_t ins0(uint32_t i, int pos) { return ((i >> pos) << (pos + 1)) | (i & ((1u << pos) - 1u)); }

// reduce-scatter of 32 values over the warp: lane j ends with the total of value j
template <int N, int OFF>
struct Halve {
  static __device__ __forceinline__ void run(float (&v)[32], int lane) {
    const bool up = lane & OFF;
#pragma unroll
    for (int j = 0; j < N / 2; j++) {
      const float send = up ? v[j] : v[j + N / 2];
      const float keep = up ? v[j + N / 2] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
    }
    Halve<N / 2, OFF / 2>::run(v, lane);
  }
};
template <int OFF>
struct Halve<1, OFF> { static __device__ __forceinline__ void run(float (&)[32], int) {} };

template <int MODE>
__global__ void __launch_bounds__(128, 3) k(float4* io, const __grid_constant__ Params p, int reps, float* sink, long long* cyc) {
  extern __shared__ float4 sm[];
  float4* smf = sm;
  float4* smb = sm + 2048;
  for (int i = threadIdx.x; i < 4096; i += 128) sm[i] = io[(size_t)blockIdx.x * 4096 + i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float tot = 0.f;
  const long long t0 = clock64();
  V4 f[4], b[4];
  if (MODE == 0) {
#pragma unroll
    for (int c = 0; c < 4; c++) { f[c] = ld4(smf + threadIdx.x * 4 + c); b[c] = ld4(smb + threadIdx.x * 4 + c); }
  }
  for (int rep = 0; rep < reps; rep++) {
    for (int g = 0; g < p.ngates; g++) {
      float2 are[16], aim[16];
#pragma unroll
      for (int k2 = 0; k2 < 16; k2++) are[k2] = aim[k2] = make_float2(0.f, 0.f);
      for (int i0 = 0; i0 < 512; i0 += 128) {
        const uint32_t base = ins0(ins0(i0 + threadIdx.x, p.lv), p.hv);
        V4 a[4], bo[4];
        if (MODE >= 1) {
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const uint32_t off = ((uint32_t)(c >> 1) << p.hv) + ((uint32_t)(c & 1) << p.lv);
            f[c] = ld4(smf + base + off);
            b[c] = ld4(smb + base + off);
          }
        }
        mv(p.inv[g], f, a);
        outer(b, a, are, aim);
        mv(p.tr[g], b, bo);
        if (MODE >= 1) {
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const uint32_t off = ((uint32_t)(c >> 1) << p.hv) + ((uint32_t)(c & 1) << p.lv);
            st4(smf + base + off, a[c]);
            st4(smb + base + off, bo[c]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 4; c++) { f[c] = a[c]; b[c] = bo[c]; }
        }
      }
      float acc[32];
#pragma unroll
      for (int k2 = 0; k2 < 16; k2++) { acc[2 * k2] = are[k2].x + are[k2].y; acc[2 * k2 + 1] = aim[k2].x + aim[k2].y; }
      if (MODE >= 3) {
        Halve<32, 16>::run(acc, lane);
        tot += acc[0];
      } else {
#pragma unroll
        for (int k2 = 0; k2 < 32; k2++) tot += acc[k2];
      }
      if (MODE >= 2) __syncthreads();
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (MODE == 0) tot += f[0].re.x + b[0].re.x;
  if (tot == 12345.678f) sink[0] = tot;
  __syncthreads();
  for (int i = threadIdx.x; i < 4096; i += 128) io[(size_t)blockIdx.x * 4096 + i] = sm[i];
}

template <int MODE>
void run(const char* name, float4* io, const Params& p, float* sink, long long* cyc, int sms, int reps) {
  const int grid = sms * 3;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  int bps = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k<MODE>, 128, 65536);
  for (int it = 0; it < 2; it++) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<grid, 128, 65536>>>(io, p, reps, sink, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long* h = new long long[grid];
    cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; i++) avg += (double)h[i]; avg /= grid;
    delete[] h;
    // per CTA: reps * ngates * 512 items * 8 amplitudes * 48 FMA ; 3 CTAs per SM share the SM for `avg` cycles
    const double fma_per_cta = (double)reps * p.ngates * 512 * 8 * 48;
    if (it == 1)
      printf("%-44s CTAs/SM %d  %.3f ms  %.0f cycles/CTA  -> %.1f FMA/clk/SM (%.0f%% of 128), effective %.0f MHz\n", name, bps, ms,
             avg, fma_per_cta * bps / avg, 100.0 * fma_per_cta * bps / avg / 128.0, avg / (ms * 1e3));
  }
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * 3;
  float4* io; cudaMalloc(&io, (size_t)grid * 4096 * sizeof(float4));
  cudaMemset(io, 0, (size_t)grid * 4096 * sizeof(float4));
  float* sink; cudaMalloc(&sink, 4);
  long long* cyc; cudaMalloc(&cyc, grid * sizeof(long long));
  Params p; p.ngates = 7; p.hv = 7; p.lv = 3;
  for (int g = 0; g < 8; g++) for (int i = 0; i < 16; i++) {
    p.inv[g].re[i] = (i % 5 == 0) ? 0.7f : 0.01f * (i + g); p.inv[g].im[i] = 0.02f * (i - g);
    p.tr[g].re[i] = (i % 5 == 0) ? 0.7f : -0.01f * (i + g); p.tr[g].im[i] = -0.02f * (i - g);
  }
  const int reps = 200;
  run<0>("0 math only", io, p, sink, cyc, sms, reps);
  run<1>("1 + LDS/STS", io, p, sink, cyc, sms, reps);
  run<2>("2 + barrier per gate", io, p, sink, cyc, sms, reps);
  run<3>("3 + fold and warp reduce-scatter per gate", io, p, sink, cyc, sms, reps);
  return 0;
}
