// Tensor-core block gradient (csrc/tc_block.cuh: k_tc_block_grad): G_W[i, j] = sum b[i] a[j] over every group of
// 64 amplitudes on 6 block qubits, against a double-precision host evaluation, and its throughput.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tc_grad_bench tc_grad_bench.cu
//   ./tc_grad_bench [n_check=24] [n_time=30] [first block qubit=8] [block list or -]
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../../differentiable-quantum-circuit-cuda_b200/csrc/tc_block.cuh"

typedef std::complex<double> zc;
#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) {                                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);             \
      exit(2);                                                                                    \
    }                                                                                             \
  } while (0)

static std::mt19937_64 rng(99);

static void run(int n, const int* block, bool check) {
  tcb::GradParams gp;
  int wbit[6];
  const char* err = tcb::make_params(block, n, &gp.geo, wbit);
  if (err) { printf("make_params: %s\n", err); exit(1); }
  int* d_err;
  CK(cudaMalloc(&d_err, sizeof(int)));
  CK(cudaMemset(d_err, 0, sizeof(int)));
  gp.geo.error_flag = d_err;
  gp.geo.w_image = nullptr;
  gp.geo.products = 6;
  const size_t N = (size_t)1 << n;
  std::vector<float2> ha(N), hb(N);
  {
    std::normal_distribution<float> nd;
    const float sc = 1.0f / std::sqrt((float)N);
    for (size_t i = 0; i < N; i++) { ha[i] = make_float2(nd(rng) * sc, nd(rng) * sc); hb[i] = make_float2(nd(rng) * sc + sc, nd(rng) * sc); }
  }
  float2 *da, *db;
  CK(cudaMalloc(&da, N * sizeof(float2)));
  CK(cudaMalloc(&db, N * sizeof(float2)));
  CK(cudaMemcpy(da, ha.data(), N * sizeof(float2), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), N * sizeof(float2), cudaMemcpyHostToDevice));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = (int)std::min<uint64_t>(gp.geo.ntiles, (uint64_t)sms);
  const size_t pbytes = (size_t)grid * 128 * 128 * sizeof(float);
  CK(cudaMalloc(&gp.partials, pbytes));
  double* d_out;
  CK(cudaMalloc(&d_out, 128 * 128 * sizeof(double)));
  CK(cudaFuncSetAttribute(tcb::k_tc_block_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, tcb::kGradSmemBytes));
  auto launch = [&]() {
    cudaMemsetAsync(gp.partials, 0, pbytes);
    tcb::k_tc_block_grad<<<grid, tcb::kThreads, tcb::kGradSmemBytes>>>(da, db, gp);
    tcb::k_tc_grad_reduce<<<(128 * 128 + 255) / 256, 256>>>(gp.partials, grid, d_out, 0);
  };
  launch();
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<double> P(128 * 128);
  CK(cudaMemcpy(P.data(), d_out, P.size() * sizeof(double), cudaMemcpyDeviceToHost));
  if (check) {
    // host: G[i][j] = sum_groups b[i] * a[j], indices in the CALLER's bit order (bit k <-> block[k])
    std::vector<zc> G(64 * 64, 0.0);
    uint64_t bmask = 0;
    for (int b = 0; b < 6; b++) bmask |= 1ull << block[b];
    for (size_t base = 0; base < N; base++) {
      if (base & bmask) continue;
      zc xa[64], xb[64];
      for (int j = 0; j < 64; j++) {
        size_t idx = base;
        for (int b = 0; b < 6; b++) idx |= (size_t)((j >> b) & 1) << block[b];
        xa[j] = zc(ha[idx].x, ha[idx].y);
        xb[j] = zc(hb[idx].x, hb[idx].y);
      }
      for (int i = 0; i < 64; i++)
        for (int j = 0; j < 64; j++) G[i * 64 + j] += xb[i] * xa[j];
    }
    // kernel index bit k <-> caller index bit wbit[k]
    auto perm = [&](int idx) {
      int o = 0;
      for (int k = 0; k < 6; k++) o |= ((idx >> k) & 1) << wbit[k];
      return o;
    };
    double max_err = 0, max_val = 0;
    for (int i = 0; i < 64; i++)
      for (int j = 0; j < 64; j++) {
        const zc got(P[i * 128 + j] - P[(64 + i) * 128 + 64 + j], P[i * 128 + 64 + j] + P[(64 + i) * 128 + j]);
        const zc want = G[perm(i) * 64 + perm(j)];
        max_err = std::max(max_err, std::abs(got - want));
        max_val = std::max(max_val, std::abs(want));
      }
    printf("n=%d block=%d,%d,%d,%d,%d,%d  block gradient vs host double: max |err| / max |entry| = %.3e (max entry %.3e)\n", n,
           block[0], block[1], block[2], block[3], block[4], block[5], max_err / max_val, max_val);
  } else {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch();
    CK(cudaEventRecord(e0));
    const int reps = 10;
    for (int r = 0; r < reps; r++) launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    printf("n=%d gradient pass: %.3f ms = %.1f GB/s of HBM reads (state + adjoint), %.1f TFLOP/s bf16 tensor (24 MMAs of 128x128x16 per tile)\n",
           n, ms, 2.0 * N * sizeof(float2) / ms * 1e-6, 24.0 * 2 * 128 * 128 * 16 * (double)gp.geo.ntiles / ms * 1e-9);
  }
  int herr = 0;
  CK(cudaMemcpy(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (herr) printf("watchdog flag: %d\n", herr);
  cudaFree(da); cudaFree(db); cudaFree(gp.partials); cudaFree(d_out); cudaFree(d_err);
}

int main(int argc, char** argv) {
  const int n_check = argc > 1 ? atoi(argv[1]) : 24;
  const int n_time = argc > 2 ? atoi(argv[2]) : 30;
  const int q0 = argc > 3 ? atoi(argv[3]) : 8;
  int block[6];
  for (int b = 0; b < 6; b++) block[b] = q0 + b;
  if (argc > 4 && strcmp(argv[4], "-") != 0) {
    if (sscanf(argv[4], "%d,%d,%d,%d,%d,%d", block, block + 1, block + 2, block + 3, block + 4, block + 5) != 6) return 1;
  }
  if (n_check > 0) run(n_check, block, true);
  if (n_time > 0) run(n_time, block, false);
  return 0;
}
