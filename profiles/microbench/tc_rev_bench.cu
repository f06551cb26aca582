// Fused tensor-core reverse step of a block (csrc/tc_rev.cuh: k_tc_block_rev): state <- W^dagger state,
// adjoint <- W^T adjoint, H += adjoint (x) state in ONE sweep; against a double-precision host evaluation, and its
// throughput against the HBM roofline (4 * S of traffic).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tc_rev_bench tc_rev_bench.cu
//   ./tc_rev_bench [n_check=22] [n_time=30] [first block qubit=8] [block list or -] [products=8] [h_products=6]
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../../differentiable-quantum-circuit-cuda_b200/csrc/tc_rev.cuh"

typedef std::complex<double> zc;
#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) {                                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);             \
      exit(2);                                                                                    \
    }                                                                                             \
  } while (0)

static std::mt19937_64 rng(4242);

static void haar4(zc* u) {
  std::normal_distribution<double> nd;
  zc a[4][4];
  for (auto& r : a) for (auto& x : r) x = zc(nd(rng), nd(rng));
  for (int c = 0; c < 4; c++) {
    for (int k = 0; k < c; k++) {
      zc dot = 0;
      for (int r = 0; r < 4; r++) dot += std::conj(a[r][k]) * a[r][c];
      for (int r = 0; r < 4; r++) a[r][c] -= dot * a[r][k];
    }
    double nrm = 0;
    for (int r = 0; r < 4; r++) nrm += std::norm(a[r][c]);
    nrm = std::sqrt(nrm);
    for (int r = 0; r < 4; r++) a[r][c] /= nrm;
  }
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) u[4 * r + c] = a[r][c];
}

static void apply_gate_to_rows(std::vector<zc>& w, const zc* g, int hi, int lo) {
  std::vector<zc> out(w.size());
  for (int i = 0; i < 64; i++) {
    const int bi = 2 * ((i >> hi) & 1) + ((i >> lo) & 1);
    for (int j = 0; j < 64; j++) {
      zc s = 0;
      for (int b = 0; b < 4; b++) {
        const int src = (i & ~((1 << hi) | (1 << lo))) | (((b >> 1) & 1) << hi) | ((b & 1) << lo);
        s += g[4 * bi + b] * w[src * 64 + j];
      }
      out[i * 64 + j] = s;
    }
  }
  w.swap(out);
}

// timing runs: pseudo-random amplitudes generated on the device (the host generator needs ~30 s at 30 qubits)
__global__ void k_fill_random(float2* x, size_t n, float scale, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u ^ (uint32_t)(i >> 32) * 40503u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    const float a = ((h & 0xffff) - 32768.0f) / 32768.0f, b = ((h >> 16) - 32768.0f) / 32768.0f;
    x[i] = make_float2(a * scale, b * scale);
  }
}

static int g_h_products = 6;
static void run(int n, const int* block, int products, bool check) {
  tcb::RevParams rp;
  int wbit[6];
  const char* err = tcb::make_params(block, n, &rp.geo, wbit);
  if (err) { printf("make_params: %s\n", err); exit(1); }
  auto perm = [&](int idx) {   // kernel index -> caller index
    int o = 0;
    for (int k = 0; k < 6; k++) o |= ((idx >> k) & 1) << wbit[k];
    return o;
  };
  // W = brickwork diamond of 9 gates on the 6 block qubits (caller bit order)
  std::vector<zc> w(64 * 64, 0.0);
  for (int i = 0; i < 64; i++) w[i * 64 + i] = 1.0;
  const int pairs[9][2] = {{1, 0}, {3, 2}, {5, 4}, {2, 1}, {4, 3}, {1, 0}, {3, 2}, {5, 4}, {3, 2}};
  for (auto& pr : pairs) {
    zc g[16];
    haar4(g);
    apply_gate_to_rows(w, g, pr[0], pr[1]);
  }
  std::vector<double> flat(64 * 64 * 2);
  for (int i = 0; i < 64; i++)
    for (int j = 0; j < 64; j++) {
      const zc v = std::conj(w[perm(j) * 64 + perm(i)]);   // W^dagger in kernel bit order
      flat[2 * (i * 64 + j)] = v.real();
      flat[2 * (i * 64 + j) + 1] = v.imag();
    }
  const std::vector<uint32_t> img = tcb::make_w_image(flat.data());
  uint32_t* d_img;
  CK(cudaMalloc(&d_img, img.size() * 4));
  CK(cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
  int* d_err;
  CK(cudaMalloc(&d_err, sizeof(int)));
  CK(cudaMemset(d_err, 0, sizeof(int)));
  rp.geo.error_flag = d_err;
  rp.geo.w_image = d_img;
  rp.geo.products = products;
  rp.h_products = g_h_products;
  const size_t N = (size_t)1 << n;
  std::vector<float2> hx(check ? N : 0), hy(check ? N : 0);
  if (check) {
    std::normal_distribution<float> nd;
    const float sc = 1.0f / std::sqrt((float)N);
    for (size_t i = 0; i < N; i++) { hx[i] = make_float2(nd(rng) * sc, nd(rng) * sc); hy[i] = make_float2(nd(rng) * sc + sc, nd(rng) * sc); }
  }
  float2 *dx, *dy;
  CK(cudaMalloc(&dx, N * sizeof(float2)));
  CK(cudaMalloc(&dy, N * sizeof(float2)));
  if (check) {
    CK(cudaMemcpy(dx, hx.data(), N * sizeof(float2), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dy, hy.data(), N * sizeof(float2), cudaMemcpyHostToDevice));
  } else {
    const float sc = 1.0f / std::sqrt((float)N);
    k_fill_random<<<1184, 256>>>(dx, N, sc, 17u);
    k_fill_random<<<1184, 256>>>(dy, N, sc, 99u);
    CK(cudaDeviceSynchronize());
  }
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = (int)std::min<uint64_t>(rp.geo.ntiles, (uint64_t)sms);
  const size_t pbytes = (size_t)grid * 128 * 128 * sizeof(float);
  CK(cudaMalloc(&rp.partials, pbytes));
  double* d_out;
  CK(cudaMalloc(&d_out, 128 * 128 * sizeof(double)));
  CK(cudaFuncSetAttribute(tcb::k_tc_block_rev, cudaFuncAttributeMaxDynamicSharedMemorySize, tcb::kRevSmemBytes));
#ifdef TC_REV_TRACE
  CK(cudaMalloc(&rp.trace, 8 * 32 * sizeof(long long)));
  CK(cudaMemset(rp.trace, 0, 8 * 32 * sizeof(long long)));
#endif
  auto launch = [&]() {
    cudaMemsetAsync(rp.partials, 0, pbytes);
    tcb::k_tc_block_rev<<<grid, tcb::kRevThreads, tcb::kRevSmemBytes>>>(dx, dy, rp);
    tcb::k_tc_grad_reduce<<<(128 * 128 + 255) / 256, 256>>>(rp.partials, grid, d_out, 0);
  };
  launch();
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  if (check) {
    std::vector<double> P(128 * 128);
    CK(cudaMemcpy(P.data(), d_out, P.size() * sizeof(double), cudaMemcpyDeviceToHost));
    std::vector<float2> ox(N), oy(N);
    CK(cudaMemcpy(ox.data(), dx, N * sizeof(float2), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(oy.data(), dy, N * sizeof(float2), cudaMemcpyDeviceToHost));
    std::vector<zc> H(64 * 64, 0.0);
    uint64_t bmask = 0;
    for (int b = 0; b < 6; b++) bmask |= 1ull << block[b];
    double ex = 0, ey = 0, mxv = 0, myv = 0;
    size_t groups = 0;
    for (size_t base = 0; base < N; base++) {
      if (base & bmask) continue;
      zc x[64], y[64];
      size_t idx[64];
      for (int j = 0; j < 64; j++) {
        size_t id = base;
        for (int b = 0; b < 6; b++) id |= (size_t)((j >> b) & 1) << block[b];
        idx[j] = id;
        x[j] = zc(hx[id].x, hx[id].y);
        y[j] = zc(hy[id].x, hy[id].y);
      }
      for (int i = 0; i < 64; i++)
        for (int j = 0; j < 64; j++) H[i * 64 + j] += y[i] * x[j];
      if ((groups++ & 63) == 0) {   // sampled groups: the two block products
        for (int i = 0; i < 64; i++) {
          zc sx = 0, sy = 0;
          for (int j = 0; j < 64; j++) {
            sx += std::conj(w[j * 64 + i]) * x[j];   // (W^dagger x)_i
            sy += w[j * 64 + i] * y[j];              // (W^T y)_i
          }
          ex = std::max(ex, std::abs(sx - zc(ox[idx[i]].x, ox[idx[i]].y)));
          ey = std::max(ey, std::abs(sy - zc(oy[idx[i]].x, oy[idx[i]].y)));
          mxv = std::max(mxv, std::abs(sx));
          myv = std::max(myv, std::abs(sy));
        }
      }
    }
    double eh = 0, mh = 0;
    for (int i = 0; i < 64; i++)
      for (int k = 0; k < 64; k++) {
        const zc got(P[i * 128 + k] + P[(64 + i) * 128 + 64 + k], P[i * 128 + 64 + k] - P[(64 + i) * 128 + k]);
        const zc want = H[perm(i) * 64 + perm(k)];
        eh = std::max(eh, std::abs(got - want));
        mh = std::max(mh, std::abs(want));
      }
    printf("h_products=%d ", g_h_products);
    printf("n=%d block=%d,%d,%d,%d,%d,%d products=%d vs host double, max |err| / max |value|: state %.3e  adjoint %.3e  H %.3e (max entry %.3e)\n",
           n, block[0], block[1], block[2], block[3], block[4], block[5], products, ex / mxv, ey / myv, eh / mh, mh);
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
    const double mmas = (products >= 8 ? 128.0 : 96.0) * 2 * 128 * 64 * 16 + (g_h_products >= 6 ? 24.0 : 12.0) * 2 * 128 * 128 * 16;
    printf("h_products=%d ", g_h_products);
    printf("n=%d products=%d fused reverse step: %.3f ms = %.1f GB/s of HBM traffic (4 * S), %.1f TFLOP/s bf16 tensor\n", n, products, ms,
           4.0 * N * sizeof(float2) / ms * 1e-6, mmas * (double)rp.geo.ntiles / ms * 1e-9);
  }
#ifdef TC_REV_TRACE
  if (!check) {
    // timeline of CTA 0, tiles 8..15 (SM cycles relative to the fill start of tile 8); slots: see tc_rev.cuh
    long long tr[8 * 32];
    CK(cudaMemcpy(tr, rp.trace, sizeof(tr), cudaMemcpyDeviceToHost));
    const char* names[32] = {"fill:start", "fill:x loaded+max", "fill:empty ok", "fill:x sliced", "fill:y max", "fill:y sliced", "fill:fenced", "",
                             "mma:full ok", "mma:H_a issued", "mma:acc free(Y')", "mma:X' issued", "mma:H_b issued", "mma:acc free(X')", "mma:Y' issued", "",
                             "drain:A", "drain:X' in regs", "drain:B", "drain:X' staged(+flush)", "drain:bar", "drain:X' written", "drain:C",
                             "drain:Y' in regs", "drain:Y' staged+bar", "drain:Y' written", "drain:done", "", "", "", "", ""};
    const long long t0 = tr[0];
    printf("%-26s", "slot \\ tile");
    for (int k = 0; k < 8; k++) printf("%9d", 8 + k);
    printf("\n");
    for (int sl = 0; sl < 27; sl++) {
      if (!names[sl][0]) continue;
      printf("%-26s", names[sl]);
      for (int k = 0; k < 8; k++) printf("%9lld", tr[k * 32 + sl] - t0);
      printf("\n");
    }
  }
#endif
  int herr = 0;
  CK(cudaMemcpy(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost));
  if (herr) printf("watchdog flag: %d\n", herr);
  cudaFree(dx); cudaFree(dy); cudaFree(rp.partials); cudaFree(d_out); cudaFree(d_err); cudaFree(d_img);
}

int main(int argc, char** argv) {
  const int n_check = argc > 1 ? atoi(argv[1]) : 22;
  const int n_time = argc > 2 ? atoi(argv[2]) : 30;
  const int q0 = argc > 3 ? atoi(argv[3]) : 8;
  int block[6];
  for (int b = 0; b < 6; b++) block[b] = q0 + b;
  if (argc > 4 && strcmp(argv[4], "-") != 0) {
    if (sscanf(argv[4], "%d,%d,%d,%d,%d,%d", block, block + 1, block + 2, block + 3, block + 4, block + 5) != 6) return 1;
  }
  const int products = argc > 5 ? atoi(argv[5]) : 8;
  g_h_products = argc > 6 ? atoi(argv[6]) : 6;
  if (n_check > 0) run(n_check, block, products, true);
  if (n_time > 0) run(n_time, block, products, false);
  return 0;
}
