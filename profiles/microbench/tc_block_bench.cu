// Tensor-core fused 6-qubit block (csrc/tc_block.cuh): correctness against a double-precision host
// evaluation, accuracy drift over many sequential blocks, and throughput against the HBM roofline.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -o tc_block_bench tc_block_bench.cu
//   ./tc_block_bench [n_qubits=28] [first block qubit=8] [rounds=10]
//
// Workload: W = product of 9 Haar-random two-qubit gates in a brickwork diamond on 6 qubits (what the scheduler's
// windows hold), state = random normalised vector.  `rounds` x (W, W^dagger) measures the drift: the state
// must come back, so the deviation after 2 * rounds blocks is pure arithmetic error.
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

static std::mt19937_64 rng(1234);

static void haar4(zc* u) {   // 4 x 4 Haar unitary by Gram-Schmidt of a Gaussian matrix
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

// W <- (gate on block index bits (hi, lo)) * W, 64 x 64 row-major
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

static bool g_odd_low = false;
static std::vector<uint32_t> image_of(const std::vector<zc>& w, const int* w_bit_of_jbit) {
  // kernel index bit k <-> caller index bit w_bit_of_jbit[k]
  auto perm = [&](int idx) {
    int o = 0;
    for (int k = 0; k < 6; k++) o |= ((idx >> k) & 1) << w_bit_of_jbit[k];
    return o;
  };
  std::vector<double> flat(64 * 64 * 2);
  for (int i = 0; i < 64; i++)
    for (int j = 0; j < 64; j++) {
      const zc v = w[perm(i) * 64 + perm(j)];
      flat[2 * (i * 64 + j)] = v.real();
      flat[2 * (i * 64 + j) + 1] = v.imag();
    }
  return tcb::make_w_image(flat.data(), g_odd_low);
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 28;
  const int q0 = argc > 2 ? atoi(argv[2]) : 8;
  const int rounds = argc > 3 ? atoi(argv[3]) : 10;
  int block[6];
  for (int b = 0; b < 6; b++) block[b] = q0 + b;
  if (argc > 4 && strcmp(argv[4], "-") != 0) {  // scattered block, e.g. "3,4,9,10,17,20"
    if (sscanf(argv[4], "%d,%d,%d,%d,%d,%d", block, block + 1, block + 2, block + 3, block + 4, block + 5) != 6) return 1;
  }
  const int products = argc > 5 ? atoi(argv[5]) : 8;   // 8 or 6 slice products per block
  g_odd_low = argc > 6 && atoi(argv[6]) != 0;
  // W = brickwork diamond of 9 gates on the 6 block qubits
  std::vector<zc> w(64 * 64, 0.0);
  for (int i = 0; i < 64; i++) w[i * 64 + i] = 1.0;
  const int pairs[9][2] = {{1, 0}, {3, 2}, {5, 4}, {2, 1}, {4, 3}, {1, 0}, {3, 2}, {5, 4}, {3, 2}};
  for (auto& pr : pairs) {
    zc g[16];
    haar4(g);
    apply_gate_to_rows(w, g, pr[0], pr[1]);
  }
  std::vector<zc> wdag(64 * 64);
  for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) wdag[i * 64 + j] = std::conj(w[j * 64 + i]);

  tcb::Params p;
  int wbit[6];
  const char* err = tcb::make_params(block, n, &p, wbit);
  if (err) { printf("make_params: %s\n", err); return 1; }
  std::vector<uint32_t> img = image_of(w, wbit), img_dag = image_of(wdag, wbit);
  uint32_t *d_img, *d_img_dag;
  p.products = products;
  int* d_err;
  CK(cudaMalloc(&d_img, img.size() * 4));
  CK(cudaMalloc(&d_img_dag, img.size() * 4));
  CK(cudaMalloc(&d_err, sizeof(int)));
  CK(cudaMemset(d_err, 0, sizeof(int)));
  CK(cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_img_dag, img_dag.data(), img.size() * 4, cudaMemcpyHostToDevice));
  p.error_flag = d_err;

  const size_t N = (size_t)1 << n;
  std::vector<float2> h(N);
  {
    std::normal_distribution<float> nd;
    double nrm = 0;
    for (size_t i = 0; i < N; i++) { h[i] = make_float2(nd(rng), nd(rng)); nrm += (double)h[i].x * h[i].x + (double)h[i].y * h[i].y; }
    const float sc = (float)(1.0 / std::sqrt(nrm));
    for (size_t i = 0; i < N; i++) { h[i].x *= sc; h[i].y *= sc; }
  }
  float2* d_state;
  CK(cudaMalloc(&d_state, N * sizeof(float2)));
  CK(cudaMemcpy(d_state, h.data(), N * sizeof(float2), cudaMemcpyHostToDevice));

  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(tcb::k_tc_block_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, tcb::kSmemBytes));
  const int grid = (int)std::min<uint64_t>(p.ntiles, (uint64_t)sms);
  auto launch = [&](const uint32_t* image) {
    p.w_image = image;
    tcb::k_tc_block_fwd<<<grid, tcb::kThreads, tcb::kSmemBytes>>>(d_state, p);
  };

  // ---- 1. one block against the host (double), on sampled groups of 64 amplitudes
  launch(d_img);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float2> out(N);
  CK(cudaMemcpy(out.data(), d_state, N * sizeof(float2), cudaMemcpyDeviceToHost));
  double max_err = 0, max_val = 0;
  std::uniform_int_distribution<size_t> pick(0, N - 1);
  uint64_t bmask = 0;
  for (int b = 0; b < 6; b++) bmask |= 1ull << block[b];
  for (int trial = 0; trial < 2000; trial++) {
    const size_t base = pick(rng) & ~bmask;
    zc x[64];
    for (int j = 0; j < 64; j++) {
      size_t idx = base;
      for (int b = 0; b < 6; b++) idx |= (size_t)((j >> b) & 1) << block[b];
      x[j] = zc(h[idx].x, h[idx].y);
    }
    for (int i = 0; i < 64; i++) {
      zc s = 0;
      for (int j = 0; j < 64; j++) s += w[i * 64 + j] * x[j];
      size_t idx = base;
      for (int b = 0; b < 6; b++) idx |= (size_t)((i >> b) & 1) << block[b];
      max_err = std::max(max_err, std::abs(s - zc(out[idx].x, out[idx].y)));
      max_val = std::max(max_val, std::abs(s));
    }
  }
  printf("n=%d block=%d,%d,%d,%d,%d,%d products=%d odd_low=%d  one block vs host double: max |err| / max |value| = %.3e\n", n, block[0], block[1],
         block[2], block[3], block[4], block[5], products, (int)g_odd_low, max_err / max_val);

  // ---- 2. drift: `rounds` DIFFERENT random blocks, then their inverses in reverse order (a forward sweep followed by
  //         its un-computation): the state must come back, the deviation is pure arithmetic error of 2 * rounds blocks
  {
    CK(cudaMemcpy(d_state, h.data(), N * sizeof(float2), cudaMemcpyHostToDevice));
    std::vector<uint32_t*> d_fw(rounds), d_bw(rounds);
    for (int r = 0; r < rounds; r++) {
      std::vector<zc> wr(64 * 64, 0.0), wrd(64 * 64);
      for (int i = 0; i < 64; i++) wr[i * 64 + i] = 1.0;
      for (auto& pr : pairs) {
        zc g[16];
        haar4(g);
        apply_gate_to_rows(wr, g, pr[0], pr[1]);
      }
      for (int i = 0; i < 64; i++) for (int j = 0; j < 64; j++) wrd[i * 64 + j] = std::conj(wr[j * 64 + i]);
      std::vector<uint32_t> a = image_of(wr, wbit), bimg = image_of(wrd, wbit);
      CK(cudaMalloc(&d_fw[r], a.size() * 4));
      CK(cudaMalloc(&d_bw[r], a.size() * 4));
      CK(cudaMemcpy(d_fw[r], a.data(), a.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(d_bw[r], bimg.data(), a.size() * 4, cudaMemcpyHostToDevice));
    }
    for (int r = 0; r < rounds; r++) launch(d_fw[r]);
    for (int r = rounds - 1; r >= 0; r--) launch(d_bw[r]);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), d_state, N * sizeof(float2), cudaMemcpyDeviceToHost));
    double dev = 0, mx = 0, n2 = 0, d2 = 0;
    for (size_t i = 0; i < N; i++) {
      const double d = std::hypot(out[i].x - h[i].x, out[i].y - h[i].y);
      dev = std::max(dev, d);
      d2 += d * d;
      mx = std::max(mx, (double)std::hypot(h[i].x, h[i].y));
      n2 += (double)out[i].x * out[i].x + (double)out[i].y * out[i].y;
    }
    printf("after %d distinct blocks and their inverses (%d block applications): max |deviation| / max |amplitude| = %.3e, "
           "||deviation||_2 / ||state||_2 = %.3e, norm^2 - 1 = %.3e\n", rounds, 2 * rounds, dev / mx, std::sqrt(d2), n2 - 1.0);
    for (int r = 0; r < rounds; r++) { cudaFree(d_fw[r]); cudaFree(d_bw[r]); }
  }

  // ---- 3. throughput
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int r = 0; r < 2; r++) { launch(d_img); launch(d_img_dag); }
  CK(cudaEventRecord(e0));
  const int reps = 10;
  for (int r = 0; r < reps; r++) { launch(d_img); launch(d_img_dag); }
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  ms /= 2 * reps;
  const double bytes = 2.0 * N * sizeof(float2);
  printf("one block pass: %.3f ms = %.1f GB/s of HBM traffic (read + write); 9 gates per block -> %.2f us per gate, "
         "%.1f TFLOP/s bf16 tensor (%d MMAs of 128x64x16 per tile)\n",
         ms, bytes / ms * 1e-6, ms * 1e3 / 9, 8.0 * products * 2 * 128 * 64 * 16 * (double)p.ntiles / ms * 1e-9, 8 * products);
  int herr = 0;
  CK(cudaMemcpy(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost));
  printf("watchdog flag: %d\n", herr);
  return 0;
}
