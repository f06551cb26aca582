// Host mathematics of the tensor-core fused blocks (no CUDA): 64 x 64 complex matrices in double precision --
// the block matrix of a window and the chain rule from its block gradient to the member gates (tc_exec.cuh).
// Kept free of device code so that the CPU test suite can compile it with g++ (tests/test_tc_host.py).
#pragma once
#include <complex>
#include <vector>

typedef std::complex<double> zc;
typedef std::vector<zc> Mat64;  // 64 x 64, row-major

struct TcGate {
  int inst;        // instruction index
  int nq;          // 1 or 2 qubits
  bool diag;
  int b2, b1;      // kernel index bits of (pos2, pos1) (b1 = -1 for one-qubit gates)
  zc m[16];        // dense matrix in (pos2, pos1) index order (2 x 2 in m[0..3] for one-qubit gates)
};

struct TcPass {
  std::vector<TcGate> gates;   // forward execution order
  int grad_slot = -1;          // slot in d_tcgrad_ (reverse pass with a live adjoint)
  bool from_h = false;         // the slot holds P' of the fused reverse kernel (tc_chain_rule_from_h), not P
};

// M <- (g embedded on index bits) * M   (acts on the row index)
static inline void tc_apply_rows(Mat64& m, const zc* g, int nq, int b2, int b1) {
  if (nq == 1) {
    const int bit = 1 << b2;
    for (int i = 0; i < 64; i++) {
      if (i & bit) continue;
      for (int j = 0; j < 64; j++) {
        const zc x0 = m[i * 64 + j], x1 = m[(i | bit) * 64 + j];
        m[i * 64 + j] = g[0] * x0 + g[1] * x1;
        m[(i | bit) * 64 + j] = g[2] * x0 + g[3] * x1;
      }
    }
    return;
  }
  const int h = 1 << b2, l = 1 << b1;
  for (int i = 0; i < 64; i++) {
    if (i & (h | l)) continue;
    const int r[4] = {i, i | l, i | h, i | h | l};
    for (int j = 0; j < 64; j++) {
      const zc x[4] = {m[r[0] * 64 + j], m[r[1] * 64 + j], m[r[2] * 64 + j], m[r[3] * 64 + j]};
      for (int o = 0; o < 4; o++) m[r[o] * 64 + j] = g[4 * o] * x[0] + g[4 * o + 1] * x[1] + g[4 * o + 2] * x[2] + g[4 * o + 3] * x[3];
    }
  }
}

// M <- M * (g embedded)^T :  (M g^T)[p, q'] = sum_q M[p, q] g[q', q]   (acts on the column index)
static inline void tc_apply_cols(Mat64& m, const zc* g, int nq, int b2, int b1) {
  if (nq == 1) {
    const int bit = 1 << b2;
    for (int p = 0; p < 64; p++)
      for (int j = 0; j < 64; j++) {
        if (j & bit) continue;
        const zc x0 = m[p * 64 + j], x1 = m[p * 64 + (j | bit)];
        m[p * 64 + j] = g[0] * x0 + g[1] * x1;
        m[p * 64 + (j | bit)] = g[2] * x0 + g[3] * x1;
      }
    return;
  }
  const int h = 1 << b2, l = 1 << b1;
  for (int p = 0; p < 64; p++)
    for (int j = 0; j < 64; j++) {
      if (j & (h | l)) continue;
      const int c[4] = {j, j | l, j | h, j | h | l};
      const zc x[4] = {m[p * 64 + c[0]], m[p * 64 + c[1]], m[p * 64 + c[2]], m[p * 64 + c[3]]};
      for (int o = 0; o < 4; o++) m[p * 64 + c[o]] = g[4 * o] * x[0] + g[4 * o + 1] * x[1] + g[4 * o + 2] * x[2] + g[4 * o + 3] * x[3];
    }
}

// partial trace of E over the other index bits: out[x * K + y] = sum_r E[(x, r), (y, r)]
static inline void tc_partial_trace(const Mat64& e, int nq, int b2, int b1, zc* out) {
  const int K = nq == 1 ? 2 : 4;
  for (int k = 0; k < K * K; k++) out[k] = zc(0, 0);
  const int mask = nq == 1 ? (1 << b2) : ((1 << b2) | (1 << b1));
  auto place = [&](int x) { return nq == 1 ? (x << b2) : ((((x >> 1) & 1) << b2) | ((x & 1) << b1)); };
  for (int r = 0; r < 64; r++) {
    if (r & mask) continue;
    for (int x = 0; x < K; x++)
      for (int y = 0; y < K; y++) out[x * K + y] += e[(r | place(x)) * 64 + (r | place(y))];
  }
}

// Block gradient -> gate gradients.  e: G_W[i, j] = sum adjoint_after[i] * state_before[j] (kernel bit order), destroyed.
// `want(k)` says whether gate k needs a gradient; `emit(k, values, count)` receives it in the reference's layout
// (row-major 2 x 2 / 4 x 4, or the 4 diagonal entries).
template <class Want, class Emit>
static inline void tc_chain_rule_g(const TcPass& pass, Mat64& e, Want want, Emit emit) {
  const int m = (int)pass.gates.size();
  zc tmp[16];
  // E_1 = U_2^T ( ... (U_m^T G_W))
  for (int k = m - 1; k >= 1; k--) {
    const TcGate& g = pass.gates[k];
    const int K = g.nq == 1 ? 2 : 4;
    for (int r = 0; r < K; r++)
      for (int c = 0; c < K; c++) tmp[r * K + c] = g.m[c * K + r];
    tc_apply_rows(e, tmp, g.nq, g.b2, g.b1);
  }
  for (int k = 0; k < m; k++) {
    const TcGate& g = pass.gates[k];
    if (want(k)) {
      zc full[16];
      tc_partial_trace(e, g.nq, g.b2, g.b1, full);
      if (g.diag) {
        const zc d[4] = {full[0], full[5], full[10], full[15]};
        emit(k, d, 4);
      } else {
        emit(k, full, g.nq == 1 ? 4 : 16);
      }
    }
    if (k + 1 < m) {
      const TcGate& gn = pass.gates[k + 1];
      const int K = gn.nq == 1 ? 2 : 4;
      for (int i = 0; i < K * K; i++) tmp[i] = std::conj(gn.m[i]);
      tc_apply_rows(e, tmp, gn.nq, gn.b2, gn.b1);    // (U_{k+1}^T)^-1 = conj(U_{k+1}) for unitary gates
      tc_apply_cols(e, g.m, g.nq, g.b2, g.b1);       // ... E U_k^T
    }
  }
}

// P: the 128 x 128 real matrix of k_tc_block_grad (row = real-ified adjoint index, column = real-ified index of the
// state BEFORE the block, kernel bit order).
template <class Want, class Emit>
static inline void tc_chain_rule(const TcPass& pass, const double* p, Want want, Emit emit) {
  Mat64 e(64 * 64);
  for (int i = 0; i < 64; i++)
    for (int j = 0; j < 64; j++)
      e[i * 64 + j] = zc(p[i * 128 + j] - p[(64 + i) * 128 + 64 + j], p[i * 128 + 64 + j] + p[(64 + i) * 128 + j]);
  tc_chain_rule_g(pass, e, want, emit);
}

// P': the 128 x 128 real matrix of the fused reverse kernel k_tc_block_rev (tc_rev.cuh): row = real-ified index of the
// CONJUGATED adjoint after the block, column = real-ified index of the state AFTER the block (both as loaded), i.e.
// H[i, k] = sum adjoint[i] * state_after[k] has Re = P'[i, k] + P'[64 + i, 64 + k], Im = P'[i, 64 + k] - P'[64 + i, k].
// With state_before = W^dagger state_after:  G_W = H conj(W) = H conj(U_m) conj(U_{m-1}) ... conj(U_1).
template <class Want, class Emit>
static inline void tc_chain_rule_from_h(const TcPass& pass, const double* p, Want want, Emit emit) {
  Mat64 e(64 * 64);
  for (int i = 0; i < 64; i++)
    for (int k = 0; k < 64; k++)
      e[i * 64 + k] = zc(p[i * 128 + k] + p[(64 + i) * 128 + 64 + k], p[i * 128 + 64 + k] - p[(64 + i) * 128 + k]);
  zc tmp[16];
  for (int k = (int)pass.gates.size() - 1; k >= 0; k--) {
    const TcGate& g = pass.gates[k];
    const int K = g.nq == 1 ? 2 : 4;
    for (int r = 0; r < K; r++)   // M conj(U) = M (U^dagger)^T
      for (int c = 0; c < K; c++) tmp[r * K + c] = std::conj(g.m[c * K + r]);
    tc_apply_cols(e, tmp, g.nq, g.b2, g.b1);
  }
  tc_chain_rule_g(pass, e, want, emit);
}

static inline void tc_block_matrix(const TcPass& pass, Mat64& w) {
  w.assign(64 * 64, zc(0, 0));
  for (int i = 0; i < 64; i++) w[i * 64 + i] = zc(1, 0);
  for (const TcGate& g : pass.gates) tc_apply_rows(w, g.m, g.nq, g.b2, g.b1);
}

