// CPU check of csrc/tc_host.hpp: the chain rule from a block gradient G_W = sum b (x) a to the gradients of the
// member gates equals the reverse pass of src/circuit.rs:320-392 carried out gate by gate on the same vectors.
//   g++ -O2 -std=c++17 -I <csrc> tc_host_check.cpp -o tc_host_check && ./tc_host_check
#include <cstdio>
#include <random>

#include "tc_host.hpp"

static std::mt19937_64 rng(7);
static std::normal_distribution<double> nd;

static void haar(int K, zc* u) {
  std::vector<zc> a(K * K);
  for (auto& x : a) x = zc(nd(rng), nd(rng));
  for (int c = 0; c < K; c++) {
    for (int k = 0; k < c; k++) {
      zc dot = 0;
      for (int r = 0; r < K; r++) dot += std::conj(a[r * K + k]) * a[r * K + c];
      for (int r = 0; r < K; r++) a[r * K + c] -= dot * a[r * K + k];
    }
    double nrm = 0;
    for (int r = 0; r < K; r++) nrm += std::norm(a[r * K + c]);
    for (int r = 0; r < K; r++) a[r * K + c] /= std::sqrt(nrm);
  }
  for (int i = 0; i < K * K; i++) u[i] = a[i];
}

int main() {
  double worst = 0;
  for (int trial = 0; trial < 20; trial++) {
    TcPass pass;
    const int m = 3 + (int)(rng() % 10);
    for (int k = 0; k < m; k++) {
      TcGate g;
      g.inst = k;
      const int kind = (int)(rng() % 3);   // 0: dense q2, 1: q1, 2: diagonal q2
      g.nq = kind == 1 ? 1 : 2;
      g.diag = kind == 2;
      g.b2 = (int)(rng() % 6);
      g.b1 = -1;
      if (g.nq == 2) do { g.b1 = (int)(rng() % 6); } while (g.b1 == g.b2);
      for (auto& x : g.m) x = zc(0, 0);
      if (kind == 0) haar(4, g.m);
      else if (kind == 1) haar(2, g.m);
      else for (int i = 0; i < 4; i++) g.m[5 * i] = std::polar(1.0, nd(rng));
      pass.gates.push_back(g);
    }
    // 64 samples as the columns of A (states before the block) and B (adjoints after the block)
    Mat64 A(64 * 64), B(64 * 64);
    for (auto& x : A) x = zc(nd(rng), nd(rng));
    for (auto& x : B) x = zc(nd(rng), nd(rng));
    // P[mu, nu] = sum_r B~[mu, r] A~[nu, r]
    std::vector<double> P(128 * 128, 0.0);
    for (int mu = 0; mu < 128; mu++)
      for (int nu = 0; nu < 128; nu++) {
        double s = 0;
        for (int r = 0; r < 64; r++) {
          const zc b = B[(mu & 63) * 64 + r], a = A[(nu & 63) * 64 + r];
          s += (mu < 64 ? b.real() : b.imag()) * (nu < 64 ? a.real() : a.imag());
        }
        P[mu * 128 + nu] = s;
      }
    std::vector<std::vector<zc>> got(m);
    tc_chain_rule(pass, P.data(), [](int) { return true; },
                  [&](int k, const zc* v, int count) { got[k].assign(v, v + count); });
    // the fused reverse kernel's form: P'[mu, nu] = sum_r conj(B)~[mu, r] X~[nu, r] with X = W A the state AFTER the block
    {
      Mat64 X(A);
      for (const TcGate& g : pass.gates) tc_apply_rows(X, g.m, g.nq, g.b2, g.b1);
      std::vector<double> Pp(128 * 128, 0.0);
      for (int mu = 0; mu < 128; mu++)
        for (int nu = 0; nu < 128; nu++) {
          double s = 0;
          for (int r = 0; r < 64; r++) {
            const zc b = std::conj(B[(mu & 63) * 64 + r]), x = X[(nu & 63) * 64 + r];
            s += (mu < 64 ? b.real() : b.imag()) * (nu < 64 ? x.real() : x.imag());
          }
          Pp[mu * 128 + nu] = s;
        }
      std::vector<std::vector<zc>> got_h(m);
      tc_chain_rule_from_h(pass, Pp.data(), [](int) { return true; },
                           [&](int k, const zc* v, int count) { got_h[k].assign(v, v + count); });
      for (int k = 0; k < m; k++) {
        if (got_h[k].size() != got[k].size()) { printf("FAIL: from_h gradient length\n"); return 1; }
        for (size_t i = 0; i < got[k].size(); i++) worst = std::max(worst, std::abs(got_h[k][i] - got[k][i]) / 64.0);
      }
    }
    // the block matrix is the ordered product of the gates
    {
      Mat64 w, seq(A);
      tc_block_matrix(pass, w);
      for (const TcGate& g : pass.gates) tc_apply_rows(seq, g.m, g.nq, g.b2, g.b1);
      for (int i = 0; i < 64; i++)
        for (int r = 0; r < 64; r++) {
          zc s = 0;
          for (int j = 0; j < 64; j++) s += w[i * 64 + j] * A[j * 64 + r];
          worst = std::max(worst, std::abs(s - seq[i * 64 + r]));
        }
    }
    // direct: S = U_{k-1} .. U_1 A,  T = U_{k+1}^T .. U_m^T B,  G_k = partial trace of T S^T
    for (int k = 0; k < m; k++) {
      Mat64 S(A), T(B);
      for (int i = 0; i < k; i++) tc_apply_rows(S, pass.gates[i].m, pass.gates[i].nq, pass.gates[i].b2, pass.gates[i].b1);
      for (int i = m - 1; i > k; i--) {
        const TcGate& g = pass.gates[i];
        const int K = g.nq == 1 ? 2 : 4;
        zc t[16];
        for (int r = 0; r < K; r++)
          for (int c = 0; c < K; c++) t[r * K + c] = g.m[c * K + r];
        tc_apply_rows(T, t, g.nq, g.b2, g.b1);
      }
      Mat64 E(64 * 64);
      for (int p = 0; p < 64; p++)
        for (int q = 0; q < 64; q++) {
          zc s = 0;
          for (int r = 0; r < 64; r++) s += T[p * 64 + r] * S[q * 64 + r];
          E[p * 64 + q] = s;
        }
      const TcGate& g = pass.gates[k];
      zc full[16];
      tc_partial_trace(E, g.nq, g.b2, g.b1, full);
      const int K = g.nq == 1 ? 2 : 4;
      for (size_t i = 0; i < got[k].size(); i++) {
        const zc want = g.diag ? full[5 * i] : full[i];
        worst = std::max(worst, std::abs(got[k][i] - want) / 64.0);
      }
      if ((int)got[k].size() != (g.diag ? 4 : K * K)) { printf("FAIL: wrong gradient length\n"); return 1; }
    }
  }
  printf("max deviation %.3e\n", worst);
  if (!(worst < 1e-11)) { printf("FAIL\n"); return 1; }
  printf("OK\n");
  return 0;
}
