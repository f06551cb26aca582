// Host-side launchers for the streaming kernels: geometry selection, gate
// re-ordering to (hi, lo) physical-bit order, host-side inverses for NonU
// gates, and the reduction workspace.  Shared by the 18 legacy C symbols
// (primitives_abi.cuh) and the circuit executor (circuit.cuh).
#pragma once
#include <complex>
#include <vector>

#include "stream_kernels.cuh"

typedef std::complex<double> zc;

// ---------------------------------------------------------------- workspace
struct Workspace {
  double* partials = nullptr;  // [cap_blocks][32]
  int cap_blocks = 0;
  double* red_out = nullptr;   // scratch for one reduction result (32 doubles)
  int device = -1;
};

static inline const char* ws_ensure(Workspace& ws, int blocks) {
  int dev = 0;
  QDC_CUDA(cudaGetDevice(&dev));
  if (ws.device != dev) {  // first use on this device (buffers of another device are leaked)
    ws = Workspace();
    ws.device = dev;
  }
  if (ws.red_out == nullptr) QDC_CUDA(cudaMalloc(&ws.red_out, 32 * sizeof(double)));
  if (blocks > ws.cap_blocks) {
    if (ws.partials) QDC_CUDA(cudaFree(ws.partials));
    int cap = blocks < 2048 ? 2048 : blocks;
    QDC_CUDA(cudaMalloc(&ws.partials, (size_t)cap * 32 * sizeof(double)));
    ws.cap_blocks = cap;
  }
  return nullptr;
}

static inline void ws_release(Workspace& ws) {
  if (ws.partials) cudaFree(ws.partials);
  if (ws.red_out) cudaFree(ws.red_out);
  ws = Workspace();
}

// ------------------------------------------------------------ host gate math
static inline int perm2(int j) { return ((j & 1) << 1) | ((j >> 1) & 1); }  // swap the two index bits

// flat row-major KxK complex (build precision) -> double
template <int K>
static inline void load_gate(const cplx_t* g, zc (&m)[K * K]) {
  for (int i = 0; i < K * K; i++) m[i] = zc((double)g[i].x, (double)g[i].y);
}

template <int K>
static inline void transpose(zc (&m)[K * K]) {
  for (int r = 0; r < K; r++)
    for (int c = r + 1; c < K; c++) std::swap(m[r * K + c], m[c * K + r]);
}

template <int K>
static inline void conjugate(zc (&m)[K * K]) {
  for (int i = 0; i < K * K; i++) m[i] = std::conj(m[i]);
}

// Gauss-Jordan inverse with partial pivoting, double precision.  Replaces the
// cuBLAS matinvBatched call of the reference (src/primitives.cu:114-138) and
// keeps its error text for a singular pivot.
template <int K>
static inline const char* invert(zc (&m)[K * K]) {
  zc a[K][2 * K];
  for (int r = 0; r < K; r++)
    for (int c = 0; c < K; c++) {
      a[r][c] = m[r * K + c];
      a[r][K + c] = (r == c) ? zc(1, 0) : zc(0, 0);
    }
  for (int col = 0; col < K; col++) {
    int piv = col;
    double best = std::abs(a[col][col]);
    for (int r = col + 1; r < K; r++)
      if (std::abs(a[r][col]) > best) {
        best = std::abs(a[r][col]);
        piv = r;
      }
    if (best == 0.0) return qdc_errf("U(%d, %d) is zero.", col + 1, col + 1);
    if (piv != col)
      for (int c = 0; c < 2 * K; c++) std::swap(a[piv][c], a[col][c]);
    const zc inv = zc(1, 0) / a[col][col];
    for (int c = 0; c < 2 * K; c++) a[col][c] *= inv;
    for (int r = 0; r < K; r++) {
      if (r == col) continue;
      const zc f = a[r][col];
      if (f == zc(0, 0)) continue;
      for (int c = 0; c < 2 * K; c++) a[r][c] -= f * a[col][c];
    }
  }
  for (int r = 0; r < K; r++)
    for (int c = 0; c < K; c++) m[r * K + c] = a[r][K + c];
  return nullptr;
}

// 4x4 in (pos2,pos1) order -> (hi,lo) order when pos2 is the lower bit.
static inline void to_hilo(zc (&m)[16], bool swap) {
  if (!swap) return;
  zc t[16];
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) t[r * 4 + c] = m[perm2(r) * 4 + perm2(c)];
  for (int i = 0; i < 16; i++) m[i] = t[i];
}

template <int KK>
static inline void split(const zc (&m)[KK], real_t* re, real_t* im) {
  for (int i = 0; i < KK; i++) {
    re[i] = (real_t)m[i].real();
    im[i] = (real_t)m[i].imag();
  }
}

// ------------------------------------------------------------------ launch
static inline int pick_grid(uint64_t nitems, int U, int blocks_per_sm, int sms) {
  const uint64_t per_block = (uint64_t)QDC_BLOCK * U;
  uint64_t need = (nitems + per_block - 1) / per_block;
  uint64_t cap = (uint64_t)sms * (blocks_per_sm > 0 ? blocks_per_sm : 1);
  uint64_t g = need < cap ? need : cap;
  return (int)(g < 1 ? 1 : g);
}

// out_dev: device double[Op::NRED] receiving the reduced values (nullptr -> ws.red_out)
template <class Geo, class Op, int U>
static const char* launch_stream(cudaStream_t st, Workspace& ws, cplx_t* a0, cplx_t* a1, const Geo& geo,
                                 const Op& op, uint64_t nitems, double* out_dev, int accumulate) {
  static thread_local int bps = 0;
  DeviceInfo di;
  QDC_TRY(qdc_device_info(&di));
  if (bps == 0)
    QDC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_stream<Geo, Op, U>, QDC_BLOCK, 0));
  const int grid = pick_grid(nitems, U, bps, di.sm_count);
  if (Op::NRED > 0) QDC_TRY(ws_ensure(ws, grid));
  k_stream<Geo, Op, U><<<grid, QDC_BLOCK, 0, st>>>((vec_t*)a0, (vec_t*)a1, geo, op, nitems, ws.partials);
  QDC_CUDA(cudaGetLastError());
  if (Op::NRED > 0) {
    k_final_reduce<<<1, 32, 0, st>>>(ws.partials, grid, Op::NRED, out_dev ? out_dev : ws.red_out, accumulate);
    QDC_CUDA(cudaGetLastError());
  }
  return nullptr;
}

template <class Op, int U>
static const char* launch_elem(cudaStream_t st, Workspace& ws, cplx_t* a0, cplx_t* a1, const Op& op, int n,
                               double* out_dev, int accumulate) {
  static thread_local int bps = 0;
  DeviceInfo di;
  QDC_TRY(qdc_device_info(&di));
  if (bps == 0) QDC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_elem<Op, U>, QDC_BLOCK, 0));
  const uint64_t nvec = n >= QDC_LV ? (1ull << (n - QDC_LV)) : 1ull;
  const int grid = pick_grid(nvec, U, bps, di.sm_count);
  if (Op::NRED > 0) QDC_TRY(ws_ensure(ws, grid));
  k_elem<Op, U><<<grid, QDC_BLOCK, 0, st>>>((vec_t*)a0, (vec_t*)a1, op, nvec, ws.partials);
  QDC_CUDA(cudaGetLastError());
  if (Op::NRED > 0) {
    k_final_reduce<<<1, 32, 0, st>>>(ws.partials, grid, Op::NRED, out_dev ? out_dev : ws.red_out, accumulate);
    QDC_CUDA(cudaGetLastError());
  }
  return nullptr;
}

// Dispatch a K=2 op over the q1 geometries.
template <class Op>
static const char* run_q1(cudaStream_t st, Workspace& ws, cplx_t* a0, cplx_t* a1, const Op& op, int pos, int n,
                          double* out_dev, int acc) {
  constexpr int U = (Op::R1 || Op::W1) ? 2 : 4;
#ifndef QDC_F64
  if (pos == 0) {
    GeoQ1L geo;
    return launch_stream<GeoQ1L, Op, 2 * U>(st, ws, a0, a1, geo, op, 1ull << (n - 1), out_dev, acc);
  }
#endif
  GeoQ1H geo;
  geo.pv = pos - QDC_LV;
  return launch_stream<GeoQ1H, Op, U>(st, ws, a0, a1, geo, op, 1ull << (n - 1 - QDC_LV), out_dev, acc);
}

// Dispatch a K=4 op (already in (hi,lo) order) over the q2 geometries.
template <class Op>
static const char* run_q2(cudaStream_t st, Workspace& ws, cplx_t* a0, cplx_t* a1, const Op& op, int lo, int hi,
                          int n, double* out_dev, int acc) {
  constexpr int U = (Op::R1 || Op::W1) ? 1 : 2;
#ifndef QDC_F64
  if (lo == 0) {
    GeoQ2LH geo;
    geo.hv = hi - 1;
    return launch_stream<GeoQ2LH, Op, 2 * U>(st, ws, a0, a1, geo, op, 1ull << (n - 2), out_dev, acc);
  }
#endif
  GeoQ2HH geo;
  geo.lv = lo - QDC_LV;
  geo.hv = hi - QDC_LV;
  return launch_stream<GeoQ2HH, Op, U>(st, ws, a0, a1, geo, op, 1ull << (n - 2 - QDC_LV), out_dev, acc);
}

// -------------------------------------------------------------- engine API
// All functions are asynchronous on `st`; gate pointers are host memory in
// the build's precision, flat row-major, (pos2,pos1) index order as in the
// reference ABI.

enum GateForm { FORM_PLAIN = 0, FORM_TR = 1, FORM_CONJ_TR = 2, FORM_INV = 3 };

template <int K>
static inline const char* make_form(const cplx_t* g, int form, zc (&m)[K * K]) {
  load_gate<K>(g, m);
  switch (form) {
    case FORM_PLAIN: break;
    case FORM_TR: transpose<K>(m); break;
    case FORM_CONJ_TR: conjugate<K>(m); transpose<K>(m); break;
    case FORM_INV: return invert<K>(m);
  }
  return nullptr;
}

static const char* eng_q1gate(cudaStream_t st, Workspace& ws, cplx_t* state, const cplx_t* gate, int form,
                              int pos, int n) {
  zc m[4];
  QDC_TRY(make_form<2>(gate, form, m));
  OpApply<2> op;
  split<4>(m, op.re, op.im);
  return run_q1(st, ws, state, nullptr, op, pos, n, nullptr, 0);
}

static const char* eng_q2gate(cudaStream_t st, Workspace& ws, cplx_t* state, const cplx_t* gate, int form,
                              int pos2, int pos1, int n) {
  zc m[16];
  QDC_TRY(make_form<4>(gate, form, m));
  to_hilo(m, pos2 < pos1);
  OpApply<4> op;
  split<16>(m, op.re, op.im);
  const int lo = pos2 < pos1 ? pos2 : pos1, hi = pos2 < pos1 ? pos1 : pos2;
  return run_q2(st, ws, state, nullptr, op, lo, hi, n, nullptr, 0);
}

static const char* eng_q2diag(cudaStream_t st, Workspace& ws, cplx_t* state, const cplx_t* gate, bool conj,
                              int pos2, int pos1, int n) {
  EOpDiagApply op;
  op.sel.pos2 = pos2;
  op.sel.pos1 = pos1;
  for (int j = 0; j < 4; j++) {
    op.re[j] = gate[j].x;
    op.im[j] = conj ? -gate[j].y : gate[j].y;
  }
  return launch_elem<EOpDiagApply, 8>(st, ws, state, nullptr, op, n, nullptr, 0);
}

// Fused reverse step through a q1 gate.  inv_form: FORM_CONJ_TR (unitary) or
// FORM_INV (NonU).  grad_dev: device double[8] (kernel order == reference
// order for q1) or nullptr for a constant gate.
static const char* eng_rev_q1(cudaStream_t st, Workspace& ws, cplx_t* fwd, cplx_t* bwd, const cplx_t* gate,
                              int inv_form, int pos, int n, double* grad_dev) {
  zc mi[4], mt[4];
  QDC_TRY(make_form<2>(gate, inv_form, mi));
  QDC_TRY(make_form<2>(gate, FORM_TR, mt));
  if (grad_dev) {
    OpRev<2> op;
    split<4>(mi, op.ire, op.iim);
    split<4>(mt, op.tre, op.tim);
    return run_q1(st, ws, fwd, bwd, op, pos, n, grad_dev, 0);
  }
  OpRevNoGrad<2> op;
  split<4>(mi, op.ire, op.iim);
  split<4>(mt, op.tre, op.tim);
  return run_q1(st, ws, fwd, bwd, op, pos, n, nullptr, 0);
}

// q2: grad_dev receives 32 doubles in (hi,lo) kernel order; use
// unpermute_q2() after the copy back.
static const char* eng_rev_q2(cudaStream_t st, Workspace& ws, cplx_t* fwd, cplx_t* bwd, const cplx_t* gate,
                              int inv_form, int pos2, int pos1, int n, double* grad_dev) {
  zc mi[16], mt[16];
  QDC_TRY(make_form<4>(gate, inv_form, mi));
  QDC_TRY(make_form<4>(gate, FORM_TR, mt));
  to_hilo(mi, pos2 < pos1);
  to_hilo(mt, pos2 < pos1);
  const int lo = pos2 < pos1 ? pos2 : pos1, hi = pos2 < pos1 ? pos1 : pos2;
  if (grad_dev) {
    OpRev<4> op;
    split<16>(mi, op.ire, op.iim);
    split<16>(mt, op.tre, op.tim);
    return run_q2(st, ws, fwd, bwd, op, lo, hi, n, grad_dev, 0);
  }
  OpRevNoGrad<4> op;
  split<16>(mi, op.ire, op.iim);
  split<16>(mt, op.tre, op.tim);
  return run_q2(st, ws, fwd, bwd, op, lo, hi, n, nullptr, 0);
}

static const char* eng_rev_diag(cudaStream_t st, Workspace& ws, cplx_t* fwd, cplx_t* bwd, const cplx_t* gate,
                                int pos2, int pos1, int n, double* grad_dev) {
  if (grad_dev) {
    EOpDiagRev<true> op;
    op.sel.pos2 = pos2;
    op.sel.pos1 = pos1;
    for (int j = 0; j < 4; j++) {
      op.re[j] = gate[j].x;
      op.im[j] = gate[j].y;
      op.ire[j] = gate[j].x;
      op.iim[j] = -gate[j].y;
    }
    return launch_elem<EOpDiagRev<true>, 4>(st, ws, fwd, bwd, op, n, grad_dev, 0);
  }
  EOpDiagRev<false> op;
  op.sel.pos2 = pos2;
  op.sel.pos1 = pos1;
  for (int j = 0; j < 4; j++) {
    op.re[j] = gate[j].x;
    op.im[j] = gate[j].y;
    op.ire[j] = gate[j].x;
    op.iim[j] = -gate[j].y;
  }
  return launch_elem<EOpDiagRev<false>, 4>(st, ws, fwd, bwd, op, n, nullptr, 0);
}

static const char* eng_grad_q1(cudaStream_t st, Workspace& ws, const cplx_t* fwd, const cplx_t* bwd, int pos,
                               int n, double* out_dev) {
  OpGrad<2> op;
  return run_q1(st, ws, (cplx_t*)fwd, (cplx_t*)bwd, op, pos, n, out_dev, 0);
}

static const char* eng_grad_q2(cudaStream_t st, Workspace& ws, const cplx_t* fwd, const cplx_t* bwd, int pos2,
                               int pos1, int n, double* out_dev) {
  OpGrad<4> op;
  const int lo = pos2 < pos1 ? pos2 : pos1, hi = pos2 < pos1 ? pos1 : pos2;
  return run_q2(st, ws, (cplx_t*)fwd, (cplx_t*)bwd, op, lo, hi, n, out_dev, 0);
}

static const char* eng_grad_diag(cudaStream_t st, Workspace& ws, const cplx_t* fwd, const cplx_t* bwd,
                                 int pos2, int pos1, int n, double* out_dev) {
  EOpDiagGrad op;
  op.sel.pos2 = pos2;
  op.sel.pos1 = pos1;
  return launch_elem<EOpDiagGrad, 4>(st, ws, (cplx_t*)fwd, (cplx_t*)bwd, op, n, out_dev, 0);
}

static const char* eng_dens_q1(cudaStream_t st, Workspace& ws, const cplx_t* state, int pos, int n,
                               double* out_dev) {
  OpDens<2> op;
  return run_q1(st, ws, (cplx_t*)state, nullptr, op, pos, n, out_dev, 0);
}

static const char* eng_dens_q2(cudaStream_t st, Workspace& ws, const cplx_t* state, int pos2, int pos1, int n,
                               double* out_dev) {
  OpDens<4> op;
  const int lo = pos2 < pos1 ? pos2 : pos1, hi = pos2 < pos1 ? pos1 : pos2;
  return run_q2(st, ws, (cplx_t*)state, nullptr, op, lo, hi, n, out_dev, 0);
}

// bwd (+)= apply_gate(G^T, 2 conj(fwd)) where G is the (already conjugated)
// density cotangent, flat row-major (src/circuit.rs:393-420).
static const char* eng_seed_q1(cudaStream_t st, Workspace& ws, const cplx_t* fwd, cplx_t* bwd, const cplx_t* g,
                               int pos, int n, bool accumulate) {
  zc m[4];
  QDC_TRY(make_form<2>(g, FORM_TR, m));
  if (accumulate) {
    OpSeed<2, true> op;
    split<4>(m, op.tre, op.tim);
    return run_q1(st, ws, (cplx_t*)fwd, bwd, op, pos, n, nullptr, 0);
  }
  OpSeed<2, false> op;
  split<4>(m, op.tre, op.tim);
  return run_q1(st, ws, (cplx_t*)fwd, bwd, op, pos, n, nullptr, 0);
}

static const char* eng_seed_q2(cudaStream_t st, Workspace& ws, const cplx_t* fwd, cplx_t* bwd, const cplx_t* g,
                               int pos2, int pos1, int n, bool accumulate) {
  zc m[16];
  QDC_TRY(make_form<4>(g, FORM_TR, m));
  to_hilo(m, pos2 < pos1);
  const int lo = pos2 < pos1 ? pos2 : pos1, hi = pos2 < pos1 ? pos1 : pos2;
  if (accumulate) {
    OpSeed<4, true> op;
    split<16>(m, op.tre, op.tim);
    return run_q2(st, ws, (cplx_t*)fwd, bwd, op, lo, hi, n, nullptr, 0);
  }
  OpSeed<4, false> op;
  split<16>(m, op.tre, op.tim);
  return run_q2(st, ws, (cplx_t*)fwd, bwd, op, lo, hi, n, nullptr, 0);
}

static const char* eng_conj_and_double(cudaStream_t st, Workspace& ws, const cplx_t* src, cplx_t* dst, int n) {
  EOpConjDouble op;
  return launch_elem<EOpConjDouble, 8>(st, ws, (cplx_t*)src, dst, op, n, nullptr, 0);
}

static const char* eng_add(cudaStream_t st, Workspace& ws, const cplx_t* src, cplx_t* dst, int n) {
  EOpAdd op;
  return launch_elem<EOpAdd, 4>(st, ws, (cplx_t*)src, dst, op, n, nullptr, 0);
}

static const char* eng_copy(cudaStream_t st, const cplx_t* src, cplx_t* dst, int n) {
  QDC_CUDA(cudaMemcpyAsync(dst, src, sizeof(cplx_t) << n, cudaMemcpyDeviceToDevice, st));
  return nullptr;
}

static const char* eng_set_standard(cudaStream_t st, cplx_t* state, int n) {
  QDC_CUDA(cudaMemsetAsync(state, 0, sizeof(cplx_t) << n, st));
  k_set_one<<<1, 1, 0, st>>>(state);
  QDC_CUDA(cudaGetLastError());
  return nullptr;
}

// Reduced values come back as 2*K*K doubles (re,im interleaved) in kernel
// (hi,lo) order; convert to the reference's (pos2,pos1) flat order.
static inline void unpermute_q2(const double* in, bool swap, zc* out16) {
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      const int rr = swap ? perm2(r) : r, cc = swap ? perm2(c) : c;
      out16[r * 4 + c] = zc(in[2 * (rr * 4 + cc)], in[2 * (rr * 4 + cc) + 1]);
    }
}
