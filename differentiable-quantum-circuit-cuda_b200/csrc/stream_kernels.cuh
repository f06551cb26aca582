// Streaming (one HBM pass per call) kernels for sm_100a.
//
// Every kernel here is HBM-bandwidth bound, so the design rules are the
// memory ones: 128-bit (float4 / double2) coalesced accesses, >= 8 independent
// 128-bit loads in flight per thread, grids sized to SM count x resident CTAs,
// 64-bit index arithmetic (n up to 34 local qubits), gate matrices in kernel
// parameters.  They replace the reference's `<<<128,128>>>` scalar kernels
// (/root/reference/src/primitives.cu:202-253, 295-354, 398-452, 513-532,
// 573-606, 649-672, 689-739, 779-837, 879-939) and add the fused reverse step
// (un-compute + gradient reduction + adjoint pull-back in one pass, 4*S bytes
// instead of the 6*S of src/circuit.rs:320-333, 348-363) and the fused
// density seed (2*S / 3*S instead of 7*S, src/circuit.rs:393-420).
#pragma once
#include "qdc_common.cuh"

#define QDC_BLOCK 256
#define QDC_FLUSH_EVERY 4

// ------------------------------------------------------------ complex math
__device__ __forceinline__ void cmac(cplx_t& o, real_t gr, real_t gi, const cplx_t a) {
  o.x = fma(gr, a.x, o.x);
  o.x = fma(-gi, a.y, o.x);
  o.y = fma(gr, a.y, o.y);
  o.y = fma(gi, a.x, o.y);
}

template <int K>
__device__ __forceinline__ void matvec(const real_t* __restrict__ gre, const real_t* __restrict__ gim,
                                       cplx_t (&a)[K]) {
  cplx_t o[K];
#pragma unroll
  for (int r = 0; r < K; r++) {
    o[r].x = 0;
    o[r].y = 0;
#pragma unroll
    for (int c = 0; c < K; c++) cmac(o[r], gre[r * K + c], gim[r * K + c], a[c]);
  }
#pragma unroll
  for (int r = 0; r < K; r++) a[r] = o[r];
}

// acc[2(pK+q)] += b[p]*a[q]   (no conjugation; src/primitives.cu:222-228, 323-329)
template <int K>
__device__ __forceinline__ void outer_acc(const cplx_t (&b)[K], const cplx_t (&a)[K], real_t* acc) {
#pragma unroll
  for (int p = 0; p < K; p++)
#pragma unroll
    for (int q = 0; q < K; q++) {
      real_t& re = acc[2 * (p * K + q)];
      real_t& im = acc[2 * (p * K + q) + 1];
      re = fma(b[p].x, a[q].x, re);
      re = fma(-b[p].y, a[q].y, re);
      im = fma(b[p].x, a[q].y, im);
      im = fma(b[p].y, a[q].x, im);
    }
}

// acc[2(pK+q)] += a[p]*conj(a[q])   (src/primitives.cu:707-715, 805-813)
template <int K>
__device__ __forceinline__ void dens_acc(const cplx_t (&a)[K], real_t* acc) {
#pragma unroll
  for (int p = 0; p < K; p++)
#pragma unroll
    for (int q = 0; q < K; q++) {
      real_t& re = acc[2 * (p * K + q)];
      real_t& im = acc[2 * (p * K + q) + 1];
      re = fma(a[p].x, a[q].x, re);
      re = fma(a[p].y, a[q].y, re);
      im = fma(a[p].y, a[q].x, im);
      im = fma(-a[p].x, a[q].y, im);
    }
}

// ------------------------------------------------------------- reductions
// Warp "reduce-scatter": NV per-thread partial sums -> lane L ends up holding
// the warp total of value index L >> (5 - log2 NV), using NV-1 (+ a few)
// shuffles instead of 5*NV.  The warp total is then accumulated in double so
// that long grid-stride loops do not pile up f32 rounding error.
template <int C, int M>
struct WarpHalve {
  static __device__ __forceinline__ void run(real_t* acc, int lane) {
    if constexpr (M >= 1) {
      if constexpr (C > 1) {
        constexpr int h = C / 2;
        const bool up = (lane & M) != 0;
#pragma unroll
        for (int k = 0; k < h; k++) {
          const real_t send = up ? acc[k] : acc[k + h];
          const real_t keep = up ? acc[k + h] : acc[k];
          acc[k] = keep + __shfl_xor_sync(0xffffffffu, send, M);
        }
        WarpHalve<h, M / 2>::run(acc, lane);
      } else {
        acc[0] += __shfl_xor_sync(0xffffffffu, acc[0], M);
        WarpHalve<1, M / 2>::run(acc, lane);
      }
    }
  }
};

template <int NV>
__device__ __forceinline__ void warp_flush(real_t (&acc)[NV], double& dacc, int lane) {
  WarpHalve<NV, 16>::run(acc, lane);
  dacc += (double)acc[0];
#pragma unroll
  for (int k = 0; k < NV; k++) acc[k] = 0;
}

// Block stage: one double per (warp, value) -> partials[block][NV].
template <int NV>
__device__ __forceinline__ void block_reduce_store(double dacc, double* __restrict__ partials) {
  __shared__ double sm[QDC_BLOCK / 32][NV];
  constexpr int REP = 32 / NV;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((lane & (REP - 1)) == 0) sm[warp][lane / REP] = dacc;
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < QDC_BLOCK / 32; w++) s += sm[w][threadIdx.x];
    partials[(size_t)blockIdx.x * NV + threadIdx.x] = s;
  }
}

// Deterministic second stage: out[j] (+)= sum_b partials[b][j].
__global__ void k_final_reduce(const double* __restrict__ partials, int nblocks, int nv,
                               double* __restrict__ out, int accumulate) {
  const int j = threadIdx.x;
  if (j >= nv) return;
  double s = 0;
  for (int b = 0; b < nblocks; b++) s += partials[(size_t)b * nv + j];
  out[j] = accumulate ? out[j] + s : s;
}

// ---------------------------------------------------------- access geometry
// An "item" is what one thread processes per unrolled slot: NVEC 128-bit
// vectors holding NG groups of K amplitudes (K = 2: a pair, K = 4: a quad).
// Quads are always presented in (hi, lo) physical-bit order: index 2*b_hi+b_lo;
// the host permutes gate matrices / results when pos2 is the lower bit.

// q1 gate on bit pos >= QDC_LV: two vectors 2^pos apart, QDC_VA pairs.
struct GeoQ1H {
  static constexpr int NVEC = 2, NG = QDC_VA, K = 2;
  int pv;  // bit position in vector-index space
  __device__ __forceinline__ uint64_t base(uint64_t i) const { return ins0(i, pv); }
  __device__ __forceinline__ uint64_t off(int c) const { return (uint64_t)c << pv; }
  __device__ __forceinline__ uint32_t base32(uint32_t i) const { return ins0_32(i, pv); }
  __device__ __forceinline__ uint32_t off32(int c) const { return (uint32_t)c << pv; }
  static __device__ __forceinline__ void unpack(const VecU (&v)[NVEC], cplx_t (&a)[NG][K]) {
#pragma unroll
    for (int e = 0; e < NG; e++)
#pragma unroll
      for (int c = 0; c < K; c++) {
        a[e][c].x = v[c].r[2 * e];
        a[e][c].y = v[c].r[2 * e + 1];
      }
  }
  static __device__ __forceinline__ void pack(VecU (&v)[NVEC], const cplx_t (&a)[NG][K]) {
#pragma unroll
    for (int e = 0; e < NG; e++)
#pragma unroll
      for (int c = 0; c < K; c++) {
        v[c].r[2 * e] = a[e][c].x;
        v[c].r[2 * e + 1] = a[e][c].y;
      }
  }
};

// q2 gate on bits lo, hi >= QDC_LV: four vectors, QDC_VA quads.
struct GeoQ2HH {
  static constexpr int NVEC = 4, NG = QDC_VA, K = 4;
  int lv, hv;  // bit positions (lo < hi) in vector-index space
  __device__ __forceinline__ uint64_t base(uint64_t i) const { return ins0(ins0(i, lv), hv); }
  __device__ __forceinline__ uint64_t off(int c) const {
    return ((uint64_t)(c >> 1) << hv) + ((uint64_t)(c & 1) << lv);
  }
  __device__ __forceinline__ uint32_t base32(uint32_t i) const { return ins0_32(ins0_32(i, lv), hv); }
  __device__ __forceinline__ uint32_t off32(int c) const {
    return ((uint32_t)(c >> 1) << hv) + ((uint32_t)(c & 1) << lv);
  }
  static __device__ __forceinline__ void unpack(const VecU (&v)[NVEC], cplx_t (&a)[NG][K]) {
#pragma unroll
    for (int e = 0; e < NG; e++)
#pragma unroll
      for (int c = 0; c < K; c++) {
        a[e][c].x = v[c].r[2 * e];
        a[e][c].y = v[c].r[2 * e + 1];
      }
  }
  static __device__ __forceinline__ void pack(VecU (&v)[NVEC], const cplx_t (&a)[NG][K]) {
#pragma unroll
    for (int e = 0; e < NG; e++)
#pragma unroll
      for (int c = 0; c < K; c++) {
        v[c].r[2 * e] = a[e][c].x;
        v[c].r[2 * e + 1] = a[e][c].y;
      }
  }
};

#ifndef QDC_F64
// f32 only: the pair on bit 0 lives inside one float4.
struct GeoQ1L {
  static constexpr int NVEC = 1, NG = 1, K = 2;
  __device__ __forceinline__ uint64_t base(uint64_t i) const { return i; }
  __device__ __forceinline__ uint64_t off(int) const { return 0; }
  __device__ __forceinline__ uint32_t base32(uint32_t i) const { return i; }
  __device__ __forceinline__ uint32_t off32(int) const { return 0; }
  static __device__ __forceinline__ void unpack(const VecU (&v)[NVEC], cplx_t (&a)[NG][K]) {
    a[0][0].x = v[0].r[0];
    a[0][0].y = v[0].r[1];
    a[0][1].x = v[0].r[2];
    a[0][1].y = v[0].r[3];
  }
  static __device__ __forceinline__ void pack(VecU (&v)[NVEC], const cplx_t (&a)[NG][K]) {
    v[0].r[0] = a[0][0].x;
    v[0].r[1] = a[0][0].y;
    v[0].r[2] = a[0][1].x;
    v[0].r[3] = a[0][1].y;
  }
};

// f32 only: lo == 0 (inside the float4), hi >= 1: two vectors, one quad.
struct GeoQ2LH {
  static constexpr int NVEC = 2, NG = 1, K = 4;
  int hv;  // hi - 1
  __device__ __forceinline__ uint64_t base(uint64_t i) const { return ins0(i, hv); }
  __device__ __forceinline__ uint64_t off(int c) const { return (uint64_t)c << hv; }
  __device__ __forceinline__ uint32_t base32(uint32_t i) const { return ins0_32(i, hv); }
  __device__ __forceinline__ uint32_t off32(int c) const { return (uint32_t)c << hv; }
  static __device__ __forceinline__ void unpack(const VecU (&v)[NVEC], cplx_t (&a)[NG][K]) {
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int l = 0; l < 2; l++) {
        a[0][2 * h + l].x = v[h].r[2 * l];
        a[0][2 * h + l].y = v[h].r[2 * l + 1];
      }
  }
  static __device__ __forceinline__ void pack(VecU (&v)[NVEC], const cplx_t (&a)[NG][K]) {
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int l = 0; l < 2; l++) {
        v[h].r[2 * l] = a[0][2 * h + l].x;
        v[h].r[2 * l + 1] = a[0][2 * h + l].y;
      }
  }
};
#endif

// -------------------------------------------------------------------- ops
// Array 0 is always read.  R1: read array 1; W0/W1: write back; NRED: number
// of real reduction outputs (0 = none).
template <int K>
struct OpApply {  // a <- G a
  static constexpr bool R1 = false, W0 = true, W1 = false;
  static constexpr int NRED = 0;
  real_t re[K * K], im[K * K];
  __device__ __forceinline__ void operator()(cplx_t (&a)[K], cplx_t (&)[K], real_t*) const {
    matvec<K>(re, im, a);
  }
};

template <int K>
struct OpRev {  // fwd <- Ginv fwd ; acc += bwd (x) fwd ; bwd <- G^T bwd
  static constexpr bool R1 = true, W0 = true, W1 = true;
  static constexpr int NRED = 2 * K * K;
  real_t ire[K * K], iim[K * K], tre[K * K], tim[K * K];
  __device__ __forceinline__ void operator()(cplx_t (&a)[K], cplx_t (&b)[K], real_t* acc) const {
    matvec<K>(ire, iim, a);
    outer_acc<K>(b, a, acc);
    matvec<K>(tre, tim, b);
  }
};

template <int K>
struct OpRevNoGrad {  // const gate with a live adjoint: both updates, one launch
  static constexpr bool R1 = true, W0 = true, W1 = true;
  static constexpr int NRED = 0;
  real_t ire[K * K], iim[K * K], tre[K * K], tim[K * K];
  __device__ __forceinline__ void operator()(cplx_t (&a)[K], cplx_t (&b)[K], real_t*) const {
    matvec<K>(ire, iim, a);
    matvec<K>(tre, tim, b);
  }
};

template <int K>
struct OpGrad {  // acc += bwd (x) fwd
  static constexpr bool R1 = true, W0 = false, W1 = false;
  static constexpr int NRED = 2 * K * K;
  __device__ __forceinline__ void operator()(cplx_t (&a)[K], cplx_t (&b)[K], real_t* acc) const {
    outer_acc<K>(b, a, acc);
  }
};

template <int K>
struct OpDens {  // acc += psi (x) conj(psi)
  static constexpr bool R1 = false, W0 = false, W1 = false;
  static constexpr int NRED = 2 * K * K;
  __device__ __forceinline__ void operator()(cplx_t (&a)[K], cplx_t (&)[K], real_t* acc) const {
    dens_acc<K>(a, acc);
  }
};

template <int K, bool ACC>
struct OpSeed {  // bwd (+)= G^T (2 conj fwd)   (src/circuit.rs:393-420 in one pass)
  static constexpr bool R1 = ACC, W0 = false, W1 = true;
  static constexpr int NRED = 0;
  real_t tre[K * K], tim[K * K];
  __device__ __forceinline__ void operator()(cplx_t (&a)[K], cplx_t (&b)[K], real_t*) const {
    cplx_t t[K];
#pragma unroll
    for (int c = 0; c < K; c++) {
      t[c].x = 2 * a[c].x;
      t[c].y = -2 * a[c].y;
    }
    matvec<K>(tre, tim, t);
#pragma unroll
    for (int c = 0; c < K; c++) {
      if (ACC) {
        b[c].x += t[c].x;
        b[c].y += t[c].y;
      } else {
        b[c] = t[c];
      }
    }
  }
};

// ------------------------------------------------------- the streaming kernel
template <class Geo, class Op, int U>
__global__ void __launch_bounds__(QDC_BLOCK)
    k_stream(vec_t* __restrict__ a0, vec_t* __restrict__ a1, const Geo geo, const Op op,
             const uint64_t nitems, double* __restrict__ partials) {
  constexpr int NVEC = Geo::NVEC, NG = Geo::NG, K = Geo::K;
  constexpr int NRED = Op::NRED;
  real_t acc[NRED > 0 ? NRED : 1];
#pragma unroll
  for (int k = 0; k < (NRED > 0 ? NRED : 1); k++) acc[k] = 0;
  double dacc = 0;
  int since = 0;
  const int lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * QDC_BLOCK;
  // loop bound is warp-uniform (nitems is a power of two): shuffles stay legal
  for (uint64_t i0 = (uint64_t)blockIdx.x * QDC_BLOCK + threadIdx.x; i0 - lane < nitems;
       i0 += stride * U) {
    VecU va[U][NVEC], vb[U][NVEC];
    uint64_t base[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t i = i0 + (uint64_t)u * stride;
      base[u] = geo.base(i);
      if (i < nitems) {
#pragma unroll
        for (int c = 0; c < NVEC; c++) {
          va[u][c].v = a0[base[u] + geo.off(c)];
          if (Op::R1) vb[u][c].v = a1[base[u] + geo.off(c)];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t i = i0 + (uint64_t)u * stride;
      if (i < nitems) {
        cplx_t a[NG][K], b[NG][K];
        Geo::unpack(va[u], a);
        if (Op::R1) Geo::unpack(vb[u], b);
#pragma unroll
        for (int g = 0; g < NG; g++) op(a[g], b[g], acc);
        if (Op::W0) {
          Geo::pack(va[u], a);
#pragma unroll
          for (int c = 0; c < NVEC; c++) a0[base[u] + geo.off(c)] = va[u][c].v;
        }
        if (Op::W1) {
          Geo::pack(vb[u], b);
#pragma unroll
          for (int c = 0; c < NVEC; c++) a1[base[u] + geo.off(c)] = vb[u][c].v;
        }
      }
    }
    if constexpr (NRED > 0) {
      if (++since == QDC_FLUSH_EVERY) {
        warp_flush<(NRED > 0 ? NRED : 1)>(acc, dacc, lane);
        since = 0;
      }
    }
  }
  if constexpr (NRED > 0) {
    warp_flush<(NRED > 0 ? NRED : 1)>(acc, dacc, lane);
    block_reduce_store<(NRED > 0 ? NRED : 1)>(dacc, partials);
  }
}

// ------------------------------------------------------ elementwise kernels
// Item = one 128-bit vector (QDC_VA consecutive amplitudes).
struct DiagSel {
  int pos2, pos1;
  __device__ __forceinline__ int operator()(uint64_t amp) const {
    return 2 * (int)((amp >> pos2) & 1) + (int)((amp >> pos1) & 1);
  }
};

__device__ __forceinline__ real_t sel4(const real_t* d, int j) {
  return (j & 2) ? ((j & 1) ? d[3] : d[2]) : ((j & 1) ? d[1] : d[0]);
}

struct EOpDiagApply {  // psi *= d[2 b2 + b1]   (src/primitives.cu:649-672)
  static constexpr bool R1 = false, W0 = true, W1 = false;
  static constexpr int NRED = 0;
  DiagSel sel;
  real_t re[4], im[4];
  __device__ __forceinline__ void operator()(uint64_t amp, cplx_t& a, cplx_t&, real_t*) const {
    const int j = sel(amp);
    const real_t dr = sel4(re, j), di = sel4(im, j);
    const real_t x = a.x * dr - a.y * di, y = a.x * di + a.y * dr;
    a.x = x;
    a.y = y;
  }
};

template <bool GRAD>
struct EOpDiagRev {  // fwd *= dinv ; acc[j] += bwd fwd ; bwd *= d
  static constexpr bool R1 = true, W0 = true, W1 = true;
  static constexpr int NRED = GRAD ? 8 : 0;
  DiagSel sel;
  real_t ire[4], iim[4], re[4], im[4];
  __device__ __forceinline__ void operator()(uint64_t amp, cplx_t& a, cplx_t& b, real_t* acc) const {
    const int j = sel(amp);
    real_t dr = sel4(ire, j), di = sel4(iim, j);
    real_t x = a.x * dr - a.y * di, y = a.x * di + a.y * dr;
    a.x = x;
    a.y = y;
    if (GRAD) {
      const real_t pr = b.x * a.x - b.y * a.y, pi = b.x * a.y + b.y * a.x;
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        acc[2 * jj] += (j == jj) ? pr : (real_t)0;
        acc[2 * jj + 1] += (j == jj) ? pi : (real_t)0;
      }
    }
    dr = sel4(re, j);
    di = sel4(im, j);
    x = b.x * dr - b.y * di;
    y = b.x * di + b.y * dr;
    b.x = x;
    b.y = y;
  }
};

struct EOpDiagGrad {  // acc[j] += bwd fwd   (src/primitives.cu:398-452)
  static constexpr bool R1 = true, W0 = false, W1 = false;
  static constexpr int NRED = 8;
  DiagSel sel;
  __device__ __forceinline__ void operator()(uint64_t amp, cplx_t& a, cplx_t& b, real_t* acc) const {
    const int j = sel(amp);
    const real_t pr = b.x * a.x - b.y * a.y, pi = b.x * a.y + b.y * a.x;
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      acc[2 * jj] += (j == jj) ? pr : (real_t)0;
      acc[2 * jj + 1] += (j == jj) ? pi : (real_t)0;
    }
  }
};

struct EOpConjDouble {  // a1 = 2 conj(a0)   (src/primitives.cu:904-915)
  static constexpr bool R1 = false, W0 = false, W1 = true;
  static constexpr int NRED = 0;
  __device__ __forceinline__ void operator()(uint64_t, cplx_t& a, cplx_t& b, real_t*) const {
    b.x = 2 * a.x;
    b.y = -2 * a.y;
  }
};

struct EOpAdd {  // a1 += a0   (src/primitives.cu:931-939)
  static constexpr bool R1 = true, W0 = false, W1 = true;
  static constexpr int NRED = 0;
  __device__ __forceinline__ void operator()(uint64_t, cplx_t& a, cplx_t& b, real_t*) const {
    b.x += a.x;
    b.y += a.y;
  }
};

template <class Op, int U>
__global__ void __launch_bounds__(QDC_BLOCK)
    k_elem(vec_t* __restrict__ a0, vec_t* __restrict__ a1, const Op op, const uint64_t nvec,
           double* __restrict__ partials) {
  constexpr int NRED = Op::NRED;
  real_t acc[NRED > 0 ? NRED : 1];
#pragma unroll
  for (int k = 0; k < (NRED > 0 ? NRED : 1); k++) acc[k] = 0;
  double dacc = 0;
  int since = 0;
  const int lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * QDC_BLOCK;
  for (uint64_t i0 = (uint64_t)blockIdx.x * QDC_BLOCK + threadIdx.x; i0 - lane < nvec;
       i0 += stride * U) {
    VecU va[U], vb[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t i = i0 + (uint64_t)u * stride;
      if (i < nvec) {
        va[u].v = a0[i];
        if (Op::R1) vb[u].v = a1[i];
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint64_t i = i0 + (uint64_t)u * stride;
      if (i < nvec) {
#pragma unroll
        for (int e = 0; e < QDC_VA; e++) {
          cplx_t a, b;
          a.x = va[u].r[2 * e];
          a.y = va[u].r[2 * e + 1];
          b.x = Op::R1 ? vb[u].r[2 * e] : (real_t)0;
          b.y = Op::R1 ? vb[u].r[2 * e + 1] : (real_t)0;
          op(i * QDC_VA + e, a, b, acc);
          va[u].r[2 * e] = a.x;
          va[u].r[2 * e + 1] = a.y;
          vb[u].r[2 * e] = b.x;
          vb[u].r[2 * e + 1] = b.y;
        }
        if (Op::W0) a0[i] = va[u].v;
        if (Op::W1) a1[i] = vb[u].v;
      }
    }
    if constexpr (NRED > 0) {
      if (++since == 2 * QDC_FLUSH_EVERY) {
        warp_flush<(NRED > 0 ? NRED : 1)>(acc, dacc, lane);
        since = 0;
      }
    }
  }
  if constexpr (NRED > 0) {
    warp_flush<(NRED > 0 ? NRED : 1)>(acc, dacc, lane);
    block_reduce_store<(NRED > 0 ? NRED : 1)>(dacc, partials);
  }
}

// |0...0>: zero fill is a cudaMemsetAsync; this writes the single 1.
__global__ void k_set_one(cplx_t* state) {
  state[0].x = 1;
  state[0].y = 0;
}
