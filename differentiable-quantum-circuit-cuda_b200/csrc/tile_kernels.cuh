// Tiled multi-gate passes: several gates per HBM round trip.
//
// A pass fixes a set of T physical bit positions (the low L bits, so that every
// global access is a >= 128-byte contiguous run, plus T-L arbitrary higher
// positions).  The 2^n state decomposes into 2^(n-T) independent tiles of 2^T
// amplitudes; a persistent CTA loads a tile into shared memory with coalesced
// 128-bit accesses, applies EVERY gate of the pass in shared memory (128-bit
// LDS/STS, gate matrices read from the kernel-parameter constant bank, so FFMAs
// take constant operands), and writes the tile back once.
//
//   forward : 2*S bytes for the whole group instead of 2*S per gate.
//   backward: the tile holds state AND adjoint; per gate (reverse order)
//             fwd <- U^-1 fwd, grad += bwd (x) fwd, bwd <- U^T bwd; gradient
//             partials are reduced warp -> CTA (deterministic order) and kept
//             in shared memory across the CTA's tiles; 4*S bytes per group
//             instead of 4*S (reference: 6*S, src/circuit.rs:320-333) per gate.
//
// Once a pass holds more than ~4 gates the kernel leaves the HBM-bound regime
// and becomes FP32/FP64-pipe bound (16 FMA per amplitude per 2-qubit gate).
#pragma once
#include <mutex>

#include "circuit.cuh"

#define QDC_TILE_NT_F 128  // threads per CTA, forward tile kernel (6 CTAs / SM: more independent barrier domains)
#define QDC_TILE_NT_B 128  // backward tile kernel: fewer threads, more registers each (3 CTAs / SM)

#ifdef QDC_F64
#define QDC_TILE_MAXG_F 48
#define QDC_TILE_MAXG_B 24
#else
#define QDC_TILE_MAXG_F 64
#define QDC_TILE_MAXG_B 32
#endif

// out = OR_k ((x >> src_k) & (2^width_k - 1)) << dst_k : a software bit-deposit
struct BitDeposit {
  int nseg;
  unsigned char src[8], dst[8], width[8];
  __host__ __device__ __forceinline__ uint64_t operator()(uint64_t x) const {
    uint64_t out = 0;
    for (int k = 0; k < nseg; k++) out |= ((x >> src[k]) & ((1ull << width[k]) - 1ull)) << dst[k];
    return out;
  }
};

static inline bool make_deposit(const std::vector<int>& positions, BitDeposit* d) {
  d->nseg = 0;
  size_t i = 0;
  while (i < positions.size()) {
    size_t j = i + 1;
    while (j < positions.size() && positions[j] == positions[j - 1] + 1) j++;
    if (d->nseg == 8) return false;
    d->src[d->nseg] = (unsigned char)i;
    d->dst[d->nseg] = (unsigned char)positions[i];
    d->width[d->nseg] = (unsigned char)(j - i);
    d->nseg++;
    i = j;
  }
  return true;
}

struct TileGeo {
  int T, L;
  BitDeposit hi;    // run index within the tile (T-L bits) -> amplitude offset
  BitDeposit tile;  // tile number (n-T bits)               -> amplitude base
  uint64_t ntiles;
};

enum { TG_Q1 = 0, TG_Q2 = 1, TG_DIAG = 2 };

// Gate matrix as the tile kernels consume it: plain re / im scalars in the
// kernel-parameter constant bank.  The gate index is CTA-uniform, so ptxas keeps
// the entries in UNIFORM registers (LDCU) and feeds them to the packed FP32 FMA
// of sm_100 (FFMA2, `fma.rn.f32x2`) as broadcast scalars; the swap and the lane
// negation of the complex product are FFMA2 operand modifiers:
//   o += g * a   ==   FFMA2 o, a.HI_LO, g.re.F32, o ;  FFMA2 o, -a.LO_HI.NP, g.im.F32, o
// i.e. two instructions per complex multiply-accumulate and NO operand
// preparation (the previous pre-packed (re,re)/(-im,im) pairs cost one LDC.64 per
// two FFMA2 and a pair of MOVs per adjoint element; SASS-checked).
struct GateMat {
  real_t re[16], im[16];
};
__host__ __device__ __forceinline__ real_t gm_re(const GateMat& m, int j) { return m.re[j]; }
__host__ __device__ __forceinline__ real_t gm_im(const GateMat& m, int j) { return m.im[j]; }
static inline void gm_fill(GateMat& m, const real_t* re, const real_t* im) {
  for (int i = 0; i < 16; i++) {
    m.re[i] = re[i];
    m.im[i] = im[i];
  }
}
#ifndef QDC_F64
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
  unsigned long long ra = *reinterpret_cast<const unsigned long long*>(&a),
                     rb = *reinterpret_cast<const unsigned long long*>(&b),
                     rc = *reinterpret_cast<const unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(const float2 a, const float2 b) {
  unsigned long long ra = *reinterpret_cast<const unsigned long long*>(&a),
                     rb = *reinterpret_cast<const unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
// o += (gr + i gi) * a
__device__ __forceinline__ float2 cmac2(const float gr, const float gi, const float2 a, const float2 o) {
  return ffma2(make_float2(gi, gi), make_float2(-a.y, a.x), ffma2(make_float2(gr, gr), a, o));
}
#endif

__device__ __forceinline__ real_t gm_sel4_re(const GateMat& m, int j) {
  return (j & 2) ? ((j & 1) ? gm_re(m, 3) : gm_re(m, 2)) : ((j & 1) ? gm_re(m, 1) : gm_re(m, 0));
}
__device__ __forceinline__ real_t gm_sel4_im(const GateMat& m, int j) {
  return (j & 2) ? ((j & 1) ? gm_im(m, 3) : gm_im(m, 2)) : ((j & 1) ? gm_im(m, 1) : gm_im(m, 0));
}

struct TileGateF {  // one matrix
  int type, a, b, pad;  // tile-local bits; q2: a > b and the matrix is in (a,b) order; diag: j = 2 bit(a) + bit(b)
  GateMat m;
};
struct TileGateB {  // inverse (for the state) and transpose (for the adjoint)
  int type, a, b, slot;  // slot < 0: constant gate (no gradient)
  GateMat inv, tr;
};
struct TileFwdParams {
  TileGeo geo;
  int ngates;
  TileGateF g[QDC_TILE_MAXG_F];
};
struct TileBwdParams {
  TileGeo geo;
  int ngates;
  TileGateB g[QDC_TILE_MAXG_B];
};

// ------------------------------------------------------------ tile <-> HBM
// Per-thread addressing of a tile, computed ONCE per kernel: vector v = tid +
// i * NT of the tile lives at  tile_base + lo + it[i]  (in vec_t units), where
// `lo` depends on the thread only and `it[i]` on the iteration only (the run
// index splits into disjoint thread / iteration bits).
template <int NT>
struct TileAddr {
  static constexpr int MAXI = 16;  // nvec / NT for the largest supported tile
  uint64_t lo;
  uint32_t it[MAXI];  // in units of runs (2^(L - LV) vectors): fits 32 bits for n <= 34
  int niter, runv_log;
  __device__ __forceinline__ void init(const TileGeo& geo) {
    const int nvec = 1 << (geo.T - QDC_LV);
    runv_log = geo.L - QDC_LV;
    const int tid = threadIdx.x;
    niter = nvec / NT;
    lo = (geo.hi((uint64_t)(tid >> runv_log)) >> QDC_LV) + (uint64_t)(tid & ((1 << runv_log) - 1));
#pragma unroll
    for (int i = 0; i < MAXI; i++)
      it[i] = (i < niter) ? (uint32_t)(geo.hi((uint64_t)i * (NT >> runv_log)) >> geo.L) : 0u;
  }
};

template <int NT, bool LOAD>
__device__ __forceinline__ void tile_io(vec_t* __restrict__ gmem, vec_t* __restrict__ smv, const TileAddr<NT>& ta,
                                        uint64_t tile_base_vec) {
  const uint64_t base = tile_base_vec + ta.lo;
  const int tid = threadIdx.x;
  constexpr int UNR = 4;
#pragma unroll
  for (int i0 = 0; i0 < TileAddr<NT>::MAXI; i0 += UNR) {
    if (i0 < ta.niter) {
      vec_t tmp[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++)
        if (LOAD) tmp[u] = gmem[base + ((uint64_t)ta.it[i0 + u] << ta.runv_log)];
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        const int v = tid + (i0 + u) * NT;
        if (LOAD) smv[v] = tmp[u]; else gmem[base + ((uint64_t)ta.it[i0 + u] << ta.runv_log)] = smv[v];
      }
    }
  }
}

// ------------------------------------------------- in-tile gate application
template <int K>
__device__ __forceinline__ void mv(const GateMat& G, cplx_t (&a)[K]) {
#ifdef QDC_F64
  cplx_t o[K];
#pragma unroll
  for (int r = 0; r < K; r++) {
    o[r].x = 0;
    o[r].y = 0;
#pragma unroll
    for (int c = 0; c < K; c++) cmac(o[r], G.re[r * K + c], G.im[r * K + c], a[c]);
  }
#pragma unroll
  for (int r = 0; r < K; r++) a[r] = o[r];
#else
  float2 o[K];
#pragma unroll
  for (int r = 0; r < K; r++) {
    o[r] = fmul2(make_float2(G.re[r * K], G.re[r * K]), a[0]);
    o[r] = ffma2(make_float2(G.im[r * K], G.im[r * K]), make_float2(-a[0].y, a[0].x), o[r]);
#pragma unroll
    for (int c = 1; c < K; c++) o[r] = cmac2(G.re[r * K + c], G.im[r * K + c], a[c], o[r]);
  }
#pragma unroll
  for (int r = 0; r < K; r++) a[r] = o[r];
#endif
}

// acc[2(pK+q)] += b[p] * a[q] (no conjugation), packed on f32
template <int K>
__device__ __forceinline__ void outer_tile(const cplx_t (&b)[K], const cplx_t (&a)[K], real_t* acc) {
#ifdef QDC_F64
  outer_acc<K>(b, a, acc);
#else
#pragma unroll
  for (int p = 0; p < K; p++) {
#pragma unroll
    for (int q = 0; q < K; q++) {
      float2 t = make_float2(acc[2 * (p * K + q)], acc[2 * (p * K + q) + 1]);
      t = cmac2(b[p].x, b[p].y, a[q], t);
      acc[2 * (p * K + q)] = t.x;
      acc[2 * (p * K + q) + 1] = t.y;
    }
  }
#endif
}

// forward: a <- G a over every item of geometry Geo in the shared tile
template <int NT, class Geo>
__device__ __forceinline__ void tile_apply(vec_t* smv, const Geo& geo, int nitems, const TileGateF& G) {
  constexpr int K = Geo::K;
  for (int i0 = 0; i0 < nitems; i0 += NT) {  // uniform trip count (nitems % NT == 0): convergent loop
    const int i = i0 + threadIdx.x;
    VecU v[Geo::NVEC];
    const uint32_t base = geo.base32((uint32_t)i);
#pragma unroll
    for (int c = 0; c < Geo::NVEC; c++) v[c].v = smv[base + geo.off32(c)];
    cplx_t a[Geo::NG][K];
    Geo::unpack(v, a);
#pragma unroll
    for (int e = 0; e < Geo::NG; e++) mv<K>(G.m, a[e]);
    Geo::pack(v, a);
#pragma unroll
    for (int c = 0; c < Geo::NVEC; c++) smv[base + geo.off32(c)] = v[c].v;
  }
}

template <int NT>
__device__ __forceinline__ void tile_diag(vec_t* smv, int nvec, const GateMat& D, int a, int b) {
  for (int i0 = 0; i0 < nvec; i0 += NT) {  // uniform trip count
    const int i = i0 + threadIdx.x;
    VecU v;
    v.v = smv[i];
#pragma unroll
    for (int e = 0; e < QDC_VA; e++) {
      const int amp = i * QDC_VA + e;
      const int j = 2 * ((amp >> a) & 1) + ((amp >> b) & 1);
      const real_t dr = gm_sel4_re(D, j), di = gm_sel4_im(D, j);
      const real_t x = v.r[2 * e] * dr - v.r[2 * e + 1] * di, y = v.r[2 * e] * di + v.r[2 * e + 1] * dr;
      v.r[2 * e] = x;
      v.r[2 * e + 1] = y;
    }
    smv[i] = v.v;
  }
}

// Diagonal gate on tile bits a > b through the quad geometry: the four vectors of an item are the four
// settings of (bit a, bit b), so vector c takes entry d[c] -- compile-time indices, no per-amplitude
// selection (the select-based tile_diag / tile_rev_diag above compile to divergent branches over the
// constant bank: 300+ instructions per vector).  One amplitude per vector (f64) or b >= 1.
template <int NT>
__device__ __forceinline__ void tile_diag_quad(vec_t* smv, const GeoQ2HH& geo, int nitems, const GateMat& D) {
  for (int i0 = 0; i0 < nitems; i0 += NT) {  // uniform trip count
    const uint32_t base = geo.base32((uint32_t)(i0 + threadIdx.x));
#pragma unroll
    for (int c = 0; c < 4; c++) {
      VecU v;
      v.v = smv[base + geo.off32(c)];
#pragma unroll
      for (int e = 0; e < QDC_VA; e++) {
        const real_t x = v.r[2 * e] * D.re[c] - v.r[2 * e + 1] * D.im[c];
        const real_t y = v.r[2 * e] * D.im[c] + v.r[2 * e + 1] * D.re[c];
        v.r[2 * e] = x;
        v.r[2 * e + 1] = y;
      }
      smv[base + geo.off32(c)] = v.v;
    }
  }
}

template <int NT>
__device__ __forceinline__ void tile_rev_diag_quad(vec_t* smf, vec_t* smb, const GeoQ2HH& geo, int nitems,
                                                   const TileGateB& G, real_t* acc) {
  for (int i0 = 0; i0 < nitems; i0 += NT) {  // uniform trip count
    const uint32_t base = geo.base32((uint32_t)(i0 + threadIdx.x));
#pragma unroll
    for (int c = 0; c < 4; c++) {
      VecU vf, vb;
      vf.v = smf[base + geo.off32(c)];
      vb.v = smb[base + geo.off32(c)];
#pragma unroll
      for (int e = 0; e < QDC_VA; e++) {
        const real_t fx = vf.r[2 * e] * G.inv.re[c] - vf.r[2 * e + 1] * G.inv.im[c];
        const real_t fy = vf.r[2 * e] * G.inv.im[c] + vf.r[2 * e + 1] * G.inv.re[c];
        vf.r[2 * e] = fx;
        vf.r[2 * e + 1] = fy;
        const real_t bx = vb.r[2 * e], by = vb.r[2 * e + 1];
        if (G.slot >= 0) {
          acc[2 * c] += bx * fx - by * fy;
          acc[2 * c + 1] += bx * fy + by * fx;
        }
        vb.r[2 * e] = bx * G.tr.re[c] - by * G.tr.im[c];
        vb.r[2 * e + 1] = bx * G.tr.im[c] + by * G.tr.re[c];
      }
      smf[base + geo.off32(c)] = vf.v;
      smb[base + geo.off32(c)] = vb.v;
    }
  }
}

__global__ void __launch_bounds__(QDC_TILE_NT_F, 6)
    k_tile_fwd(cplx_t* __restrict__ state, const __grid_constant__ TileFwdParams p) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  vec_t* smv = (vec_t*)tile_smem;
  const int nvec = 1 << (p.geo.T - QDC_LV);
  TileAddr<QDC_TILE_NT_F> ta;
  ta.init(p.geo);
  for (uint64_t tile = blockIdx.x; tile < p.geo.ntiles; tile += gridDim.x) {
    const uint64_t tbase = p.geo.tile(tile) >> QDC_LV;
    tile_io<QDC_TILE_NT_F, true>((vec_t*)state, smv, ta, tbase);
    __syncthreads();
    for (int g = 0; g < p.ngates; g++) {
      const TileGateF& G = p.g[g];
      if (G.type == TG_Q2) {
#ifndef QDC_F64
        if (G.b == 0) {
          GeoQ2LH geo;
          geo.hv = G.a - 1;
          tile_apply<QDC_TILE_NT_F>(smv, geo, nvec / 2, G);
        } else
#endif
        {
          GeoQ2HH geo;
          geo.lv = G.b - QDC_LV;
          geo.hv = G.a - QDC_LV;
          tile_apply<QDC_TILE_NT_F>(smv, geo, nvec / 4, G);
        }
      } else if (G.type == TG_Q1) {
#ifndef QDC_F64
        if (G.a == 0) {
          GeoQ1L geo;
          tile_apply<QDC_TILE_NT_F>(smv, geo, nvec, G);
        } else
#endif
        {
          GeoQ1H geo;
          geo.pv = G.a - QDC_LV;
          tile_apply<QDC_TILE_NT_F>(smv, geo, nvec / 2, G);
        }
      } else if (G.b >= QDC_LV) {
        GeoQ2HH geo;
        geo.lv = G.b - QDC_LV;
        geo.hv = G.a - QDC_LV;
        tile_diag_quad<QDC_TILE_NT_F>(smv, geo, nvec / 4, G.m);
      } else {
        tile_diag<QDC_TILE_NT_F>(smv, nvec, G.m, G.a, G.b);
      }
      __syncthreads();
    }
    tile_io<QDC_TILE_NT_F, false>((vec_t*)state, smv, ta, tbase);
    __syncthreads();
  }
}

// ---------------------------------------------------------------- backward
template <int NT, class Geo>
__device__ __forceinline__ void tile_rev(vec_t* smf, vec_t* smb, const Geo& geo, int nitems, const TileGateB& G,
                                         real_t* acc) {
  constexpr int K = Geo::K;
  // ONE sweep per gate: un-compute the state, gradient from (pre-gate state, post-gate adjoint), pull
  // the adjoint back -- 32 B of shared-memory traffic per amplitude instead of the 40 B of separate
  // un-compute / gradient sweeps (shared bandwidth is 128 B/clk/SM against 128 (64 f64) FMA/clk/SM).
  // The gate matrices are CTA-uniform kernel parameters (uniform registers), so holding both costs
  // no vector registers.
  for (int i0 = 0; i0 < nitems; i0 += NT) {  // uniform trip count (nitems % NT == 0): convergent loop
    const int i = i0 + threadIdx.x;
    VecU vf[Geo::NVEC], vb[Geo::NVEC];
    const uint32_t base = geo.base32((uint32_t)i);
#pragma unroll
    for (int c = 0; c < Geo::NVEC; c++) {
      vf[c].v = smf[base + geo.off32(c)];
      vb[c].v = smb[base + geo.off32(c)];
    }
    cplx_t a[Geo::NG][K], b[Geo::NG][K];
    Geo::unpack(vf, a);
    Geo::unpack(vb, b);
#pragma unroll
    for (int e = 0; e < Geo::NG; e++) mv<K>(G.inv, a[e]);
    Geo::pack(vf, a);
#pragma unroll
    for (int c = 0; c < Geo::NVEC; c++) smf[base + geo.off32(c)] = vf[c].v;
    if (G.slot >= 0) {
#pragma unroll
      for (int e = 0; e < Geo::NG; e++) outer_tile<K>(b[e], a[e], acc);
    }
#pragma unroll
    for (int e = 0; e < Geo::NG; e++) mv<K>(G.tr, b[e]);
    Geo::pack(vb, b);
#pragma unroll
    for (int c = 0; c < Geo::NVEC; c++) smb[base + geo.off32(c)] = vb[c].v;
  }
}

template <int NT>
__device__ __forceinline__ void tile_rev_diag(vec_t* smf, vec_t* smb, int nvec, const TileGateB& G, real_t* acc) {
  for (int i0 = 0; i0 < nvec; i0 += NT) {  // uniform trip count
    const int i = i0 + threadIdx.x;
    VecU vf, vb;
    vf.v = smf[i];
    vb.v = smb[i];
#pragma unroll
    for (int e = 0; e < QDC_VA; e++) {
      const int amp = i * QDC_VA + e;
      const int j = 2 * ((amp >> G.a) & 1) + ((amp >> G.b) & 1);
      real_t dr = gm_sel4_re(G.inv, j), di = gm_sel4_im(G.inv, j);
      const real_t fx = vf.r[2 * e] * dr - vf.r[2 * e + 1] * di, fy = vf.r[2 * e] * di + vf.r[2 * e + 1] * dr;
      vf.r[2 * e] = fx;
      vf.r[2 * e + 1] = fy;
      const real_t bx = vb.r[2 * e], by = vb.r[2 * e + 1];
      if (G.slot >= 0) {
        const real_t pr = bx * fx - by * fy, pi = bx * fy + by * fx;
#pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          acc[2 * jj] += (j == jj) ? pr : (real_t)0;
          acc[2 * jj + 1] += (j == jj) ? pi : (real_t)0;
        }
      }
      dr = gm_sel4_re(G.tr, j);
      di = gm_sel4_im(G.tr, j);
      vb.r[2 * e] = bx * dr - by * di;
      vb.r[2 * e + 1] = bx * di + by * dr;
    }
    smf[i] = vf.v;
    smb[i] = vb.v;
  }
}

// partials: [gridDim.x][ngates][32] doubles
__global__ void __launch_bounds__(QDC_TILE_NT_B, 3)
    k_tile_bwd(cplx_t* __restrict__ fwd, cplx_t* __restrict__ bwd, const __grid_constant__ TileBwdParams p,
               double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  const int nvec = 1 << (p.geo.T - QDC_LV);
  vec_t* smf = (vec_t*)tile_smem;
  vec_t* smb = smf + nvec;
  double* sm_acc = (double*)(smb + nvec);                    // [MAXG_B][32]
  real_t* sm_part = (real_t*)(sm_acc + QDC_TILE_MAXG_B * 32);  // [2][warps][32]
  constexpr int NW = QDC_TILE_NT_B / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < QDC_TILE_MAXG_B * 32; i += QDC_TILE_NT_B) sm_acc[i] = 0.0;
  __syncthreads();
  TileAddr<QDC_TILE_NT_B> ta;
  ta.init(p.geo);
  for (uint64_t tile = blockIdx.x; tile < p.geo.ntiles; tile += gridDim.x) {
    const uint64_t tbase = p.geo.tile(tile) >> QDC_LV;
    tile_io<QDC_TILE_NT_B, true>((vec_t*)fwd, smf, ta, tbase);
    tile_io<QDC_TILE_NT_B, true>((vec_t*)bwd, smb, ta, tbase);
    __syncthreads();
    for (int g = 0; g < p.ngates; g++) {
      const TileGateB& G = p.g[g];
      real_t acc[32];
#pragma unroll
      for (int k = 0; k < 32; k++) acc[k] = 0;
      if (G.type == TG_Q2) {
#ifndef QDC_F64
        if (G.b == 0) {
          GeoQ2LH geo;
          geo.hv = G.a - 1;
          tile_rev<QDC_TILE_NT_B>(smf, smb, geo, nvec / 2, G, acc);
        } else
#endif
        {
          GeoQ2HH geo;
          geo.lv = G.b - QDC_LV;
          geo.hv = G.a - QDC_LV;
          tile_rev<QDC_TILE_NT_B>(smf, smb, geo, nvec / 4, G, acc);
        }
      } else if (G.type == TG_Q1) {
#ifndef QDC_F64
        if (G.a == 0) {
          GeoQ1L geo;
          tile_rev<QDC_TILE_NT_B>(smf, smb, geo, nvec, G, acc);
        } else
#endif
        {
          GeoQ1H geo;
          geo.pv = G.a - QDC_LV;
          tile_rev<QDC_TILE_NT_B>(smf, smb, geo, nvec / 2, G, acc);
        }
      } else if (G.b >= QDC_LV) {
        GeoQ2HH geo;
        geo.lv = G.b - QDC_LV;
        geo.hv = G.a - QDC_LV;
        tile_rev_diag_quad<QDC_TILE_NT_B>(smf, smb, geo, nvec / 4, G, acc);
      } else {
        tile_rev_diag<QDC_TILE_NT_B>(smf, smb, nvec, G, acc);
      }
      real_t* part = sm_part + (size_t)(g & 1) * NW * 32;
      if (G.slot >= 0) {
        double d = 0.0;
        warp_flush<32>(acc, d, lane);  // lane j now holds the warp total of value j
        part[warp * 32 + lane] = (real_t)d;
      }
      __syncthreads();
      if (G.slot >= 0 && warp == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; w++) s += (double)part[w * 32 + lane];
        sm_acc[g * 32 + lane] += s;
      }
    }
    tile_io<QDC_TILE_NT_B, false>((vec_t*)fwd, smf, ta, tbase);
    tile_io<QDC_TILE_NT_B, false>((vec_t*)bwd, smb, ta, tbase);
    __syncthreads();
  }
  for (int i = threadIdx.x; i < p.ngates * 32; i += QDC_TILE_NT_B)
    partials[(size_t)blockIdx.x * p.ngates * 32 + i] = sm_acc[i];
}

// out[slot(g)][j] += sum_cta partials[cta][g][j]    (one block per gate, 32 threads)
struct TileSlots {
  int s[QDC_TILE_MAXG_B];
};
__global__ void k_tile_final(const double* __restrict__ partials, int ncta, int ngates, const TileSlots slots,
                             double* __restrict__ out) {
  const int g = blockIdx.x, j = threadIdx.x;
  const int slot = slots.s[g];
  if (slot < 0) return;
  double s = 0;
  for (int c = 0; c < ncta; c++) s += partials[((size_t)c * ngates + g) * 32 + j];
  out[(size_t)slot * 32 + j] += s;
}

// ------------------------------------------------------------ host launch
#ifndef QDC_F64
// pair-lane layout kernels (tile_soa_kernels.cuh), the f32 default
__global__ void __launch_bounds__(QDC_TILE_NT_F, 6)
    k_tile_fwd_soa(cplx_t* __restrict__ state, const __grid_constant__ TileFwdParams p);
__global__ void __launch_bounds__(QDC_TILE_NT_B, 3)
    k_tile_bwd_soa(cplx_t* __restrict__ fwd, cplx_t* __restrict__ bwd, const __grid_constant__ TileBwdParams p,
                   double* __restrict__ partials);
#define QDC_KFWD (opt_soa_ ? k_tile_fwd_soa : k_tile_fwd)
#define QDC_KBWD (opt_soa_ ? k_tile_bwd_soa : k_tile_bwd)
#else
#define QDC_KFWD k_tile_fwd
#define QDC_KBWD k_tile_bwd
#endif
static inline const char* make_tile_geo_bits(std::vector<int> bits, int n_loc, int low_bits, TileGeo* geo,
                                             std::vector<int>* tile_pos_of);
static inline const char* make_tile_geo(const qdc::Plan& plan, const qdc::Step& t, int n_loc, int low_bits,
                                        TileGeo* geo, std::vector<int>* tile_pos_of) {
  return make_tile_geo_bits(
      std::vector<int>(plan.tile_bits.begin() + t.tb_first, plan.tile_bits.begin() + t.tb_first + t.tb_count),
      n_loc, low_bits, geo, tile_pos_of);
}

// Geometry of a tile over the physical positions `bits` (padded with the lowest unused positions).
static inline const char* make_tile_geo_bits(std::vector<int> bits, int n_loc, int low_bits, TileGeo* geo,
                                             std::vector<int>* tile_pos_of) {
  // pad with the lowest unused positions so that T >= log2(threads * vector) and runs stay whole
  int T = (int)bits.size();
  // every thread gets at least one 4-vector quad item (QDC_LV + 10), and the register-blocked kernels need one
  // 2^4-amplitude block per thread of a 128-thread CTA (11): a 2^10 f64 tile made their upper 64 threads redo
  // the blocks of the lower 64 (found with window-grown tiles that close before the bit budget is used up)
  const int min_T = (QDC_LV + 8 + 2) > 11 ? (QDC_LV + 8 + 2) : 11;
  for (int p = 0; T < min_T && p < n_loc; p++) {
    if (std::find(bits.begin(), bits.end(), p) == bits.end()) {
      bits.push_back(p);
      T++;
    }
  }
  std::sort(bits.begin(), bits.end());
  if (T < min_T) return qdc_errf("register too small for a tiled pass.");
  if (T - QDC_LV > 11) return qdc_errf("tile_bits too large: at most %d.", 11 + QDC_LV);
  int L = 0;
  while (L < T && bits[L] == L) L++;
  if (L < low_bits && L < T) return qdc_errf("tile does not contain the low bits.");
  if (L > T - 1) L = T - 1;
  // run = the low L contiguous bits; cap so that a run does not exceed one thread-row
  const int maxL = QDC_LV + 7;  // vectors per run <= threads of the narrower (backward) kernel
  if (L > maxL) L = maxL;
  geo->T = T;
  geo->L = L;
  std::vector<int> hi(bits.begin() + L, bits.end());
  std::vector<int> rest;
  for (int p = 0; p < n_loc; p++)
    if (std::find(bits.begin(), bits.end(), p) == bits.end()) rest.push_back(p);
  if (!make_deposit(hi, &geo->hi) || !make_deposit(rest, &geo->tile))
    return qdc_errf("tile bit set too fragmented.");
  geo->ntiles = 1ull << (n_loc - T);
  tile_pos_of->assign(n_loc, -1);
  for (int k = 0; k < T; k++) (*tile_pos_of)[bits[k]] = k;
  return nullptr;
}

inline void Circuit::release_tiles() {
  if (tile_partials_) cudaFree(tile_partials_);
  tile_partials_ = nullptr;
  tile_partials_cap_ = 0;
}

static inline const char* tile_grid(const void* kernel, int threads, size_t smem, uint64_t ntiles, int* grid) {
  DeviceInfo di;
  QDC_TRY(qdc_device_info(&di));
  // resident CTAs per SM of (kernel, threads, smem) on this device: queried once (the attribute call and the
  // occupancy query cost more than the launch itself on small registers)
  // The attribute is per kernel and device for the whole PROCESS: the bookkeeping is process-wide too.
  struct Entry { const void* kernel; int threads, device; size_t smem; int bps; };
  static std::vector<Entry> cache;
  static std::mutex cache_mutex;
  std::lock_guard<std::mutex> lock(cache_mutex);
  int bps = -1;
  for (const Entry& e : cache)
    if (e.kernel == kernel && e.threads == threads && e.smem == smem && e.device == di.device) bps = e.bps;
  if (bps < 0) {
    size_t allowed = 0;  // the attribute is a per-kernel MAXIMUM: only ever raise it
    for (const Entry& e : cache)
      if (e.kernel == kernel && e.device == di.device && e.smem > allowed) allowed = e.smem;
    if (smem > allowed)
      QDC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QDC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, threads, smem));
    cache.push_back(Entry{kernel, threads, di.device, smem, bps});
  }
  if (bps < 1) return qdc_errf("tile kernel does not fit on an SM (%zu bytes of shared memory).", smem);
  const uint64_t cap = (uint64_t)di.sm_count * bps;
  *grid = (int)(ntiles < cap ? ntiles : cap);
  return nullptr;
}

// Fill the (hi,lo)-ordered matrix of a dense gate, or the 4 diagonal entries.
static inline const char* tile_matrix(const cplx_t* gate, int kind, int form, bool swap, GateMat* out) {
  real_t re[16], im[16];
  for (int i = 0; i < 16; i++) re[i] = im[i] = 0;
  if (kind_is_q1(kind)) {
    zc m[4];
    QDC_TRY(make_form<2>(gate, form, m));
    split<4>(m, re, im);
  } else if (kind_is_q2dense(kind)) {
    zc m[16];
    QDC_TRY(make_form<4>(gate, form, m));
    to_hilo(m, swap);
    split<16>(m, re, im);
  } else {
    for (int j = 0; j < 4; j++) {  // entries in (hi,lo) order: j = 2 bit_hi + bit_lo
      const int src = swap ? (((j & 1) << 1) | (j >> 1)) : j;
      re[j] = gate[src].x;
      im[j] = (form == FORM_CONJ_TR) ? -gate[src].y : gate[src].y;
    }
  }
  gm_fill(*out, re, im);
  return nullptr;
}

inline const char* Circuit::run_tile_forward(const qdc::Step& t, const std::vector<const cplx_t*>& gp) {
  static thread_local TileFwdParams p;  // large: keep off the stack
  std::vector<int> tpos;
  QDC_TRY(make_tile_geo(plan_, t, n_loc_, 0, &p.geo, &tpos));
  if (t.count > QDC_TILE_MAXG_F) return qdc_errf("tile pass holds too many gates.");
  p.ngates = t.count;
  for (int k = 0; k < t.count; k++) {
    const qdc::Step& st = plan_.tile_steps[t.first + k];
    const int kind = insts_[st.inst].kind;
    TileGateF& G = p.g[k];
    G.pad = 0;
    const bool swap = st.p1 >= 0 && st.p2 < st.p1;
    QDC_TRY(tile_matrix(gp[st.inst], kind, FORM_PLAIN, swap, &G.m));
    if (kind_is_q1(kind)) {
      G.type = TG_Q1;
      G.a = tpos[st.p2];
      G.b = -1;
    } else if (kind_is_q2dense(kind)) {
      G.type = TG_Q2;
      G.a = tpos[swap ? st.p1 : st.p2];
      G.b = tpos[swap ? st.p2 : st.p1];
    } else {
      G.type = TG_DIAG;
      G.a = tpos[swap ? st.p1 : st.p2];
      G.b = tpos[swap ? st.p2 : st.p1];
    }
  }
  const size_t smem = sizeof(cplx_t) << p.geo.T;
  int grid = 0;
  QDC_TRY(tile_grid((const void*)QDC_KFWD, QDC_TILE_NT_F, smem, p.geo.ntiles, &grid));
  cudaEvent_t pa = nullptr;
  if (prof_.on) pa = prof_.begin(stream_);
  QDC_KFWD<<<grid, QDC_TILE_NT_F, smem, stream_>>>(state_, p);
  QDC_CUDA(cudaGetLastError());
  if (prof_.on) prof_.end(stream_, CAT_TILE_FWD, pa, 2ull * t.count * bytes());
  stats_.kernel_launches += 1;
  stats_.hbm_passes += 1;
  stats_.algorithmic_bytes += 2ull * t.count * bytes();
  return nullptr;
}

inline const char* Circuit::run_tile_backward(const qdc::Step& t, const std::vector<const cplx_t*>& gp,
                                              const std::vector<long>& vslot, bool live) {
  if (!live) {
    // no adjoint yet: un-compute the group with the forward kernel (inverse matrices, reverse order)
    static thread_local TileFwdParams p;
    std::vector<int> tpos;
    QDC_TRY(make_tile_geo(plan_, t, n_loc_, 0, &p.geo, &tpos));
    p.ngates = t.count;
    for (int k = 0; k < t.count; k++) {
      const qdc::Step& st = plan_.tile_steps[t.first + t.count - 1 - k];
      const int kind = insts_[st.inst].kind;
      TileGateF& G = p.g[k];
      G.pad = 0;
      const bool swap = st.p1 >= 0 && st.p2 < st.p1;
      QDC_TRY(tile_matrix(gp[st.inst], kind, kind_is_nonu(kind) ? FORM_INV : FORM_CONJ_TR, swap, &G.m));
      G.type = kind_is_q1(kind) ? TG_Q1 : (kind_is_q2dense(kind) ? TG_Q2 : TG_DIAG);
      if (G.type != TG_Q1) {
        G.a = tpos[swap ? st.p1 : st.p2];
        G.b = tpos[swap ? st.p2 : st.p1];
      } else {
        G.a = tpos[st.p2];
        G.b = -1;
      }
    }
    const size_t smem = sizeof(cplx_t) << p.geo.T;
    int grid = 0;
    QDC_TRY(tile_grid((const void*)QDC_KFWD, QDC_TILE_NT_F, smem, p.geo.ntiles, &grid));
    cudaEvent_t pa = nullptr;
    if (prof_.on) pa = prof_.begin(stream_);
    QDC_KFWD<<<grid, QDC_TILE_NT_F, smem, stream_>>>(state_, p);
    QDC_CUDA(cudaGetLastError());
    if (prof_.on) prof_.end(stream_, CAT_UNCOMPUTE, pa, 2ull * t.count * bytes());
    stats_.kernel_launches += 1;
    stats_.hbm_passes += 1;
    stats_.algorithmic_bytes += 2ull * t.count * bytes();
    return nullptr;
  }
  static thread_local TileBwdParams p;
  std::vector<int> tpos;
  QDC_TRY(make_tile_geo(plan_, t, n_loc_, 0, &p.geo, &tpos));
  if (t.count > QDC_TILE_MAXG_B) return qdc_errf("tile pass holds too many gates for the backward kernel.");
  p.ngates = t.count;
  TileSlots h_slots;
  for (int k = 0; k < t.count; k++) {
    const qdc::Step& st = plan_.tile_steps[t.first + t.count - 1 - k];
    const int kind = insts_[st.inst].kind;
    TileGateB& G = p.g[k];
    const bool swap = st.p1 >= 0 && st.p2 < st.p1;
    QDC_TRY(tile_matrix(gp[st.inst], kind, kind_is_nonu(kind) ? FORM_INV : FORM_CONJ_TR, swap, &G.inv));
    QDC_TRY(tile_matrix(gp[st.inst], kind, kind_is_diag(kind) ? FORM_PLAIN : FORM_TR, swap, &G.tr));
    G.type = kind_is_q1(kind) ? TG_Q1 : (kind_is_q2dense(kind) ? TG_Q2 : TG_DIAG);
    if (G.type != TG_Q1) {
      G.a = tpos[swap ? st.p1 : st.p2];
      G.b = tpos[swap ? st.p2 : st.p1];
    } else {
      G.a = tpos[st.p2];
      G.b = -1;
    }
    if (G.type == TG_DIAG) diag_hilo_[st.inst] = swap;  // the gradient comes back in (hi,lo) order
    G.slot = (int)vslot[st.inst];
    h_slots.s[k] = G.slot;
  }
  const size_t tile_bytes = sizeof(cplx_t) << p.geo.T;
  const size_t smem = 2 * tile_bytes + QDC_TILE_MAXG_B * 32 * sizeof(double) +
                      2 * (QDC_TILE_NT_B / 32) * 32 * sizeof(real_t);
  int grid = 0;
  QDC_TRY(tile_grid((const void*)QDC_KBWD, QDC_TILE_NT_B, smem, p.geo.ntiles, &grid));
  const size_t need = (size_t)grid * QDC_TILE_MAXG_B * 32;
  if (need > tile_partials_cap_) {
    if (tile_partials_) QDC_CUDA(cudaFree(tile_partials_));
    QDC_CUDA(cudaMalloc((void**)&tile_partials_, need * sizeof(double)));
    tile_partials_cap_ = need;
  }
  cudaEvent_t pa = nullptr;
  if (prof_.on) pa = prof_.begin(stream_);
  QDC_KBWD<<<grid, QDC_TILE_NT_B, smem, stream_>>>(state_, bwd_, p, tile_partials_);
  QDC_CUDA(cudaGetLastError());
  k_tile_final<<<t.count, 32, 0, stream_>>>(tile_partials_, grid, t.count, h_slots, d_res_);
  QDC_CUDA(cudaGetLastError());
  if (prof_.on) prof_.end(stream_, CAT_TILE_BWD, pa, 4ull * t.count * bytes());
  stats_.kernel_launches += 2;
  stats_.hbm_passes += 2;
  stats_.algorithmic_bytes += 4ull * t.count * bytes();
  return nullptr;
}
