// Tiled multi-gate passes, f32, "pair-lane" shared-memory layout.
//
// Same tiling / parameters / host code as tile_kernels.cuh, but the tile is
// held in shared memory with the two amplitudes of a 128-bit vector
// de-interleaved:
//
//   HBM   float4 = (re0, im0, re1, im1)      amplitudes 2v, 2v+1 (tile bit 0)
//   smem  float4 = (re0, re1, im0, im1)      the swap is done in registers by tile_io
//
// A gate that does not act on tile bit 0 treats the two amplitudes of a vector
// identically, so (re0, re1) and (im0, im1) are the two lanes of the packed FP32
// FMA of sm_100 (FFMA2) and a complex multiply-accumulate with a CTA-uniform
// gate entry g is
//
//   o.re += g.re * a.re ; o.re += (-g.im) * a.im ; o.im += g.re * a.im ; o.im += g.im * a.re
//
// = 4 FFMA2 for two amplitudes with the gate entry as a broadcast UNIFORM-register
// operand (LDCU from the kernel-parameter bank) and negation as an operand
// modifier: no swaps, no pre-packed constants, no MOVs.  SASS of the inner loops
// (cuobjdump): forward 64 packed math + 4 LDS.128 + 4 STS.128 + ~20
// others per item (was 64 + ~60); reverse 192 + 8 + 8 + ~40 (was 192 + ~195).
// The gradient outer product accumulates per lane (acc.re / acc.im pairs) and the
// lanes are folded once per gate, before the warp reduce-scatter.
//
// The reverse pass is ONE sweep per gate (un-compute, gradient, adjoint pull-back
// on the same registers): shared-memory traffic 32 B per amplitude per gate
// instead of 40 B -- at 128 B/clk/SM shared bandwidth vs 128 FMA/clk/SM the two
// sweeps of the older kernel were as expensive as its FMAs.
//
// Gates that DO act on tile bit 0 (physical qubit 0 only) mix the lanes; they are
// unpacked to (re, im) pairs and go through the complex helpers of
// tile_kernels.cuh (rare: 1 in n-1 brickwork gates).
#pragma once
#include "tile_kernels.cuh"

#ifndef QDC_F64

struct V4 {
  float2 re, im;  // (lane 0, lane 1)
};
__device__ __forceinline__ V4 ld4(const vec_t* p) {
  const float4 t = *p;
  V4 v;
  v.re = make_float2(t.x, t.y);
  v.im = make_float2(t.z, t.w);
  return v;
}
__device__ __forceinline__ void st4(vec_t* p, const V4& v) { *p = make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }
__device__ __forceinline__ float2 bc2(const float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 neg2(const float2 a) { return make_float2(-a.x, -a.y); }

// o = G a for both lanes; G is CTA-uniform
template <int K>
__device__ __forceinline__ void mv_soa(const GateMat& G, const V4 (&a)[K], V4 (&o)[K]) {
#pragma unroll
  for (int r = 0; r < K; r++) {
    o[r].re = fmul2(bc2(G.re[r * K]), a[0].re);
    o[r].im = fmul2(bc2(G.re[r * K]), a[0].im);
    o[r].re = ffma2(bc2(-G.im[r * K]), a[0].im, o[r].re);
    o[r].im = ffma2(bc2(G.im[r * K]), a[0].re, o[r].im);
#pragma unroll
    for (int c = 1; c < K; c++) {
      o[r].re = ffma2(bc2(G.re[r * K + c]), a[c].re, o[r].re);
      o[r].im = ffma2(bc2(G.re[r * K + c]), a[c].im, o[r].im);
      o[r].re = ffma2(bc2(-G.im[r * K + c]), a[c].im, o[r].re);
      o[r].im = ffma2(bc2(G.im[r * K + c]), a[c].re, o[r].im);
    }
  }
}

// per-lane accumulation of b[p] * a[q] (no conjugation)
template <int K>
__device__ __forceinline__ void outer_soa(const V4 (&b)[K], const V4 (&a)[K], float2 (&are)[16], float2 (&aim)[16]) {
#pragma unroll
  for (int p = 0; p < K; p++) {
    const float2 nbi = neg2(b[p].im);
#pragma unroll
    for (int q = 0; q < K; q++) {
      are[p * K + q] = ffma2(b[p].re, a[q].re, are[p * K + q]);
      are[p * K + q] = ffma2(nbi, a[q].im, are[p * K + q]);
      aim[p * K + q] = ffma2(b[p].re, a[q].im, aim[p * K + q]);
      aim[p * K + q] = ffma2(b[p].im, a[q].re, aim[p * K + q]);
    }
  }
}

// ------------------------------------------------------------ tile <-> HBM
// Batches of QDC_TILE_IO_UNR independent 128-bit accesses per thread.  (Issuing all 16 loads of a
// 2^12 tile at once was measured and is NOT faster -- 576 vs 561 ms/step at 28 q -- the tile fill of
// one CTA hides behind the other CTAs' math; it only costs registers.)
#ifndef QDC_TILE_IO_UNR
#define QDC_TILE_IO_UNR 4
#endif
template <int NT, bool LOAD>
__device__ __forceinline__ void tile_io_soa(vec_t* __restrict__ gmem, vec_t* __restrict__ smv, const TileAddr<NT>& ta,
                                            uint64_t tile_base_vec) {
  const uint64_t base = tile_base_vec + ta.lo;
  const int tid = threadIdx.x;
  constexpr int UNR = QDC_TILE_IO_UNR;
#pragma unroll
  for (int i0 = 0; i0 < TileAddr<NT>::MAXI; i0 += UNR) {
    if (i0 < ta.niter) {
      vec_t tmp[UNR];
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        if (LOAD) tmp[u] = gmem[base + ((uint64_t)ta.it[i0 + u] << ta.runv_log)];
        else tmp[u] = smv[tid + (i0 + u) * NT];
      }
#pragma unroll
      for (int u = 0; u < UNR; u++) {
        const vec_t s = make_float4(tmp[u].x, tmp[u].z, tmp[u].y, tmp[u].w);  // (re0,im0,re1,im1) <-> (re0,re1,im0,im1)
        if (LOAD) smv[tid + (i0 + u) * NT] = s;
        else gmem[base + ((uint64_t)ta.it[i0 + u] << ta.runv_log)] = s;
      }
    }
  }
}

// ---------------------------------------------------------------- forward
template <int NT, class Geo>
__device__ __forceinline__ void tile_apply_soa(vec_t* smv, const Geo& geo, int nitems, const GateMat& G) {
  constexpr int K = Geo::K;
  for (int i0 = 0; i0 < nitems; i0 += NT) {  // uniform trip count (nitems % NT == 0)
    const uint32_t base = geo.base32((uint32_t)(i0 + threadIdx.x));
    V4 a[K], o[K];
#pragma unroll
    for (int c = 0; c < K; c++) a[c] = ld4(smv + base + geo.off32(c));
    mv_soa<K>(G, a, o);
#pragma unroll
    for (int c = 0; c < K; c++) st4(smv + base + geo.off32(c), o[c]);
  }
}

// gate on tile bit 0 (lane-mixing): K = 4: bits (hv + 1, 0), two vectors; K = 2: bit 0, one vector
template <int K>
__device__ __forceinline__ void unpack_lane_mix(const V4 (&v)[K / 2], cplx_t (&a)[K]) {
#pragma unroll
  for (int h = 0; h < K / 2; h++) {
    a[2 * h] = make_float2(v[h].re.x, v[h].im.x);
    a[2 * h + 1] = make_float2(v[h].re.y, v[h].im.y);
  }
}
template <int K>
__device__ __forceinline__ void pack_lane_mix(V4 (&v)[K / 2], const cplx_t (&a)[K]) {
#pragma unroll
  for (int h = 0; h < K / 2; h++) {
    v[h].re = make_float2(a[2 * h].x, a[2 * h + 1].x);
    v[h].im = make_float2(a[2 * h].y, a[2 * h + 1].y);
  }
}

template <int NT, int K>
__device__ __forceinline__ void tile_apply_mix(vec_t* smv, int hv, int nitems, const GateMat& G) {
  for (int i0 = 0; i0 < nitems; i0 += NT) {
    const uint32_t i = (uint32_t)(i0 + threadIdx.x);
    const uint32_t base = (K == 4) ? ins0_32(i, hv) : i;
    V4 v[K / 2];
#pragma unroll
    for (int h = 0; h < K / 2; h++) v[h] = ld4(smv + base + ((uint32_t)h << hv));
    cplx_t a[K];
    unpack_lane_mix<K>(v, a);
    mv<K>(G, a);
    pack_lane_mix<K>(v, a);
#pragma unroll
    for (int h = 0; h < K / 2; h++) st4(smv + base + ((uint32_t)h << hv), v[h]);
  }
}

// Diagonal gates, entries in (hi,lo) order d[2 bit_a + bit_b], a > b (host: tile_matrix).
//  * b >= 1: the four vectors of a quad item are the four settings of (bit a, bit b): vector c takes
//    the CTA-uniform entry d[c] for both lanes -- broadcast operands, compile-time indices.
//  * b == 0: the lanes ARE bit b: the two vectors of a pair item over bit a take the lane pairs
//    (d[2h], d[2h+1]).
// (A per-amplitude selection of the entry compiles to divergent branches over the constant bank:
// measured 0.44 ms per diagonal reverse step at 26 q against 0.11 ms for a dense one-qubit gate.)
__device__ __forceinline__ V4 cmul_bc(const V4& v, const float dr, const float di) {
  V4 o;
  o.re = ffma2(bc2(-di), v.im, fmul2(bc2(dr), v.re));
  o.im = ffma2(bc2(dr), v.im, fmul2(bc2(di), v.re));
  return o;
}
__device__ __forceinline__ V4 cmul_pair(const V4& v, const float2 dr, const float2 di) {
  V4 o;
  o.re = ffma2(neg2(di), v.im, fmul2(dr, v.re));
  o.im = ffma2(dr, v.im, fmul2(di, v.re));
  return o;
}

template <int NT>
__device__ __forceinline__ void tile_diag_soa_hh(vec_t* smv, const GeoQ2HH& geo, int nitems, const GateMat& D) {
  for (int i0 = 0; i0 < nitems; i0 += NT) {
    const uint32_t base = geo.base32((uint32_t)(i0 + threadIdx.x));
#pragma unroll
    for (int c = 0; c < 4; c++) st4(smv + base + geo.off32(c), cmul_bc(ld4(smv + base + geo.off32(c)), D.re[c], D.im[c]));
  }
}

template <int NT>
__device__ __forceinline__ void tile_diag_soa_lh(vec_t* smv, int hv, int nitems, const GateMat& D) {
  float2 dr[2], di[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    dr[h] = make_float2(D.re[2 * h], D.re[2 * h + 1]);
    di[h] = make_float2(D.im[2 * h], D.im[2 * h + 1]);
  }
  for (int i0 = 0; i0 < nitems; i0 += NT) {
    const uint32_t base = ins0_32((uint32_t)(i0 + threadIdx.x), hv);
#pragma unroll
    for (int h = 0; h < 2; h++)
      st4(smv + base + ((uint32_t)h << hv), cmul_pair(ld4(smv + base + ((uint32_t)h << hv)), dr[h], di[h]));
  }
}

__global__ void __launch_bounds__(QDC_TILE_NT_F, 6)
    k_tile_fwd_soa(cplx_t* __restrict__ state, const __grid_constant__ TileFwdParams p) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  vec_t* smv = (vec_t*)tile_smem;
  const int nvec = 1 << (p.geo.T - QDC_LV);
  TileAddr<QDC_TILE_NT_F> ta;
  ta.init(p.geo);
  for (uint64_t tile = blockIdx.x; tile < p.geo.ntiles; tile += gridDim.x) {
    const uint64_t tbase = p.geo.tile(tile) >> QDC_LV;
    tile_io_soa<QDC_TILE_NT_F, true>((vec_t*)state, smv, ta, tbase);
    __syncthreads();
    for (int g = 0; g < p.ngates; g++) {
      const TileGateF& G = p.g[g];
      if (G.type == TG_Q2) {
        if (G.b == 0) {
          tile_apply_mix<QDC_TILE_NT_F, 4>(smv, G.a - 1, nvec / 2, G.m);
        } else {
          GeoQ2HH geo;
          geo.lv = G.b - QDC_LV;
          geo.hv = G.a - QDC_LV;
          tile_apply_soa<QDC_TILE_NT_F>(smv, geo, nvec / 4, G.m);
        }
      } else if (G.type == TG_Q1) {
        if (G.a == 0) {
          tile_apply_mix<QDC_TILE_NT_F, 2>(smv, 0, nvec, G.m);
        } else {
          GeoQ1H geo;
          geo.pv = G.a - QDC_LV;
          tile_apply_soa<QDC_TILE_NT_F>(smv, geo, nvec / 2, G.m);
        }
      } else if (G.b == 0) {
        tile_diag_soa_lh<QDC_TILE_NT_F>(smv, G.a - 1, nvec / 2, G.m);
      } else {
        GeoQ2HH geo;
        geo.lv = G.b - QDC_LV;
        geo.hv = G.a - QDC_LV;
        tile_diag_soa_hh<QDC_TILE_NT_F>(smv, geo, nvec / 4, G.m);
      }
      __syncthreads();
    }
    tile_io_soa<QDC_TILE_NT_F, false>((vec_t*)state, smv, ta, tbase);
    __syncthreads();
  }
}

// ---------------------------------------------------------------- reverse
template <int NT, class Geo>
__device__ __forceinline__ void tile_rev_soa(vec_t* smf, vec_t* smb, const Geo& geo, int nitems, const TileGateB& G,
                                             real_t (&acc)[32]) {
  constexpr int K = Geo::K;
  float2 are[16], aim[16];
#pragma unroll
  for (int k = 0; k < 16; k++) are[k] = aim[k] = make_float2(0.f, 0.f);
  for (int i0 = 0; i0 < nitems; i0 += NT) {  // uniform trip count
    const uint32_t base = geo.base32((uint32_t)(i0 + threadIdx.x));
    V4 f[K], a[K], b[K], bo[K];
#pragma unroll
    for (int c = 0; c < K; c++) {
      f[c] = ld4(smf + base + geo.off32(c));
      b[c] = ld4(smb + base + geo.off32(c));
    }
    mv_soa<K>(G.inv, f, a);                               // un-compute the state
#pragma unroll
    for (int c = 0; c < K; c++) st4(smf + base + geo.off32(c), a[c]);
    if (G.slot >= 0) outer_soa<K>(b, a, are, aim);        // gradient from (pre-gate state, post-gate adjoint)
    mv_soa<K>(G.tr, b, bo);                               // pull the adjoint back
#pragma unroll
    for (int c = 0; c < K; c++) st4(smb + base + geo.off32(c), bo[c]);
  }
#pragma unroll
  for (int k = 0; k < K * K; k++) {
    acc[2 * k] = are[k].x + are[k].y;
    acc[2 * k + 1] = aim[k].x + aim[k].y;
  }
}

template <int NT, int K>
__device__ __forceinline__ void tile_rev_mix(vec_t* smf, vec_t* smb, int hv, int nitems, const TileGateB& G,
                                             real_t (&acc)[32]) {
  for (int i0 = 0; i0 < nitems; i0 += NT) {
    const uint32_t i = (uint32_t)(i0 + threadIdx.x);
    const uint32_t base = (K == 4) ? ins0_32(i, hv) : i;
    V4 vf[K / 2], vb[K / 2];
#pragma unroll
    for (int h = 0; h < K / 2; h++) {
      vf[h] = ld4(smf + base + ((uint32_t)h << hv));
      vb[h] = ld4(smb + base + ((uint32_t)h << hv));
    }
    cplx_t a[K], b[K];
    unpack_lane_mix<K>(vf, a);
    unpack_lane_mix<K>(vb, b);
    mv<K>(G.inv, a);
    if (G.slot >= 0) outer_tile<K>(b, a, acc);
    mv<K>(G.tr, b);
    pack_lane_mix<K>(vf, a);
    pack_lane_mix<K>(vb, b);
#pragma unroll
    for (int h = 0; h < K / 2; h++) {
      st4(smf + base + ((uint32_t)h << hv), vf[h]);
      st4(smb + base + ((uint32_t)h << hv), vb[h]);
    }
  }
}

template <int NT>
__device__ __forceinline__ void tile_rev_diag_soa_hh(vec_t* smf, vec_t* smb, const GeoQ2HH& geo, int nitems,
                                                     const TileGateB& G, real_t (&acc)[32]) {
  float2 are[4], aim[4];
#pragma unroll
  for (int c = 0; c < 4; c++) are[c] = aim[c] = make_float2(0.f, 0.f);
  for (int i0 = 0; i0 < nitems; i0 += NT) {
    const uint32_t base = geo.base32((uint32_t)(i0 + threadIdx.x));
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const V4 f = cmul_bc(ld4(smf + base + geo.off32(c)), G.inv.re[c], G.inv.im[c]);
      const V4 b = ld4(smb + base + geo.off32(c));
      if (G.slot >= 0) {
        are[c] = ffma2(neg2(b.im), f.im, ffma2(b.re, f.re, are[c]));
        aim[c] = ffma2(b.im, f.re, ffma2(b.re, f.im, aim[c]));
      }
      st4(smf + base + geo.off32(c), f);
      st4(smb + base + geo.off32(c), cmul_bc(b, G.tr.re[c], G.tr.im[c]));
    }
  }
#pragma unroll
  for (int c = 0; c < 4; c++) {
    acc[2 * c] = are[c].x + are[c].y;
    acc[2 * c + 1] = aim[c].x + aim[c].y;
  }
}

template <int NT>
__device__ __forceinline__ void tile_rev_diag_soa_lh(vec_t* smf, vec_t* smb, int hv, int nitems, const TileGateB& G,
                                                     real_t (&acc)[32]) {
  float2 ir[2], ii[2], dr[2], di[2], are[2], aim[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    ir[h] = make_float2(G.inv.re[2 * h], G.inv.re[2 * h + 1]);
    ii[h] = make_float2(G.inv.im[2 * h], G.inv.im[2 * h + 1]);
    dr[h] = make_float2(G.tr.re[2 * h], G.tr.re[2 * h + 1]);
    di[h] = make_float2(G.tr.im[2 * h], G.tr.im[2 * h + 1]);
    are[h] = aim[h] = make_float2(0.f, 0.f);
  }
  for (int i0 = 0; i0 < nitems; i0 += NT) {
    const uint32_t base = ins0_32((uint32_t)(i0 + threadIdx.x), hv);
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const V4 f = cmul_pair(ld4(smf + base + ((uint32_t)h << hv)), ir[h], ii[h]);
      const V4 b = ld4(smb + base + ((uint32_t)h << hv));
      if (G.slot >= 0) {
        are[h] = ffma2(neg2(b.im), f.im, ffma2(b.re, f.re, are[h]));
        aim[h] = ffma2(b.im, f.re, ffma2(b.re, f.im, aim[h]));
      }
      st4(smf + base + ((uint32_t)h << hv), f);
      st4(smb + base + ((uint32_t)h << hv), cmul_pair(b, dr[h], di[h]));
    }
  }
#pragma unroll
  for (int h = 0; h < 2; h++) {  // lane e of pair h is the entry j = 2 h + e: no lane fold
    acc[2 * (2 * h)] = are[h].x;
    acc[2 * (2 * h) + 1] = aim[h].x;
    acc[2 * (2 * h + 1)] = are[h].y;
    acc[2 * (2 * h + 1) + 1] = aim[h].y;
  }
}

// partials: [gridDim.x][ngates][32] doubles (same contract as k_tile_bwd)
__global__ void __launch_bounds__(QDC_TILE_NT_B, 3)
    k_tile_bwd_soa(cplx_t* __restrict__ fwd, cplx_t* __restrict__ bwd, const __grid_constant__ TileBwdParams p,
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
    tile_io_soa<QDC_TILE_NT_B, true>((vec_t*)fwd, smf, ta, tbase);
    tile_io_soa<QDC_TILE_NT_B, true>((vec_t*)bwd, smb, ta, tbase);
    __syncthreads();
    for (int g = 0; g < p.ngates; g++) {
      const TileGateB& G = p.g[g];
      real_t acc[32];
#pragma unroll
      for (int k = 0; k < 32; k++) acc[k] = 0;
      if (G.type == TG_Q2) {
        if (G.b == 0) {
          tile_rev_mix<QDC_TILE_NT_B, 4>(smf, smb, G.a - 1, nvec / 2, G, acc);
        } else {
          GeoQ2HH geo;
          geo.lv = G.b - QDC_LV;
          geo.hv = G.a - QDC_LV;
          tile_rev_soa<QDC_TILE_NT_B>(smf, smb, geo, nvec / 4, G, acc);
        }
      } else if (G.type == TG_Q1) {
        if (G.a == 0) {
          tile_rev_mix<QDC_TILE_NT_B, 2>(smf, smb, 0, nvec, G, acc);
        } else {
          GeoQ1H geo;
          geo.pv = G.a - QDC_LV;
          tile_rev_soa<QDC_TILE_NT_B>(smf, smb, geo, nvec / 2, G, acc);
        }
      } else if (G.b == 0) {
        tile_rev_diag_soa_lh<QDC_TILE_NT_B>(smf, smb, G.a - 1, nvec / 2, G, acc);
      } else {
        GeoQ2HH geo;
        geo.lv = G.b - QDC_LV;
        geo.hv = G.a - QDC_LV;
        tile_rev_diag_soa_hh<QDC_TILE_NT_B>(smf, smb, geo, nvec / 4, G, acc);
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
    tile_io_soa<QDC_TILE_NT_B, false>((vec_t*)fwd, smf, ta, tbase);
    tile_io_soa<QDC_TILE_NT_B, false>((vec_t*)bwd, smb, ta, tbase);
    __syncthreads();
  }
  for (int i = threadIdx.x; i < p.ngates * 32; i += QDC_TILE_NT_B)
    partials[(size_t)blockIdx.x * p.ngates * 32 + i] = sm_acc[i];
}

#endif  // !QDC_F64
