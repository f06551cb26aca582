// Register-blocked tiled passes.
//
// Same tiling as tile_kernels.cuh (persistent CTA, tile of 2^T amplitudes in
// shared memory, one HBM round trip per pass), but the gates of a pass are
// grouped by the scheduler into GROUPS acting inside <= 4 tile positions.  For
// each group a thread pulls a block of 16 amplitudes (the 2^4 settings of the
// group's positions) into REGISTERS, applies every gate of the group there
// (fully unrolled, one switch case per position pair so all register indices
// are compile-time), and writes the block back.  Shared-memory traffic, gate
// matrix loads (LDC) and barriers are paid once per group (~3 gates for
// brickwork) instead of once per gate -- ncu showed the per-gate version
// spending as many cycles on LDS/STS + LDC + barriers as on FFMAs.
#pragma once
#include "tile_kernels.cuh"

#define QDC_TILE_NT_BRB 128  // threads per CTA of the register-blocked backward kernel
#ifdef QDC_F64
#define QDC_RB_FWD_MINB 4   // 16 double-complex amplitudes per thread: 128 registers
#else
#define QDC_RB_FWD_MINB 6
#endif
#define QDC_RB 4            // positions per register block
#define QDC_RB_AMPS 16
#define QDC_RB_MAXGRP 32  // >= the largest number of gates in a pass (every gate its own group in the worst case)

struct RbGroup {
  BitDeposit map;         // block index (T-4 bits) -> tile-local element index of the block's element 0
  int bit[QDC_RB];        // tile-local positions of the block, ascending
  int first, count;       // gates of this group: params.g[first .. first+count)
};

// gate codes: 0..5 dense q2 on block positions (1,0),(2,0),(2,1),(3,0),(3,1),(3,2);
// 6..9 dense q1 on block position 0..3; 10..15 diagonal on the same six pairs
// (entries in (hi,lo) order: d[2 bit_hi + bit_lo]).
struct TileFwdRbParams {
  TileGeo geo;
  int ngroups, ngates;
  RbGroup grp[QDC_RB_MAXGRP];
  TileGateF g[QDC_TILE_MAXG_F];  // .type = code
};
struct TileBwdRbParams {
  TileGeo geo;
  int ngroups, ngates;
  RbGroup grp[QDC_RB_MAXGRP];
  TileGateB g[QDC_TILE_MAXG_B];  // .type = code, .slot = gradient slot or -1
};

__host__ __device__ constexpr int rb_ins0(int i, int pos) { return ((i >> pos) << (pos + 1)) | (i & ((1 << pos) - 1)); }

// ------------------------------------------------------------ forward ops
template <int GA, int GB>
__device__ __forceinline__ void rb_q2(cplx_t (&x)[QDC_RB_AMPS], const GateMat& G) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int i0 = rb_ins0(rb_ins0(r, GB), GA);
    cplx_t a[4] = {x[i0], x[i0 + (1 << GB)], x[i0 + (1 << GA)], x[i0 + (1 << GA) + (1 << GB)]};
    mv<4>(G, a);
    x[i0] = a[0];
    x[i0 + (1 << GB)] = a[1];
    x[i0 + (1 << GA)] = a[2];
    x[i0 + (1 << GA) + (1 << GB)] = a[3];
  }
}

template <int P>
__device__ __forceinline__ void rb_q1(cplx_t (&x)[QDC_RB_AMPS], const GateMat& G) {
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int i0 = rb_ins0(r, P);
    cplx_t a[2] = {x[i0], x[i0 + (1 << P)]};
    mv<2>(G, a);
    x[i0] = a[0];
    x[i0 + (1 << P)] = a[1];
  }
}

template <int GA, int GB>
__device__ __forceinline__ void rb_diag(cplx_t (&x)[QDC_RB_AMPS], const GateMat& G) {
#pragma unroll
  for (int k = 0; k < QDC_RB_AMPS; k++) {
    const int j = 2 * ((k >> GA) & 1) + ((k >> GB) & 1);
    const real_t xr = x[k].x * gm_re(G, j) - x[k].y * gm_im(G, j), xi = x[k].x * gm_im(G, j) + x[k].y * gm_re(G, j);
    x[k].x = xr;
    x[k].y = xi;
  }
}

__device__ __forceinline__ void rb_apply(int code, cplx_t (&x)[QDC_RB_AMPS], const GateMat& G) {
  switch (code) {
    case 0: rb_q2<1, 0>(x, G); break;
    case 1: rb_q2<2, 0>(x, G); break;
    case 2: rb_q2<2, 1>(x, G); break;
    case 3: rb_q2<3, 0>(x, G); break;
    case 4: rb_q2<3, 1>(x, G); break;
    case 5: rb_q2<3, 2>(x, G); break;
    case 6: rb_q1<0>(x, G); break;
    case 7: rb_q1<1>(x, G); break;
    case 8: rb_q1<2>(x, G); break;
    case 9: rb_q1<3>(x, G); break;
    case 10: rb_diag<1, 0>(x, G); break;
    case 11: rb_diag<2, 0>(x, G); break;
    case 12: rb_diag<2, 1>(x, G); break;
    case 13: rb_diag<3, 0>(x, G); break;
    case 14: rb_diag<3, 1>(x, G); break;
    default: rb_diag<3, 2>(x, G); break;
  }
}

// block element k lives at tile-local element  base + sum_i bit_i(k) << grp.bit[i]
__device__ __forceinline__ void rb_offsets(const RbGroup& G, uint32_t (&off)[QDC_RB_AMPS]) {
#pragma unroll
  for (int k = 0; k < QDC_RB_AMPS; k++) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < QDC_RB; i++) o += (uint32_t)((k >> i) & 1) << G.bit[i];
    off[k] = o;
  }
}

__global__ void __launch_bounds__(QDC_TILE_NT_F, QDC_RB_FWD_MINB)
    k_tile_fwd_rb(cplx_t* __restrict__ state, const __grid_constant__ TileFwdRbParams p) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  vec_t* smv = (vec_t*)tile_smem;
  cplx_t* sme = (cplx_t*)tile_smem;
  const int nblocks = 1 << (p.geo.T - QDC_RB);
  TileAddr<QDC_TILE_NT_F> ta;
  ta.init(p.geo);
  for (uint64_t tile = blockIdx.x; tile < p.geo.ntiles; tile += gridDim.x) {
    const uint64_t tbase = p.geo.tile(tile) >> QDC_LV;
    tile_io<QDC_TILE_NT_F, true>((vec_t*)state, smv, ta, tbase);
    __syncthreads();
    for (int gi = 0; gi < p.ngroups; gi++) {
      const RbGroup& G = p.grp[gi];
      uint32_t off[QDC_RB_AMPS];
      rb_offsets(G, off);
      for (int j0 = 0; j0 < nblocks; j0 += QDC_TILE_NT_F) {  // uniform trip count: nblocks % CTA size == 0
        const int j = j0 + threadIdx.x;
        const uint32_t base = (uint32_t)G.map((uint64_t)j);
        cplx_t x[QDC_RB_AMPS];
#pragma unroll
        for (int k = 0; k < QDC_RB_AMPS; k++) x[k] = sme[base + off[k]];
        for (int q = 0; q < G.count; q++) {
          const TileGateF& M = p.g[G.first + q];
          rb_apply(M.type, x, M.m);
        }
#pragma unroll
        for (int k = 0; k < QDC_RB_AMPS; k++) sme[base + off[k]] = x[k];
      }
      __syncthreads();
    }
    tile_io<QDC_TILE_NT_F, false>((vec_t*)state, smv, ta, tbase);
    __syncthreads();
  }
}

// ----------------------------------------------------------- backward ops
template <int GA, int GB>
__device__ __forceinline__ void rb_q2_rev(cplx_t (&xf)[QDC_RB_AMPS], cplx_t (&xb)[QDC_RB_AMPS], const TileGateB& M,
                                          real_t (&acc)[32]) {
  // un-compute the state first (only the inverse matrix live), then gradient + adjoint
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int i0 = rb_ins0(rb_ins0(r, GB), GA);
    cplx_t a[4] = {xf[i0], xf[i0 + (1 << GB)], xf[i0 + (1 << GA)], xf[i0 + (1 << GA) + (1 << GB)]};
    mv<4>(M.inv, a);
    xf[i0] = a[0];
    xf[i0 + (1 << GB)] = a[1];
    xf[i0 + (1 << GA)] = a[2];
    xf[i0 + (1 << GA) + (1 << GB)] = a[3];
  }
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int i0 = rb_ins0(rb_ins0(r, GB), GA);
    cplx_t b[4] = {xb[i0], xb[i0 + (1 << GB)], xb[i0 + (1 << GA)], xb[i0 + (1 << GA) + (1 << GB)]};
    if (M.slot >= 0) {
      const cplx_t a[4] = {xf[i0], xf[i0 + (1 << GB)], xf[i0 + (1 << GA)], xf[i0 + (1 << GA) + (1 << GB)]};
      outer_tile<4>(b, a, acc);
    }
    mv<4>(M.tr, b);
    xb[i0] = b[0];
    xb[i0 + (1 << GB)] = b[1];
    xb[i0 + (1 << GA)] = b[2];
    xb[i0 + (1 << GA) + (1 << GB)] = b[3];
  }
}

template <int P>
__device__ __forceinline__ void rb_q1_rev(cplx_t (&xf)[QDC_RB_AMPS], cplx_t (&xb)[QDC_RB_AMPS], const TileGateB& M,
                                          real_t (&acc)[32]) {
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int i0 = rb_ins0(r, P);
    cplx_t a[2] = {xf[i0], xf[i0 + (1 << P)]};
    mv<2>(M.inv, a);
    xf[i0] = a[0];
    xf[i0 + (1 << P)] = a[1];
    cplx_t b[2] = {xb[i0], xb[i0 + (1 << P)]};
    if (M.slot >= 0) outer_tile<2>(b, a, acc);
    mv<2>(M.tr, b);
    xb[i0] = b[0];
    xb[i0 + (1 << P)] = b[1];
  }
}

template <int GA, int GB>
__device__ __forceinline__ void rb_diag_rev(cplx_t (&xf)[QDC_RB_AMPS], cplx_t (&xb)[QDC_RB_AMPS], const TileGateB& M,
                                            real_t (&acc)[32]) {
#pragma unroll
  for (int k = 0; k < QDC_RB_AMPS; k++) {
    const int j = 2 * ((k >> GA) & 1) + ((k >> GB) & 1);
    const real_t fx = xf[k].x * gm_re(M.inv, j) - xf[k].y * gm_im(M.inv, j), fy = xf[k].x * gm_im(M.inv, j) + xf[k].y * gm_re(M.inv, j);
    xf[k].x = fx;
    xf[k].y = fy;
    const real_t bx = xb[k].x, by = xb[k].y;
    if (M.slot >= 0) {
      acc[2 * j] += bx * fx - by * fy;
      acc[2 * j + 1] += bx * fy + by * fx;
    }
    xb[k].x = bx * gm_re(M.tr, j) - by * gm_im(M.tr, j);
    xb[k].y = bx * gm_im(M.tr, j) + by * gm_re(M.tr, j);
  }
}

__device__ __forceinline__ void rb_apply_rev(int code, cplx_t (&xf)[QDC_RB_AMPS], cplx_t (&xb)[QDC_RB_AMPS],
                                             const TileGateB& M, real_t (&acc)[32]) {
  switch (code) {
    case 0: rb_q2_rev<1, 0>(xf, xb, M, acc); break;
    case 1: rb_q2_rev<2, 0>(xf, xb, M, acc); break;
    case 2: rb_q2_rev<2, 1>(xf, xb, M, acc); break;
    case 3: rb_q2_rev<3, 0>(xf, xb, M, acc); break;
    case 4: rb_q2_rev<3, 1>(xf, xb, M, acc); break;
    case 5: rb_q2_rev<3, 2>(xf, xb, M, acc); break;
    case 6: rb_q1_rev<0>(xf, xb, M, acc); break;
    case 7: rb_q1_rev<1>(xf, xb, M, acc); break;
    case 8: rb_q1_rev<2>(xf, xb, M, acc); break;
    case 9: rb_q1_rev<3>(xf, xb, M, acc); break;
    case 10: rb_diag_rev<1, 0>(xf, xb, M, acc); break;
    case 11: rb_diag_rev<2, 0>(xf, xb, M, acc); break;
    case 12: rb_diag_rev<2, 1>(xf, xb, M, acc); break;
    case 13: rb_diag_rev<3, 0>(xf, xb, M, acc); break;
    case 14: rb_diag_rev<3, 1>(xf, xb, M, acc); break;
    default: rb_diag_rev<3, 2>(xf, xb, M, acc); break;
  }
}

// Shared memory: [fwd tile][bwd tile][sm_acc: MAXG_B x 32 double][part: warps x MAXG_B x 32 real_t]
// Gradient partials: thread -> warp reduce-scatter (lane j owns value j) -> the
// warp's private row of `part` (no atomics, fixed order) -> at the end of every
// tile warp-ordered sum into the CTA's double accumulators.
__global__ void __launch_bounds__(QDC_TILE_NT_BRB, 3)
    k_tile_bwd_rb(cplx_t* __restrict__ fwd, cplx_t* __restrict__ bwd, const __grid_constant__ TileBwdRbParams p,
                  double* __restrict__ partials) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  const int nvec = 1 << (p.geo.T - QDC_LV);
  vec_t* smf = (vec_t*)tile_smem;
  vec_t* smb = smf + nvec;
  cplx_t* ef = (cplx_t*)smf;
  cplx_t* eb = (cplx_t*)smb;
  double* sm_acc = (double*)(smb + nvec);
  real_t* part = (real_t*)(sm_acc + QDC_TILE_MAXG_B * 32);
  constexpr int NW = QDC_TILE_NT_BRB / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nblocks = 1 << (p.geo.T - QDC_RB);
  for (int i = threadIdx.x; i < QDC_TILE_MAXG_B * 32; i += QDC_TILE_NT_BRB) sm_acc[i] = 0.0;
  for (int i = threadIdx.x; i < NW * QDC_TILE_MAXG_B * 32; i += QDC_TILE_NT_BRB) part[i] = 0;
  __syncthreads();
  TileAddr<QDC_TILE_NT_BRB> ta;
  ta.init(p.geo);
  real_t* mypart = part + (size_t)warp * QDC_TILE_MAXG_B * 32;
  for (uint64_t tile = blockIdx.x; tile < p.geo.ntiles; tile += gridDim.x) {
    const uint64_t tbase = p.geo.tile(tile) >> QDC_LV;
    tile_io<QDC_TILE_NT_BRB, true>((vec_t*)fwd, smf, ta, tbase);
    tile_io<QDC_TILE_NT_BRB, true>((vec_t*)bwd, smb, ta, tbase);
    __syncthreads();
    for (int gi = 0; gi < p.ngroups; gi++) {
      const RbGroup& G = p.grp[gi];
      uint32_t off[QDC_RB_AMPS];
      rb_offsets(G, off);
      for (int j0 = 0; j0 < nblocks; j0 += QDC_TILE_NT_BRB) {  // uniform trip count: nblocks % CTA size == 0
        const int j = j0 + threadIdx.x;
        const uint32_t base = (uint32_t)G.map((uint64_t)j);
        cplx_t xf[QDC_RB_AMPS], xb[QDC_RB_AMPS];
#pragma unroll
        for (int k = 0; k < QDC_RB_AMPS; k++) {
          xf[k] = ef[base + off[k]];
          xb[k] = eb[base + off[k]];
        }
        for (int q = 0; q < G.count; q++) {
          const int gidx = G.first + q;
          const TileGateB& M = p.g[gidx];
          real_t acc[32];
#pragma unroll
          for (int k = 0; k < 32; k++) acc[k] = 0;
          rb_apply_rev(M.type, xf, xb, M, acc);
          if (M.slot >= 0) {
            double d = 0.0;
            warp_flush<32>(acc, d, lane);
            mypart[gidx * 32 + lane] += (real_t)d;
          }
        }
#pragma unroll
        for (int k = 0; k < QDC_RB_AMPS; k++) {
          ef[base + off[k]] = xf[k];
          eb[base + off[k]] = xb[k];
        }
      }
      __syncthreads();
    }
    // fold this tile's per-warp partials into the CTA accumulators (fixed order)
    for (int i = threadIdx.x; i < p.ngates * 32; i += QDC_TILE_NT_BRB) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < NW; w++) {
        s += (double)part[(size_t)w * QDC_TILE_MAXG_B * 32 + i];
        part[(size_t)w * QDC_TILE_MAXG_B * 32 + i] = 0;
      }
      sm_acc[i] += s;
    }
    tile_io<QDC_TILE_NT_BRB, false>((vec_t*)fwd, smf, ta, tbase);
    tile_io<QDC_TILE_NT_BRB, false>((vec_t*)bwd, smb, ta, tbase);
    __syncthreads();
  }
  for (int i = threadIdx.x; i < p.ngates * 32; i += QDC_TILE_NT_BRB)
    partials[(size_t)blockIdx.x * p.ngates * 32 + i] = sm_acc[i];
}

// ------------------------------------------------------------ host side
static inline int rb_pair_code(int ga, int gb) {  // ga > gb, block-local
  static const int tab[4][4] = {{-1, -1, -1, -1}, {0, -1, -1, -1}, {1, 2, -1, -1}, {3, 4, 5, -1}};
  return tab[ga][gb];
}

// Fill the groups of a tile pass.  `tpos` maps physical position -> tile-local
// position; `code_of[k]` receives the switch code of tile gate k (plan order).
static inline const char* rb_make_groups(const qdc::Plan& plan, const qdc::Step& t, const std::vector<int>& tpos,
                                         int T, const std::vector<Inst>& insts, RbGroup* grp, int* ngroups,
                                         std::vector<int>* code_of, std::vector<bool>* swap_of) {
  if (t.grp_count > QDC_RB_MAXGRP) return qdc_errf("tile pass holds too many register-block groups.");
  // the kernels give every thread of the CTA its own 2^4-amplitude block per sweep: fewer blocks than threads
  // would make threads share blocks (make_tile_geo keeps tiles at >= 2^11 amplitudes for this reason)
  if ((1 << (T - QDC_RB)) < QDC_TILE_NT_F || (1 << (T - QDC_RB)) < QDC_TILE_NT_BRB)
    return qdc_errf("tile of 2^%d amplitudes is too small for the register-blocked kernels.", T);
  *ngroups = t.grp_count;
  code_of->assign(t.count, -1);
  swap_of->assign(t.count, false);
  for (int gi = 0; gi < t.grp_count; gi++) {
    const qdc::Group& g = plan.groups[t.grp_first + gi];
    RbGroup& G = grp[gi];
    std::vector<int> bits;
    for (int k = 0; k < g.nbits; k++) bits.push_back(tpos[g.bits[k]]);
    // pad to 4 positions with the highest unused tile positions (keeps the low, bank-deciding bits free)
    for (int pos = T - 1; (int)bits.size() < QDC_RB && pos >= 0; pos--)
      if (std::find(bits.begin(), bits.end(), pos) == bits.end()) bits.push_back(pos);
    std::sort(bits.begin(), bits.end());
    for (int k = 0; k < QDC_RB; k++) G.bit[k] = bits[k];
    std::vector<int> rest;
    for (int pos = 0; pos < T; pos++)
      if (std::find(bits.begin(), bits.end(), pos) == bits.end()) rest.push_back(pos);
    if (!make_deposit(rest, &G.map)) return qdc_errf("register-block map too fragmented.");
    G.first = g.first - t.first;
    G.count = g.count;
    for (int q = 0; q < g.count; q++) {
      const int k = g.first - t.first + q;
      const qdc::Step& st = plan.tile_steps[g.first + q];
      const int kind = insts[st.inst].kind;
      auto local = [&](int phys) {
        const int tp = tpos[phys];
        for (int i = 0; i < QDC_RB; i++)
          if (G.bit[i] == tp) return i;
        return -1;
      };
      if (kind_is_q1(kind)) {
        (*code_of)[k] = 6 + local(st.p2);
      } else {
        int a = local(st.p2), b = local(st.p1);
        const bool swap = a < b;  // pos2 sits on the lower position: present the matrix in (hi,lo) order
        if (swap) std::swap(a, b);
        (*swap_of)[k] = swap;
        (*code_of)[k] = (kind_is_diag(kind) ? 10 : 0) + rb_pair_code(a, b);
      }
    }
  }
  return nullptr;
}

// diagonal entries in (hi,lo) order
static inline void rb_diag_entries(const cplx_t* d, bool swap, bool conj, GateMat* out) {
  real_t re[16], im[16];
  for (int i = 0; i < 16; i++) re[i] = im[i] = 0;
  for (int j = 0; j < 4; j++) {
    const int src = swap ? perm2(j) : j;
    re[j] = d[src].x;
    im[j] = conj ? -d[src].y : d[src].y;
  }
  gm_fill(*out, re, im);
}

// Build the gate table of a pass for the register-blocked kernels.
// reverse = false: forward matrices in plan order.
// reverse = true : groups and gates reversed; `inverse_only` selects the
//                  un-compute-only variant (forward kernel, inverse matrices).
template <class Params>
static const char* rb_fill_groups(const qdc::Plan& plan, const qdc::Step& t, const std::vector<int>& tpos, int T,
                                  const std::vector<Inst>& insts, bool reverse, Params* p, std::vector<int>* order,
                                  std::vector<int>* code_of, std::vector<bool>* swap_of) {
  static thread_local RbGroup tmp[QDC_RB_MAXGRP];
  int ng = 0;
  QDC_TRY(rb_make_groups(plan, t, tpos, T, insts, tmp, &ng, code_of, swap_of));
  p->ngroups = ng;
  p->ngates = t.count;
  order->clear();
  int next = 0;
  for (int gi = 0; gi < ng; gi++) {
    const RbGroup& src = tmp[reverse ? ng - 1 - gi : gi];
    RbGroup& dst = p->grp[gi];
    dst = src;
    dst.first = next;
    for (int q = 0; q < src.count; q++) order->push_back(reverse ? src.first + src.count - 1 - q : src.first + q);
    next += src.count;
  }
  return nullptr;
}

inline const char* Circuit::run_tile_forward_rb(const qdc::Step& t, const std::vector<const cplx_t*>& gp,
                                                bool uncompute) {
  static thread_local TileFwdRbParams p;
  std::vector<int> tpos, order, code_of;
  std::vector<bool> swap_of;
  QDC_TRY(make_tile_geo(plan_, t, n_loc_, 0, &p.geo, &tpos));
  if (t.count > QDC_TILE_MAXG_F) return qdc_errf("tile pass holds too many gates.");
  QDC_TRY(rb_fill_groups(plan_, t, tpos, p.geo.T, insts_, uncompute, &p, &order, &code_of, &swap_of));
  for (int k = 0; k < t.count; k++) {
    const int src = order[k];
    const qdc::Step& st = plan_.tile_steps[t.first + src];
    const int kind = insts_[st.inst].kind;
    TileGateF& G = p.g[k];
    G.type = code_of[src];
    G.a = G.b = G.pad = 0;
    const int form = uncompute ? (kind_is_nonu(kind) ? FORM_INV : FORM_CONJ_TR) : FORM_PLAIN;
    if (kind_is_diag(kind)) {
      rb_diag_entries(gp[st.inst], swap_of[src], uncompute, &G.m);
    } else {
      QDC_TRY(tile_matrix(gp[st.inst], kind, form, swap_of[src], &G.m));
    }
  }
  const size_t smem = sizeof(cplx_t) << p.geo.T;
  int grid = 0;
  QDC_TRY(tile_grid((const void*)k_tile_fwd_rb, QDC_TILE_NT_F, smem, p.geo.ntiles, &grid));
  cudaEvent_t pa = nullptr;
  if (prof_.on) pa = prof_.begin(stream_);
  k_tile_fwd_rb<<<grid, QDC_TILE_NT_F, smem, stream_>>>(state_, p);
  QDC_CUDA(cudaGetLastError());
  if (prof_.on) prof_.end(stream_, uncompute ? CAT_UNCOMPUTE : CAT_TILE_FWD, pa, 2ull * t.count * bytes());
  stats_.kernel_launches += 1;
  stats_.hbm_passes += 1;
  stats_.algorithmic_bytes += 2ull * t.count * bytes();
  return nullptr;
}

inline const char* Circuit::run_tile_backward_rb(const qdc::Step& t, const std::vector<const cplx_t*>& gp,
                                                 const std::vector<long>& vslot) {
  static thread_local TileBwdRbParams p;
  std::vector<int> tpos, order, code_of;
  std::vector<bool> swap_of;
  QDC_TRY(make_tile_geo(plan_, t, n_loc_, 0, &p.geo, &tpos));
  if (t.count > QDC_TILE_MAXG_B) return qdc_errf("tile pass holds too many gates for the backward kernel.");
  QDC_TRY(rb_fill_groups(plan_, t, tpos, p.geo.T, insts_, true, &p, &order, &code_of, &swap_of));
  TileSlots h_slots;
  for (int k = 0; k < t.count; k++) {
    const int src = order[k];
    const qdc::Step& st = plan_.tile_steps[t.first + src];
    const int kind = insts_[st.inst].kind;
    TileGateB& G = p.g[k];
    G.type = code_of[src];
    G.a = G.b = 0;
    if (kind_is_diag(kind)) {
      rb_diag_entries(gp[st.inst], swap_of[src], true, &G.inv);
      rb_diag_entries(gp[st.inst], swap_of[src], false, &G.tr);
      diag_hilo_[st.inst] = swap_of[src];
    } else {
      QDC_TRY(tile_matrix(gp[st.inst], kind, kind_is_nonu(kind) ? FORM_INV : FORM_CONJ_TR, swap_of[src], &G.inv));
      QDC_TRY(tile_matrix(gp[st.inst], kind, FORM_TR, swap_of[src], &G.tr));
    }
    G.slot = (int)vslot[st.inst];
    h_slots.s[k] = G.slot;
  }
  const size_t tile_bytes = sizeof(cplx_t) << p.geo.T;
  const size_t smem = 2 * tile_bytes + QDC_TILE_MAXG_B * 32 * sizeof(double) +
                      (QDC_TILE_NT_BRB / 32) * QDC_TILE_MAXG_B * 32 * sizeof(real_t);
  int grid = 0;
  QDC_TRY(tile_grid((const void*)k_tile_bwd_rb, QDC_TILE_NT_BRB, smem, p.geo.ntiles, &grid));
  const size_t need = (size_t)grid * QDC_TILE_MAXG_B * 32;
  if (need > tile_partials_cap_) {
    if (tile_partials_) QDC_CUDA(cudaFree(tile_partials_));
    QDC_CUDA(cudaMalloc((void**)&tile_partials_, need * sizeof(double)));
    tile_partials_cap_ = need;
  }
  cudaEvent_t pa = nullptr;
  if (prof_.on) pa = prof_.begin(stream_);
  k_tile_bwd_rb<<<grid, QDC_TILE_NT_BRB, smem, stream_>>>(state_, bwd_, p, tile_partials_);
  QDC_CUDA(cudaGetLastError());
  k_tile_final<<<t.count, 32, 0, stream_>>>(tile_partials_, grid, t.count, h_slots, d_res_);
  QDC_CUDA(cudaGetLastError());
  if (prof_.on) prof_.end(stream_, CAT_TILE_BWD, pa, 4ull * t.count * bytes());
  stats_.kernel_launches += 2;
  stats_.hbm_passes += 2;
  stats_.algorithmic_bytes += 4ull * t.count * bytes();
  return nullptr;
}
