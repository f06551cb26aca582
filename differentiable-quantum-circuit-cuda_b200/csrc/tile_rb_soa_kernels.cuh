// Register-blocked forward pass in the pair-lane layout (f32).
//
// Combines the two levers measured separately before:
//  * register blocking (tile_rb_kernels.cuh): the gates of a pass are grouped by <= 4 tile
//    positions and a thread applies a whole group (~3 gates for brickwork) to a block it holds in
//    registers, so shared memory is read and written once per GROUP -- the per-gate forward kernel
//    needs 8 x 512 B of shared traffic per 64 FFMA2, i.e. 93 % of the 128 B/clk/SM shared bandwidth
//    at full FP32 rate;
//  * the pair-lane layout (tile_soa_kernels.cuh): FFMA2 with uniform-register gate entries and no
//    operand preparation.
// A block is the 2^4 settings of four tile positions OTHER than position 0, each element being one
// float4 = the two amplitudes of tile bit 0 as lanes: 16 vectors (32 amplitudes, 64 registers) per
// thread, 2^(T-5) = 128 blocks per 2^12 tile = one block per thread and group.  Gates on block
// positions run on both lanes at once; a gate that involves tile position 0 mixes the lanes and goes
// through the complex helpers on unpacked (re, im) pairs.
#pragma once
#include "tile_rb_kernels.cuh"
#include "tile_soa_kernels.cuh"

#ifndef QDC_F64

#define QDC_RBS_NT 128  // threads per CTA

// gate codes: 0..5 dense q2 on block positions (1,0),(2,0),(2,1),(3,0),(3,1),(3,2); 6..9 dense q1 on
// block position 0..3; 10..15 diagonal on the same six pairs; 16..19 dense q2 on (block position k,
// lane bit); 20 dense q1 on the lane bit; 21..24 diagonal on (block position k, lane bit).
// Matrices / diagonals are in (hi,lo) order; the lane bit is always the lo one.

template <int GA, int GB>
__device__ __forceinline__ void rbs_q2(V4 (&x)[16], const GateMat& G) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const int i0 = rb_ins0(rb_ins0(r, GB), GA);
    const V4 a[4] = {x[i0], x[i0 + (1 << GB)], x[i0 + (1 << GA)], x[i0 + (1 << GA) + (1 << GB)]};
    V4 o[4];
    mv_soa<4>(G, a, o);
    x[i0] = o[0];
    x[i0 + (1 << GB)] = o[1];
    x[i0 + (1 << GA)] = o[2];
    x[i0 + (1 << GA) + (1 << GB)] = o[3];
  }
}

template <int P>
__device__ __forceinline__ void rbs_q1(V4 (&x)[16], const GateMat& G) {
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int i0 = rb_ins0(r, P);
    const V4 a[2] = {x[i0], x[i0 + (1 << P)]};
    V4 o[2];
    mv_soa<2>(G, a, o);
    x[i0] = o[0];
    x[i0 + (1 << P)] = o[1];
  }
}

template <int GA, int GB>
__device__ __forceinline__ void rbs_diag(V4 (&x)[16], const GateMat& D) {
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const int j = 2 * ((k >> GA) & 1) + ((k >> GB) & 1);
    x[k] = cmul_bc(x[k], D.re[j], D.im[j]);
  }
}

// dense q2 on (block position P, lane bit)
template <int P>
__device__ __forceinline__ void rbs_q2_lane(V4 (&x)[16], const GateMat& G) {
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const int i0 = rb_ins0(r, P);
    V4 v[2] = {x[i0], x[i0 + (1 << P)]};
    cplx_t a[4];
    unpack_lane_mix<4>(v, a);
    mv<4>(G, a);
    pack_lane_mix<4>(v, a);
    x[i0] = v[0];
    x[i0 + (1 << P)] = v[1];
  }
}

__device__ __forceinline__ void rbs_q1_lane(V4 (&x)[16], const GateMat& G) {
#pragma unroll
  for (int k = 0; k < 16; k++) {
    V4 v[1] = {x[k]};
    cplx_t a[2];
    unpack_lane_mix<2>(v, a);
    mv<2>(G, a);
    pack_lane_mix<2>(v, a);
    x[k] = v[0];
  }
}

// diagonal on (block position P, lane bit): entry 2 bit_P + lane
template <int P>
__device__ __forceinline__ void rbs_diag_lane(V4 (&x)[16], const GateMat& D) {
  float2 dr[2], di[2];
#pragma unroll
  for (int h = 0; h < 2; h++) {
    dr[h] = make_float2(D.re[2 * h], D.re[2 * h + 1]);
    di[h] = make_float2(D.im[2 * h], D.im[2 * h + 1]);
  }
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = cmul_pair(x[k], dr[(k >> P) & 1], di[(k >> P) & 1]);
}

__device__ __forceinline__ void rbs_apply(int code, V4 (&x)[16], const GateMat& G) {
  switch (code) {
    case 0: rbs_q2<1, 0>(x, G); break;
    case 1: rbs_q2<2, 0>(x, G); break;
    case 2: rbs_q2<2, 1>(x, G); break;
    case 3: rbs_q2<3, 0>(x, G); break;
    case 4: rbs_q2<3, 1>(x, G); break;
    case 5: rbs_q2<3, 2>(x, G); break;
    case 6: rbs_q1<0>(x, G); break;
    case 7: rbs_q1<1>(x, G); break;
    case 8: rbs_q1<2>(x, G); break;
    case 9: rbs_q1<3>(x, G); break;
    case 10: rbs_diag<1, 0>(x, G); break;
    case 11: rbs_diag<2, 0>(x, G); break;
    case 12: rbs_diag<2, 1>(x, G); break;
    case 13: rbs_diag<3, 0>(x, G); break;
    case 14: rbs_diag<3, 1>(x, G); break;
    case 15: rbs_diag<3, 2>(x, G); break;
    case 16: rbs_q2_lane<0>(x, G); break;
    case 17: rbs_q2_lane<1>(x, G); break;
    case 18: rbs_q2_lane<2>(x, G); break;
    case 19: rbs_q2_lane<3>(x, G); break;
    case 20: rbs_q1_lane(x, G); break;
    case 21: rbs_diag_lane<0>(x, G); break;
    case 22: rbs_diag_lane<1>(x, G); break;
    case 23: rbs_diag_lane<2>(x, G); break;
    default: rbs_diag_lane<3>(x, G); break;
  }
}

// RbGroup::bit[] holds VECTOR-index bit positions here (tile position - 1), RbGroup::map deposits the
// block number into the remaining vector-index bits.
#ifndef QDC_RBS_MINB
#define QDC_RBS_MINB 3
#endif
__global__ void __launch_bounds__(QDC_RBS_NT, QDC_RBS_MINB)
    k_tile_fwd_rbs(cplx_t* __restrict__ state, const __grid_constant__ TileFwdRbParams p) {
  extern __shared__ __align__(16) unsigned char tile_smem[];
  vec_t* smv = (vec_t*)tile_smem;
  const int nblocks = 1 << (p.geo.T - QDC_LV - QDC_RB);
  TileAddr<QDC_RBS_NT> ta;
  ta.init(p.geo);
  for (uint64_t tile = blockIdx.x; tile < p.geo.ntiles; tile += gridDim.x) {
    const uint64_t tbase = p.geo.tile(tile) >> QDC_LV;
    tile_io_soa<QDC_RBS_NT, true>((vec_t*)state, smv, ta, tbase);
    __syncthreads();
    for (int gi = 0; gi < p.ngroups; gi++) {
      const RbGroup& G = p.grp[gi];
      uint32_t off[16];
      rb_offsets(G, off);
      for (int j0 = 0; j0 < nblocks; j0 += QDC_RBS_NT) {
        // A 2^11 tile has 64 blocks for 128 threads: the upper half of the CTA recomputes blocks of the
        // lower half and only its STORES are predicated off -- the loop stays convergent, which is what
        // lets ptxas keep the gate entries in uniform registers (a `break` here turned every LDCU into
        // a per-thread LDC).
        const bool active = j0 + (int)threadIdx.x < nblocks;
        const uint32_t base = (uint32_t)G.map((uint64_t)((j0 + threadIdx.x) & (nblocks - 1)));
        V4 x[16];
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = ld4(smv + base + off[k]);
        for (int q = 0; q < G.count; q++) {
          const TileGateF& M = p.g[G.first + q];
          rbs_apply(M.type, x, M.m);
        }
        if (nblocks < QDC_RBS_NT) __syncthreads();  // (uniform) the duplicates have read before anyone writes
#pragma unroll
        for (int k = 0; k < 16; k++)
          if (active) st4(smv + base + off[k], x[k]);
      }
      __syncthreads();
    }
    tile_io_soa<QDC_RBS_NT, false>((vec_t*)state, smv, ta, tbase);
    __syncthreads();
  }
}

// ------------------------------------------------------------ host side
// Groups of a pass for the pair-lane register-blocked kernel: block positions are the group's tile
// positions other than 0 (in vector-index space), padded to 4 with the highest unused ones.
static inline const char* rbs_make_groups(const qdc::Plan& plan, const qdc::Step& t, const std::vector<int>& tpos,
                                          int T, const std::vector<Inst>& insts, RbGroup* grp, int* ngroups,
                                          std::vector<int>* code_of, std::vector<bool>* swap_of) {
  if (t.grp_count > QDC_RB_MAXGRP) return qdc_errf("tile pass holds too many register-block groups.");
  if (T - 1 < QDC_RB) return qdc_errf("tile too small for register blocks.");
  *ngroups = t.grp_count;
  code_of->assign(t.count, -1);
  swap_of->assign(t.count, false);
  for (int gi = 0; gi < t.grp_count; gi++) {
    const qdc::Group& g = plan.groups[t.grp_first + gi];
    RbGroup& G = grp[gi];
    std::vector<int> bits;  // vector-index bit positions (tile position - 1)
    for (int k = 0; k < g.nbits; k++)
      if (tpos[g.bits[k]] > 0) bits.push_back(tpos[g.bits[k]] - 1);
    for (int pos = T - 2; (int)bits.size() < QDC_RB && pos >= 0; pos--)
      if (std::find(bits.begin(), bits.end(), pos) == bits.end()) bits.push_back(pos);
    std::sort(bits.begin(), bits.end());
    for (int k = 0; k < QDC_RB; k++) G.bit[k] = bits[k];
    std::vector<int> rest;
    for (int pos = 0; pos < T - 1; pos++)
      if (std::find(bits.begin(), bits.end(), pos) == bits.end()) rest.push_back(pos);
    if (!make_deposit(rest, &G.map)) return qdc_errf("register-block map too fragmented.");
    G.first = g.first - t.first;
    G.count = g.count;
    auto local = [&](int phys) {  // block-local index of a physical position; -1 for the lane bit
      const int tp = tpos[phys];
      if (tp == 0) return -1;
      for (int i = 0; i < QDC_RB; i++)
        if (G.bit[i] == tp - 1) return i;
      return -2;
    };
    for (int q = 0; q < g.count; q++) {
      const int k = g.first - t.first + q;
      const qdc::Step& st = plan.tile_steps[g.first + q];
      const int kind = insts[st.inst].kind;
      if (kind_is_q1(kind)) {
        const int a = local(st.p2);
        if (a == -2) return qdc_errf("register-block group does not hold its gate.");
        (*code_of)[k] = a < 0 ? 20 : 6 + a;
      } else {
        const bool swap = st.p2 < st.p1;  // pos2 on the lower position: present the matrix in (hi,lo) order
        const int a = local(swap ? st.p1 : st.p2), b = local(swap ? st.p2 : st.p1);
        if (a < 0 || b == -2) return qdc_errf("register-block group does not hold its gate.");
        (*swap_of)[k] = swap;
        if (b < 0) (*code_of)[k] = (kind_is_diag(kind) ? 21 : 16) + a;
        else (*code_of)[k] = (kind_is_diag(kind) ? 10 : 0) + rb_pair_code(a, b);
      }
    }
  }
  return nullptr;
}

inline const char* Circuit::run_tile_forward_rbs(const qdc::Step& t, const std::vector<const cplx_t*>& gp,
                                                 bool uncompute) {
  static thread_local TileFwdRbParams p;
  static thread_local RbGroup tmp[QDC_RB_MAXGRP];
  std::vector<int> tpos, code_of;
  std::vector<bool> swap_of;
  QDC_TRY(make_tile_geo(plan_, t, n_loc_, 0, &p.geo, &tpos));
  if (t.count > QDC_TILE_MAXG_F) return qdc_errf("tile pass holds too many gates.");
  int ng = 0;
  QDC_TRY(rbs_make_groups(plan_, t, tpos, p.geo.T, insts_, tmp, &ng, &code_of, &swap_of));
  p.ngroups = ng;
  p.ngates = t.count;
  // forward: plan order; un-compute: groups and gates reversed, inverse matrices
  int next = 0;
  for (int gi = 0; gi < ng; gi++) {
    const RbGroup& src = tmp[uncompute ? ng - 1 - gi : gi];
    RbGroup& dst = p.grp[gi];
    dst = src;
    dst.first = next;
    for (int q = 0; q < src.count; q++) {
      const int s = uncompute ? src.first + src.count - 1 - q : src.first + q;
      const qdc::Step& st = plan_.tile_steps[t.first + s];
      const int kind = insts_[st.inst].kind;
      TileGateF& G = p.g[next + q];
      G.type = code_of[s];
      G.a = G.b = G.pad = 0;
      const int form = uncompute ? (kind_is_nonu(kind) ? FORM_INV : FORM_CONJ_TR) : FORM_PLAIN;
      QDC_TRY(tile_matrix(gp[st.inst], kind, form, swap_of[s], &G.m));
    }
    next += src.count;
  }
  const size_t smem = sizeof(cplx_t) << p.geo.T;
  int grid = 0;
  QDC_TRY(tile_grid((const void*)k_tile_fwd_rbs, QDC_RBS_NT, smem, p.geo.ntiles, &grid));
  cudaEvent_t pa = nullptr;
  if (prof_.on) pa = prof_.begin(stream_);
  k_tile_fwd_rbs<<<grid, QDC_RBS_NT, smem, stream_>>>(state_, p);
  QDC_CUDA(cudaGetLastError());
  if (prof_.on) prof_.end(stream_, uncompute ? CAT_UNCOMPUTE : CAT_TILE_FWD, pa, 2ull * t.count * bytes());
  stats_.kernel_launches += 1;
  stats_.hbm_passes += 1;
  stats_.algorithmic_bytes += 2ull * t.count * bytes();
  return nullptr;
}

#endif  // !QDC_F64
