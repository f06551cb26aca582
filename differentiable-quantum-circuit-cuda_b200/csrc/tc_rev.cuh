// Fused reverse step of a tensor-core block (f32 build): ONE sweep over state and adjoint (4*S of HBM traffic) does
// what tc_exec.cuh used to do in three (6*S):
//
//     state   <- W^dagger state                       (un-compute)
//     adjoint <- W^T adjoint = conj(W^dagger conj(adjoint))     (pull-back; the SAME W^dagger image serves both)
//     H       += adjoint (x) state                    (both as loaded, i.e. AFTER the block)
//
// and the block gradient of the reference's convention (src/primitives.cu:323-329), G_W = sum adjoint (x) (W^dagger
// state), follows on the host as G_W = H conj(W) (tc_host.hpp: tc_chain_rule_from_h) -- 64 x 64 matrices, in double.
// H needs only the slices that the two block products need anyway, so one fill of shared memory feeds three GEMMs:
//
//     X' [128 x 64]  = R(W^dagger) [128 x 128] * X~ [128 x 64]           A = W slices (tensor memory), B = X slices
//     Y'~[128 x 64]  = R(W^dagger) [128 x 128] * conj(Y)~ [128 x 64]     B = conj(Y) slices (imaginary rows negated)
//     P' [128 x 128] += conj(Y)~ [128 x 64] * X~^T [64 x 128]            A, B = the same slices read along their rows
//
// Exact 9-bit bf16 slices and the (A0 exact | A1 lower order) accumulator pairs are those of tc_block.cuh.
//
// Tensor memory (512 columns): W slices (3 x 64) | H (128) | block accumulator: A0, A1 (2 x 64).
// H is ONE accumulator for all its slice products: its truncation bias (~1e-8 per accumulation, a window of kFlush
// tiles = 96 accumulations) is a relative error of < 1e-6 of a gradient entry, once -- unlike the state and the adjoint,
// which pass through ~70 blocks and keep the exact (A0 | A1) pair.  That leaves room for all three W slices in tensor
// memory: every MMA operand read from shared memory shares the L1 data pipe with the fill / drain traffic, which is
// what bounds this kernel (profiles/r2_tc_rev_28q_v2_ncu.txt: tensor operand wavefronts 41 % + LSU 37 % of the pipe).
// Shared memory: 2 stages x (X slices 48 KiB + conj(Y) slices 48 KiB) + 32 KiB of output staging = 224 KiB.  The
// staging buffer is the drain's own, so a stage goes back to the fill the moment its last MMA has completed
// (tcgen05.commit arrives on the barriers the fill waits on): the stage cycle is fill + MMA, not fill + MMA + drain.
//
// Warp roles (416 threads, one CTA per SM, persistent over tiles):
//   warps 0-7  fill  : HBM -> registers (the next tile's loads in flight) -> running maximum exponent per warp ->
//                      slices of X and conj(Y) -> `full`
//   warp  8    MMA   : one elected lane; per tile  H_a | X' -> A | H_b -> B | Y' -> C   (H split in two so that the
//                      drain's tcgen05.ld of the single block accumulator pair always has tensor work to hide behind)
//   warps 9-12 drain : A: X' accumulators -> registers (`acc_empty`) -> staging -> registers -> state;
//                      C: Y' likewise, imaginary part negated -> adjoint.
//                      Every kFlush tiles: H accumulator -> red.global.add into the CTA's private partial.
#pragma once
#include "tc_block.cuh"

namespace tcb {

constexpr int kRevStageBytes = 2 * kStageBytes;                       // 96 KiB
constexpr int kRevStages = 2;
constexpr int kRevOutOff = kRevStages * kRevStageBytes;               // 192 KiB: output staging, 2 x 16 KiB
constexpr int kRevBarOff = kRevOutOff + 2 * kSliceBytesX;             // 224 KiB
constexpr int kRevSmemBytes = kRevBarOff + 1024 /*alignment slack*/ + 512 /*barriers*/;
constexpr int kRevThreads = kFillThreads + 32 + kDrainThreads;        // 416
#ifndef TC_REV_PREFETCH
#define TC_REV_PREFETCH 0
#endif
constexpr int kRevPrefetch = TC_REV_PREFETCH;
constexpr uint32_t kRevTmH = 192, kRevTmA0 = 320, kRevTmA1 = 384;

struct RevParams {
  Params geo;          // w_image = image of W^dagger (make_w_image); products: 8 or 6
  float* partials;     // [gridDim.x][128][128], zero on entry; row = real-ified conj(adjoint) index, column = state index
  int h_products;      // slice products of H: 6 (orders 0..2, 2^-27 per term).  3 (orders 0, 1) is a microbenchmark switch
                       // only: 2^-18 per term is 6e-5 of the largest entry of an incoherent sum (r2_tc_rev_bench_v6.txt)
#ifdef TC_REV_TRACE
  long long* trace;    // [8 tiles][32 slots] clock64 stamps of CTA 0, tiles 8..15 (profiles/microbench/tc_rev_bench.cu)
#endif
};

#ifdef TC_REV_TRACE
#define TC_TR(itv, slot)                                                                      \
  do {                                                                                        \
    if (blockIdx.x == 0 && (itv) >= 8 && (itv) < 16) rp.trace[((itv) - 8) * 32 + (slot)] = clock64(); \
  } while (0)
#else
#define TC_TR(itv, slot) do { } while (0)
#endif

// 32 lanes x 16 consecutive 32-bit columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// slices of the two items of a fill thread into one 48 KiB slice set; NEG_IM: the slices of the complex conjugate
// (the slicing is odd-symmetric, so this is an exact sign flip of the imaginary rows)
template <bool NEG_IM>
__device__ __forceinline__ void rev_fill_slices(const float4 (&v)[2][4], const uint32_t (&soff)[2], float m0, float m1, float m2,
                                                uint8_t* slices) {
#pragma unroll
  for (int it = 0; it < 2; it++) {
#pragma unroll
    for (int c = 0; c < 2; c++) {
      float q0[8], q1[8], q2[8];
#pragma unroll
      for (int h = 0; h < 4; h++) {
        const float e0 = c ? (NEG_IM ? -v[it][h].y : v[it][h].y) : v[it][h].x;
        const float e1 = c ? (NEG_IM ? -v[it][h].w : v[it][h].w) : v[it][h].z;
        slice3_pair(e0, e1, m0, m1, m2, q0[2 * h], q1[2 * h], q2[2 * h], q0[2 * h + 1], q1[2 * h + 1], q2[2 * h + 1]);
      }
      const uint32_t off = soff[it] + (uint32_t)c * 8192u;   // row (c, j) = row j + 64
      *(uint4*)(slices + 0 * kSliceBytesX + off) =
          make_uint4(pack_hi16(q0[0], q0[1]), pack_hi16(q0[2], q0[3]), pack_hi16(q0[4], q0[5]), pack_hi16(q0[6], q0[7]));
      *(uint4*)(slices + 1 * kSliceBytesX + off) =
          make_uint4(pack_hi16(q1[0], q1[1]), pack_hi16(q1[2], q1[3]), pack_hi16(q1[4], q1[5]), pack_hi16(q1[6], q1[7]));
      *(uint4*)(slices + 2 * kSliceBytesX + off) =
          make_uint4(pack_hi16(q2[0], q2[1]), pack_hi16(q2[2], q2[3]), pack_hi16(q2[4], q2[5]), pack_hi16(q2[6], q2[7]));
    }
  }
}

__global__ void __launch_bounds__(kRevThreads, 1)
    k_tc_block_rev(float2* __restrict__ state, float2* __restrict__ adj, const __grid_constant__ RevParams rp) {
  const Params& p = rp.geo;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + kRevBarOff);
  uint64_t* full = bars;             // [2] fill -> MMA           (256 arrivals)
  uint8_t* sm_out = smem + kRevOutOff;
  uint64_t* bar_a = bars + 4;        // [2] MMA -> drain: X' accumulators complete
  uint64_t* bar_b = bars + 6;        // [2] MMA -> fill: every MMA reading the X slices of the stage has completed
  uint64_t* bar_c = bars + 8;        // [2] MMA -> drain, fill: Y' accumulators complete, the stage is no longer read
  uint64_t* acc_empty = bars + 10;   // drain -> MMA: block accumulators are in registers (twice per tile)
  uint64_t* h_done = bars + 11;      // MMA -> drain: the H window is complete
  uint64_t* h_empty = bars + 12;     // drain -> MMA: H accumulators have been flushed
  uint32_t* tmem_slot = (uint32_t*)(bars + 13);
  uint32_t* e_shared = tmem_slot + 2;        // [2] maximum exponents of the CTA's first state / adjoint tile (fill_grid_seed)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    e_shared[0] = e_shared[1] = 0;
    for (int s = 0; s < 2; s++) {
      mbar_init(&full[s], kFillThreads);
      mbar_init(&bar_a[s], 1);
      mbar_init(&bar_b[s], 1);
      mbar_init(&bar_c[s], 1);
    }
    mbar_init(acc_empty, kDrainThreads);
    mbar_init(h_done, 1);
    mbar_init(h_empty, kDrainThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kFillWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp > kFillWarps) {
    // W slices -> tensor memory: lane = row mu, column = pair of bf16 along K (64 columns per slice)
    const int q4 = warp & 3, mu = q4 * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
#pragma unroll 1
    for (int sl = 0; sl < kSlices; sl++) {
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const uint4* src = (const uint4*)(p.w_image + ((size_t)sl * kDim + mu) * 64 + half * 32);
        uint32_t r[32];
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const uint4 t = __ldg(src + k);
          r[4 * k] = t.x; r[4 * k + 1] = t.y; r[4 * k + 2] = t.z; r[4 * k + 3] = t.w;
        }
        tmem_st32(lane_addr + sl * 64 + half * 32, r);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint64_t my_tiles = p.ntiles > (uint64_t)blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp < kFillWarps) {
    // =========================================================================== fill
    const int t = threadIdx.x;
    ItemAddr<2, kFillThreads> ia;
    ia.init(p, t);
    auto load = [&](const float2* base, uint64_t tbase, float4 (&v)[2][4]) {
      const float2* src = base + tbase;
#pragma unroll
      for (int it = 0; it < 2; it++) load_item(src + ia.goff[it], p, v[it]);
    };
    // L2 prefetch: threads 0..127 own one 256-byte run of the state tile each, threads 128..255 of the adjoint tile
    const uint64_t roff = run_offset(p, t & 127);
    const float2* pf_base = (t < 128 ? state : adj) + roff;
    auto prefetch = [&](uint64_t tbase) { prefetch_run_l2(pf_base + tbase); };
    TileWalk wl, wp;   // base of the tile whose loads are issued next / of the tile prefetched into L2 next
    wl.init(p, blockIdx.x, gridDim.x);
    wp = wl;
    // (optional L2 prefetch kRevPrefetch tiles ahead, one bulk instruction per thread and tile; off by default, see kPrefetch)
    for (int k = 0; k < kRevPrefetch; k++) {
      if (blockIdx.x + (uint64_t)k * gridDim.x < p.ntiles) prefetch(wp.cur);
      wp.advance();
    }
    uint32_t e_run_x = 0, e_run_y = 0;
    uint32_t it_count = 0;
    // Software pipeline over registers: the loads of the NEXT tile are issued as soon as the slices of this tile's
    // state (adjoint) have been taken, so that a whole tile pair (64 KiB, ~3000 cycles at this SM's share of the HBM
    // bandwidth) is in flight while the fill computes, fences and waits for its stage.
    float4 vx[2][4], vy[2][4];
    if ((uint64_t)blockIdx.x < p.ntiles) {
      load(state, wl.cur, vx);
      load(adj, wl.cur, vy);
      e_run_x = fill_grid_seed(vx, e_shared, lane);
      e_run_y = fill_grid_seed(vy, e_shared + 1, lane);
    }
    for (uint64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it_count++) {
      const int s = it_count & 1;
      const uint32_t use = it_count >> 1;
      uint8_t* stage = smem + s * kRevStageBytes;
      const uint64_t next = tile + gridDim.x;
      if (t == 0) TC_TR(it_count, 0);
      e_run_x = max(e_run_x, warp_max_exp(vx));
      float m0, m1, m2;
      magic_of(e_run_x, m0, m1, m2);
      if (t == 0) TC_TR(it_count, 1);
      if (use > 0) mbar_wait(&bar_b[s], (use - 1) & 1, p.error_flag);   // X slices of tile t - 2 are no longer read
      if (t == 0) TC_TR(it_count, 2);
      rev_fill_slices<false>(vx, ia.soff, m0, m1, m2, stage);
      wl.advance();
      if (next < p.ntiles) load(state, wl.cur, vx);
      if (t == 0) TC_TR(it_count, 3);
      e_run_y = max(e_run_y, warp_max_exp(vy));
      magic_of(e_run_y, m0, m1, m2);
      if (use > 0) mbar_wait(&bar_c[s], (use - 1) & 1, p.error_flag);   // nor its conj(Y) slices
      if (t == 0) TC_TR(it_count, 4);
      rev_fill_slices<true>(vy, ia.soff, m0, m1, m2, stage + kStageBytes);
      if (next < p.ntiles) load(adj, wl.cur, vy);
      if (kRevPrefetch > 0) {
        if (tile + (uint64_t)kRevPrefetch * gridDim.x < p.ntiles) prefetch(wp.cur);
        wp.advance();
      }
      if (t == 0) TC_TR(it_count, 5);
      // (no fence.proxy.async here: it waits for the loads of the next tile that this thread has in flight -- up to
      // 1800 cycles, profiles/r2_tc_rev_bench_v4.txt.  The MMA warp executes the proxy fence after it has acquired
      // `full`: arrive (release) -> wait (acquire) -> fence.proxy.async -> tcgen05.mma keeps these stores in the
      // causality order of the asynchronous reads.)
      mbar_arrive(&full[s]);
      if (t == 0) TC_TR(it_count, 6);
    }
  } else if (warp == kFillWarps) {
    // =========================================================================== MMA issue
    // The whole warp runs the loop converged and ONE elected lane issues (elect.sync): ptxas then knows that the
    // tcgen05 instructions have a single active thread.  Issued from `if (lane == 0)` every UTCHMMA sat inside a
    // compiler-generated leader-election loop, 50-80 cycles per instruction -- more than the 32 cycles the tensor pipe
    // needs for it (profiles/r2_tc_rev_trace_v1.txt).
    constexpr uint32_t idesc_blk = make_idesc(kDim, kN);   // A K-major, B MN-major, M 128, N 64
    // both operands K-major, M = N = 128
    constexpr uint32_t idesc_h = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t base = smem_u32(smem);
    const uint32_t acc0 = tmem_base + kRevTmA0, acc1 = tmem_base + kRevTmA1;
    const uint32_t hacc = tmem_base + kRevTmH;
    const bool all8 = p.products >= 8, h6 = rp.h_products >= 6;
    // R(W^dagger) * (slices whose MN-major descriptor starts at b_lo): (0,0) alone into A0 (exact), the lower-order
    // products into A1
    auto block_products = [&](uint32_t b_lo) {
      bool first1 = true;
#pragma unroll
      for (int pw = 0; pw < 3; pw++) {
#pragma unroll
        for (int px = 0; px < 3; px++) {
          if (pw + px >= 4) continue;
          if (pw + px == 3 && !all8) continue;
          const bool lead = (pw == 0 && px == 0);
#pragma unroll
          for (int ks = 0; ks < 8; ks++) {   // K = 16 per instruction = 8 TMEM columns of A, 16 rows of B
            const uint64_t bd = desc_at(b_lo, kDescHi, px * kSliceBytesX + ks * 2048);
            const uint32_t accumulate = lead ? (ks > 0) : !(first1 && ks == 0);
            umma_bf16_ts(lead ? acc0 : acc1, tmem_base + (uint32_t)(pw * 64 + ks * 8), bd, idesc_blk, accumulate);
          }
          if (!lead) first1 = false;
        }
      }
    };
    // conj(Y) slice pb (rows = M) times X slice pa (rows = N), K = the 64 rest columns of the tile
    auto h_product = [&](uint32_t y_lo, uint32_t x_lo, int pb, int pa, uint32_t acc, bool fresh) {
#pragma unroll
      for (int ks = 0; ks < 4; ks++)   // K = 16 columns n per instruction: 32 bytes along the 128-byte rows
        umma_bf16(acc, desc_at(y_lo, kDescHi, pb * kSliceBytesX + ks * 32), desc_at(x_lo, kDescHi, pa * kSliceBytesX + ks * 32), idesc_h,
                  !(fresh && ks == 0));
    };
    for (uint32_t it = 0; it < (uint32_t)my_tiles; it++) {
      const int s = it & 1;
      const uint32_t use = it >> 1, win = it / kFlush;
      const bool first = (it % kFlush) == 0, last = (it % kFlush) == kFlush - 1 || it + 1 == (uint32_t)my_tiles;
      mbar_wait(&full[s], use & 1, p.error_flag);
      fence_async_smem();   // generic-proxy stores of the fill warps (acquired through `full`) -> async-proxy reads of the MMAs
      if (first && win >= 1) mbar_wait(h_empty, (win - 1) & 1, p.error_flag);
      tc_fence_after();
      const uint32_t xs = base + s * kRevStageBytes, ys = xs + kStageBytes;
      const uint32_t xk_lo = desc_lo(xs, 16), yk_lo = desc_lo(ys, 16);          // K-major views (H)
      const uint32_t xm_lo = desc_lo(xs, 1024), ym_lo = desc_lo(ys, 1024);      // MN-major views (block products)
      if (elect_one()) {
        TC_TR(it, 8);
        // H_a: orders 0 and 1
        h_product(yk_lo, xk_lo, 0, 0, hacc, first);
        h_product(yk_lo, xk_lo, 0, 1, hacc, false);
        h_product(yk_lo, xk_lo, 1, 0, hacc, false);
        TC_TR(it, 9);
      }
      __syncwarp();
      if (it >= 1) {   // Y' of the previous tile has left the block accumulators
        mbar_wait(acc_empty, 1, p.error_flag);
        tc_fence_after();
      }
      if (elect_one()) {
        TC_TR(it, 10);
        block_products(xm_lo);
        umma_commit(&bar_a[s]);
        TC_TR(it, 11);
        // H_b: order 2
        if (h6) {
          h_product(yk_lo, xk_lo, 0, 2, hacc, false);
          h_product(yk_lo, xk_lo, 1, 1, hacc, false);
          h_product(yk_lo, xk_lo, 2, 0, hacc, false);
        }
        umma_commit(&bar_b[s]);
        if (last) umma_commit(h_done);
        TC_TR(it, 12);
      }
      __syncwarp();
      mbar_wait(acc_empty, 0, p.error_flag);   // X' of this tile has left the block accumulators
      tc_fence_after();
      if (elect_one()) {
        TC_TR(it, 13);
        block_products(ym_lo);
        umma_commit(&bar_c[s]);
        TC_TR(it, 14);
      }
      __syncwarp();
    }
  } else {
    // =========================================================================== drain
    const int t128 = threadIdx.x - (kFillThreads + 32);
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may access
    const int mu = q4 * 32 + lane;           // output row (c', i)
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
    float* mine = rp.partials + ((size_t)blockIdx.x * kDim + mu) * kDim;
    ItemAddr<4, kDrainThreads> ia;
    ia.init(p, t128);
    // block accumulators A0 + A1 -> d (row mu, n = 0..63)
    auto load_acc = [&](float (&d)[64]) {
#pragma unroll
      for (int q = 0; q < 4; q++) {
        float a0[16], a1[16];
        tmem_ld16(lane_addr + kRevTmA0 + q * 16, a0);
        tmem_ld16(lane_addr + kRevTmA1 + q * 16, a1);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; i++) d[q * 16 + i] = a0[i] + a1[i];
      }
    };
    // d -> staging (two half-buffers with the row / chunk structure of an X slice, see k_tc_block_fwd)
    auto stage_out = [&](uint8_t* region, const float (&d)[64]) {
#pragma unroll
      for (int g = 0; g < 16; g++)   // n = 4 g .. 4 g + 3: half-buffer (g & 1), chunk g / 2
        *(float4*)(region + (g & 1) * kSliceBytesX + x_chunk_byte(mu, g >> 1)) = make_float4(d[4 * g], d[4 * g + 1], d[4 * g + 2], d[4 * g + 3]);
    };
    // staging -> registers (the stage's region is free again once every drain thread has done this), registers -> HBM
    auto read_staged = [&](const uint8_t* region, float4 (&r)[4][4]) {
#pragma unroll
      for (int it = 0; it < 4; it++) {
        r[it][0] = *(const float4*)(region + ia.soff[it]);
        r[it][1] = *(const float4*)(region + kSliceBytesX + ia.soff[it]);
        r[it][2] = *(const float4*)(region + ia.soff[it] + 8192u);
        r[it][3] = *(const float4*)(region + kSliceBytesX + ia.soff[it] + 8192u);
      }
    };
    auto write_out = [&](float2* dst, const float4 (&r)[4][4], float sign_im) {
#pragma unroll
      for (int it = 0; it < 4; it++) {
        const float4 r0 = r[it][0], r1 = r[it][1];
        float4 i0 = r[it][2], i1 = r[it][3];
        i0.x *= sign_im; i0.y *= sign_im; i0.z *= sign_im; i0.w *= sign_im;
        i1.x *= sign_im; i1.y *= sign_im; i1.z *= sign_im; i1.w *= sign_im;
        const float4 o[4] = {make_float4(r0.x, i0.x, r0.y, i0.y), make_float4(r0.z, i0.z, r0.w, i0.w),
                             make_float4(r1.x, i1.x, r1.y, i1.y), make_float4(r1.z, i1.z, r1.w, i1.w)};
        store_item(dst + ia.goff[it], p, o);
      }
    };
    TileWalk wd;
    wd.init(p, blockIdx.x, gridDim.x);
    uint32_t it_count = 0;
    for (uint64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it_count++, wd.advance()) {
      const int s = it_count & 1;
      const uint32_t use = it_count >> 1, win = it_count / kFlush;
      const bool last = (it_count % kFlush) == kFlush - 1 || it_count + 1 == (uint32_t)my_tiles;
      const uint64_t tbase = wd.cur;
      float d[64];
      float4 r[4][4];
      // ---- X'
      mbar_wait(&bar_a[s], use & 1, p.error_flag);
      tc_fence_after();
      if (t128 == 0) TC_TR(it_count, 16);
      load_acc(d);
      tc_fence_before();
      mbar_arrive(acc_empty);
      if (t128 == 0) TC_TR(it_count, 17);
      stage_out(sm_out, d);
      named_bar(2, kDrainThreads);
      read_staged(sm_out, r);
      named_bar(2, kDrainThreads);                // every drain thread has read its staging rows
      if (t128 == 0) TC_TR(it_count, 18);
      if (last) {
        // H window -> the CTA's private partial (single writer per element; red = no round trip)
        mbar_wait(h_done, win & 1, p.error_flag);
        tc_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < 4; ch++) {
          float a0[32];
          tmem_ld32(lane_addr + kRevTmH + ch * 32, a0);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; g++) red_add_v4(mine + ch * 32 + 4 * g, a0[4 * g], a0[4 * g + 1], a0[4 * g + 2], a0[4 * g + 3]);
        }
        tc_fence_before();
        mbar_arrive(h_empty);
      }
      if (t128 == 0) TC_TR(it_count, 19);
      write_out(state + tbase, r, 1.0f);
      if (t128 == 0) TC_TR(it_count, 21);
      // ---- Y'
      mbar_wait(&bar_c[s], use & 1, p.error_flag);
      tc_fence_after();
      if (t128 == 0) TC_TR(it_count, 22);
      load_acc(d);
      tc_fence_before();
      mbar_arrive(acc_empty);
      if (t128 == 0) TC_TR(it_count, 23);
      stage_out(sm_out, d);
      named_bar(2, kDrainThreads);
      read_staged(sm_out, r);
      named_bar(2, kDrainThreads);
      if (t128 == 0) TC_TR(it_count, 25);
      write_out(adj + tbase, r, -1.0f);
      if (t128 == 0) TC_TR(it_count, 26);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kFillWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

}  // namespace tcb
