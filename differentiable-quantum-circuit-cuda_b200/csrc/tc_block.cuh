// Fused 6-qubit blocks on the 5th-generation tensor cores (tcgen05 / TMEM), f32 build.
//
// A window of gates acting inside 6 qubits is multiplied on the host into one dense 64 x 64 complex
// matrix W; the pass then applies W to every group of 64 amplitudes as a real GEMM
//
//     D[mu, n] = sum_kappa R(W)[mu, kappa] * X[kappa, n],      mu, kappa = (re/im, block index) in 0..127
//
// with M = 128 (one tcgen05.mma, cta_group::1), K = 128, N = 64 "rest" columns per 2^12-amplitude tile.
//
// Precision.  TF32 / BF16 products alone carry 2^-11 / 2^-8 relative error, and tensor-core FP32
// accumulation truncates (round-toward-zero), which at ~50 accumulations per output is a SYSTEMATIC
// shrink of ~5e-7 per block -- 1e-5 after the ~35 blocks a depth-100 amplitude passes through.  The
// kernel therefore works on exact slices: every f32 value x of a tile is split, on the tile-uniform
// grid g0 = 2^(E-7) (2^E > max |x| of the tile), into three BF16 numbers
//
//     x = p0 + p1 + p2 + r,   p_i = k_i * g0 * 2^(-8 i),  |k_i| <= 128,  |r| <= 2^(E-24)
//
// (three magic-number roundings, 8 FADD per value; the high 16 bits of each f32 slice ARE the bf16).
// W is sliced the same way on the host (grid 2^-7).  The leading products p0(W) * p0(X) are integers
// times g0 * 2^-7 of at most 2^14; their sum over K = 128 stays below 2^24, so the tensor core adds
// them EXACTLY in its own TMEM accumulator A0 whatever its rounding mode.  The seven lower-order
// products (all but p2 * p2) go to a second accumulator A1 whose truncation errors are 2^-8 smaller
// than the result's last bit.  D = A0 + A1 is one rounded f32 addition in the epilogue.  Error per
// block: the 2^-24 * (tile max) quantisation of the inputs, unbiased.
//
// Shared-memory layouts (what the UMMA descriptors describe):
//   * W slice i : A operand, K-major, SWIZZLE_128B, bf16: two K blocks of 64 (128 B rows), 128 rows.
//   * X slice j : B operand, MN-major (n contiguous), SWIZZLE_128B, bf16: row kappa = 128 B = 64 n.
//   * output staging (f32), two half-buffers with the SAME row / chunk structure as an X slice, so that the
//     drain is the mirror image of the fill.
// The fill goes through registers (LDG.128 -> slices -> STS.128): the slicing needs the CUDA cores
// anyway, and the interleaved (re, im) HBM layout of the reference is de-interleaved on the way.
//
// Warp roles (288 threads, one CTA per SM, persistent over tiles, 2-stage pipeline):
//   warps 0-3  fill    : HBM -> registers -> tile max -> slices -> smem stage, arrive `full`
//   warp  4    MMA     : one elected lane issues 64 tcgen05.mma per tile, tcgen05.commit -> `mma_done`
//   warps 5-8  drain   : tcgen05.ld A0, A1 -> D -> smem staging -> HBM, arrive `tmem_empty`, `empty`
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

namespace tcb {

constexpr int kBlockQubits = 6;
constexpr int kDim = 128;          // 2 * 2^6: real-ified block dimension (M and K of the GEMM)
constexpr int kN = 64;             // rest columns per tile
constexpr int kTileBits = 12;      // 2^12 amplitudes per tile
constexpr int kSlices = 3;
constexpr int kSliceBytesW = kDim * kDim * 2;          // 32 KiB per W slice
constexpr int kSliceBytesX = kDim * kN * 2;            // 16 KiB per X slice
constexpr int kStageBytes = kSlices * kSliceBytesX;    // 48 KiB per pipeline stage
constexpr int kStages = 2;
constexpr int kSmemW = kSlices * kSliceBytesW;         // 96 KiB
constexpr int kSmemBytes = kSmemW + kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int kThreads = 288;
constexpr int kTmemCols = 256;     // 2 stages x (A0, A1) x 64 columns

// Software bit deposit: tile number -> amplitude base (same role as TileGeo::tile)
struct Deposit {
  int nseg;
  unsigned char src[8], dst[8], width[8];
  __host__ __device__ __forceinline__ uint64_t operator()(uint64_t x) const {
    uint64_t out = 0;
    for (int k = 0; k < nseg; k++) out |= ((x >> src[k]) & ((1ull << width[k]) - 1ull)) << dst[k];
    return out;
  }
};

struct Params {
  // tile bit t (ascending physical position) -> physical position, and its role: index bit of the
  // block index j (0..5) or of the rest index n (0..5).  Tile bits 0..2 are physical 0..2 = n0..n2.
  int pos[kTileBits];
  int j_of[kTileBits];   // -1 if the tile bit is a rest bit
  int n_of[kTileBits];   // -1 if the tile bit is a block bit
  Deposit tile;          // tile number -> amplitude base over the other n - 12 positions
  uint64_t ntiles;
  const uint8_t* w_image;  // 96 KiB: the three W slices in their shared-memory byte image (host: make_w_image)
  int* error_flag;         // set to 1 by a watchdog if a barrier wait times out
};

// ------------------------------------------------------------------ byte images (host and device)
// element (m, k) of a W slice (A operand, K-major SW128, bf16)
__host__ __device__ __forceinline__ uint32_t w_byte(int m, int k) {
  const int kb = k >> 6, kk = k & 63;
  return (uint32_t)(kb * 16384 + (m >> 3) * 1024 + (m & 7) * 128 + ((((kk >> 3) ^ (m & 7)) & 7) << 4) + (kk & 7) * 2);
}
// 16-byte chunk holding n = 8 q .. 8 q + 7 of row kappa of an X slice (B operand, MN-major SW128, bf16)
__host__ __device__ __forceinline__ uint32_t x_chunk_byte(int kappa, int q) {
  return (uint32_t)((kappa >> 3) * 1024 + (kappa & 7) * 128 + (((q ^ (kappa & 7)) & 7) << 4));
}

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug must end in a trap (an error), never in a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int* error_flag) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; spin++) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 24)) {
      if (error_flag) atomicExch(error_flag, 1);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout type << 61 (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: D f32, A / B bf16, A K-major, B MN-major
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread = lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// --------------------------------------------------------------------------------- slicing
// x -> three bf16 slices on the grid given by the magic numbers (m0, m0 / 256, m0 / 65536)
__device__ __forceinline__ void slice3(float x, float m0, float m1, float m2, float& p0, float& p1, float& p2) {
  p0 = __fsub_rn(__fadd_rn(x, m0), m0);
  const float r1 = __fsub_rn(x, p0);
  p1 = __fsub_rn(__fadd_rn(r1, m1), m1);
  const float r2 = __fsub_rn(r1, p1);
  p2 = __fsub_rn(__fadd_rn(r2, m2), m2);
}
// the high halves of two f32 words -> one word of two bf16 (lo = first)
__device__ __forceinline__ uint32_t pack_hi16(float a, float b) {
  return __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632);
}

// per-thread addressing of the 4 items (8 consecutive amplitudes each) a fill / drain thread owns
struct ItemAddr {
  uint64_t goff[4];   // amplitude offset inside the tile span
  uint32_t soff[4];   // byte offset of the item's 16-byte chunk in row (c = 0, j) of an X slice
  __device__ __forceinline__ void init(const Params& p, int t128) {
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int item = it * 128 + t128;  // 9 bits <-> tile bits 3..11
      uint64_t g = 0;
      int j = 0, n = 0;
#pragma unroll
      for (int b = 0; b < 9; b++) {
        if ((item >> b) & 1) {
          const int tb = b + 3;
          g |= 1ull << p.pos[tb];
          if (p.j_of[tb] >= 0) j |= 1 << p.j_of[tb]; else n |= 1 << p.n_of[tb];
        }
      }
      goff[it] = g;
      soff[it] = x_chunk_byte(j, n >> 3);
    }
  }
};

// ------------------------------------------------------------------------------------ the kernel
// state: 2^n interleaved (re, im) f32 pairs, updated in place: every group of 64 amplitudes over the block
// qubits is multiplied by W.
__global__ void __launch_bounds__(kThreads, 1) k_tc_block_fwd(float2* __restrict__ state, const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B atoms: 1024-byte aligned
  uint8_t* sm_w = smem;
  uint8_t* sm_x = smem + kSmemW;
  uint64_t* bars = (uint64_t*)(smem + kSmemW + kStages * kStageBytes);
  uint64_t* full = bars;            // [2] fill -> MMA            (128 arrivals)
  uint64_t* empty = bars + 2;       // [2] drain -> fill          (128 arrivals)
  uint64_t* mma_done = bars + 4;    // [2] MMA -> drain           (tcgen05.commit)
  uint64_t* tmem_empty = bars + 6;  // [2] drain -> MMA           (128 arrivals)
  uint32_t* tmem_slot = (uint32_t*)(bars + 8);
  float* sm_max = (float*)(bars + 9);  // [4] warp maxima of the fill group

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup
  for (int i = threadIdx.x; i < kSmemW / 16; i += kThreads)
    ((uint4*)sm_w)[i] = __ldg((const uint4*)p.w_image + i);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; s++) {
      mbar_init(&full[s], 128);
      mbar_init(&empty[s], 128);
      mbar_init(&mma_done[s], 1);
      mbar_init(&tmem_empty[s], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  fence_async_smem();   // the W image was written through the generic proxy, the tensor core reads through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =========================================================================== fill
    const int t128 = threadIdx.x;
    ItemAddr ia;
    ia.init(p, t128);
    uint32_t it_count = 0;
    for (uint64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it_count++) {
      const int s = it_count & 1;
      const uint32_t use = it_count >> 1;
      if (use > 0) mbar_wait(&empty[s], (use - 1) & 1, p.error_flag);
      const float4* src = (const float4*)(state + p.tile(tile));
      float4 v[4][4];
#pragma unroll
      for (int it = 0; it < 4; it++)
#pragma unroll
        for (int h = 0; h < 4; h++) v[it][h] = __ldcs(src + (ia.goff[it] >> 1) + h);
      float mx = 0.f;
#pragma unroll
      for (int it = 0; it < 4; it++)
#pragma unroll
        for (int h = 0; h < 4; h++)
          mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[it][h].x), fabsf(v[it][h].y))), fmaxf(fabsf(v[it][h].z), fabsf(v[it][h].w)));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      if (lane == 0) sm_max[warp] = mx;
      named_bar(1, 128);
      mx = fmaxf(fmaxf(sm_max[0], sm_max[1]), fmaxf(sm_max[2], sm_max[3]));
      named_bar(1, 128);   // sm_max may be overwritten by the next tile only after everyone has read it
      uint32_t bexp = (__float_as_uint(mx) >> 23) & 0xffu;
      bexp = bexp < 24u ? 24u : (bexp > 230u ? 230u : bexp);
      const uint32_t m0b = ((bexp + 17u) << 23) | 0x400000u;   // 1.5 * 2^(E + 16), 2^E > max
      const float m0 = __uint_as_float(m0b), m1 = __uint_as_float(m0b - (8u << 23)), m2 = __uint_as_float(m0b - (16u << 23));
      uint8_t* stage = sm_x + s * kStageBytes;
#pragma unroll
      for (int it = 0; it < 4; it++) {
        float re[8], im[8];
#pragma unroll
        for (int h = 0; h < 4; h++) {
          re[2 * h] = v[it][h].x; im[2 * h] = v[it][h].y;
          re[2 * h + 1] = v[it][h].z; im[2 * h + 1] = v[it][h].w;
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
          float q0[8], q1[8], q2[8];
#pragma unroll
          for (int e = 0; e < 8; e++) slice3(c ? im[e] : re[e], m0, m1, m2, q0[e], q1[e], q2[e]);
          const uint32_t off = ia.soff[it] + (uint32_t)c * 8192u;   // row (c, j) = row j + 64
          *(uint4*)(stage + 0 * kSliceBytesX + off) =
              make_uint4(pack_hi16(q0[0], q0[1]), pack_hi16(q0[2], q0[3]), pack_hi16(q0[4], q0[5]), pack_hi16(q0[6], q0[7]));
          *(uint4*)(stage + 1 * kSliceBytesX + off) =
              make_uint4(pack_hi16(q1[0], q1[1]), pack_hi16(q1[2], q1[3]), pack_hi16(q1[4], q1[5]), pack_hi16(q1[6], q1[7]));
          *(uint4*)(stage + 2 * kSliceBytesX + off) =
              make_uint4(pack_hi16(q2[0], q2[1]), pack_hi16(q2[2], q2[3]), pack_hi16(q2[4], q2[5]), pack_hi16(q2[6], q2[7]));
        }
      }
      fence_async_smem();
      mbar_arrive(&full[s]);
    }
  } else if (warp == 4) {
    // =========================================================================== MMA issue
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kDim, kN);
      const uint32_t w_addr = smem_u32(sm_w), x_addr = smem_u32(sm_x);
      uint32_t it_count = 0;
      for (uint64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it_count++) {
        const int s = it_count & 1;
        const uint32_t use = it_count >> 1;
        mbar_wait(&full[s], use & 1, p.error_flag);
        if (use > 0) mbar_wait(&tmem_empty[s], (use - 1) & 1, p.error_flag);
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(s * 2 * kN), acc1 = acc0 + kN;
        const uint32_t xs = x_addr + s * kStageBytes;
        // (W slice, X slice) products: (0,0) alone into A0 (exact), the seven lower-order ones into A1
        bool first1 = true;
#pragma unroll
        for (int pw = 0; pw < 3; pw++) {
#pragma unroll
          for (int px = 0; px < 3; px++) {
            if (pw == 2 && px == 2) continue;
            const bool lead = (pw == 0 && px == 0);
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {   // K = 16 per instruction
              const uint64_t ad = make_desc(w_addr + pw * kSliceBytesW + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
              const uint64_t bd = make_desc(xs + px * kSliceBytesX + ks * 2048, 1024, 1024);
              const uint32_t accumulate = lead ? (ks > 0) : !(first1 && ks == 0);
              umma_bf16(lead ? acc0 : acc1, ad, bd, idesc, accumulate);
            }
            if (!lead) first1 = false;
          }
        }
        umma_commit(&mma_done[s]);
      }
    }
  } else {
    // =========================================================================== drain
    const int t128 = threadIdx.x - 160;
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may read
    const int mu = q4 * 32 + lane;           // output row (c', i)
    ItemAddr ia;
    ia.init(p, t128);
    uint32_t it_count = 0;
    for (uint64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it_count++) {
      const int s = it_count & 1;
      const uint32_t use = it_count >> 1;
      mbar_wait(&mma_done[s], use & 1, p.error_flag);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(s * 2 * kN);
      uint8_t* stage = sm_x + s * kStageBytes;    // the X slices of this stage are dead: staging for D
#pragma unroll
      for (int half = 0; half < 2; half++) {      // n = 32 half .. 32 half + 31
        float a0[32], a1[32];
        tmem_ld32(lane_addr + half * 32, a0);
        tmem_ld32(lane_addr + kN + half * 32, a1);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; g++) {             // n = 32 half + 4 g .. + 3: half-buffer (g & 1), chunk q = 4 half + g / 2
          const float4 o = make_float4(a0[4 * g] + a1[4 * g], a0[4 * g + 1] + a1[4 * g + 1], a0[4 * g + 2] + a1[4 * g + 2],
                                       a0[4 * g + 3] + a1[4 * g + 3]);
          *(float4*)(stage + (g & 1) * kSliceBytesX + x_chunk_byte(mu, 4 * half + (g >> 1))) = o;
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[s]);                // the accumulators of this stage may be overwritten
      named_bar(2, 128);
      float4* dst = (float4*)(state + p.tile(tile));
#pragma unroll
      for (int it = 0; it < 4; it++) {
        const float4 r0 = *(const float4*)(stage + ia.soff[it]);
        const float4 r1 = *(const float4*)(stage + kSliceBytesX + ia.soff[it]);
        const float4 i0 = *(const float4*)(stage + ia.soff[it] + 8192u);
        const float4 i1 = *(const float4*)(stage + kSliceBytesX + ia.soff[it] + 8192u);
        float4* d = dst + (ia.goff[it] >> 1);
        __stcs(d + 0, make_float4(r0.x, i0.x, r0.y, i0.y));
        __stcs(d + 1, make_float4(r0.z, i0.z, r0.w, i0.w));
        __stcs(d + 2, make_float4(r1.x, i1.x, r1.y, i1.y));
        __stcs(d + 3, make_float4(r1.z, i1.z, r1.w, i1.w));
      }
      named_bar(2, 128);                          // every drain thread has read its staging rows
      mbar_arrive(&empty[s]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// ------------------------------------------------------------------------------------ host side
// bf16 slices of v on the grid 2^(E - 7 - 8 i): returns the three slice values (exactly representable in bf16)
inline void host_slice3(double v, int E, float out[3]) {
  double r = v;
  for (int i = 0; i < 3; i++) {
    const double g = std::ldexp(1.0, E - 7 - 8 * i);
    const double pi = std::nearbyint(r / g) * g;
    out[i] = (float)pi;
    r -= pi;
  }
}
inline uint16_t host_bf16_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)(u >> 16);   // exact: the slices have at most 8 significant bits
}

// W: 64 x 64 complex (row-major, double re / im pairs), index bit b of the block index = block qubit j_b.
// Real-ified R[(c', i), (c, j)]: [[Wr, -Wi], [Wi, Wr]].  Returns the 96 KiB shared-memory image; E_w is chosen so
// that 2^E_w >= max |entry| (0 for unitaries).
inline std::vector<uint8_t> make_w_image(const double* w_re_im) {
  double mx = 0;
  for (int i = 0; i < 64 * 64 * 2; i++) mx = std::max(mx, std::fabs(w_re_im[i]));
  int E = 0;
  while (std::ldexp(1.0, E) < mx) E++;
  std::vector<uint8_t> img(kSmemW, 0);
  for (int m = 0; m < kDim; m++) {
    for (int k = 0; k < kDim; k++) {
      const int cp = m >> 6, i = m & 63, c = k >> 6, j = k & 63;
      const double wr = w_re_im[2 * (i * 64 + j)], wi = w_re_im[2 * (i * 64 + j) + 1];
      const double r = (cp == c) ? wr : (cp == 1 ? wi : -wi);
      float sl[3];
      host_slice3(r, E, sl);
      for (int s = 0; s < 3; s++) {
        const uint16_t b = host_bf16_bits(sl[s]);
        memcpy(&img[(size_t)s * kSliceBytesW + w_byte(m, k)], &b, 2);
      }
    }
  }
  return img;
}

// Geometry of a pass: `block` = the 6 physical positions of the block qubits in the order of W's index bits
// (index bit b <-> block[b]); every position must be >= 3.  Fills pos / j_of / n_of / tile.
inline const char* make_params(const int* block, int n_qubits, Params* p, int* w_bit_of_jbit /* [6] */) {
  std::vector<int> bits = {0, 1, 2, 3, 4};
  for (int b = 0; b < 6; b++) {
    if (block[b] < 3 || block[b] >= n_qubits) return "block qubits must lie in [3, n).";
    if (std::find(bits.begin(), bits.end(), block[b]) == bits.end()) bits.push_back(block[b]);
  }
  for (int q = 5; (int)bits.size() < kTileBits && q < n_qubits; q++)
    if (std::find(bits.begin(), bits.end(), q) == bits.end()) bits.push_back(q);
  if ((int)bits.size() != kTileBits) return "register too small for a tensor-core pass.";
  std::sort(bits.begin(), bits.end());
  auto is_block = [&](int pos) { return std::find(block, block + 6, pos) != block + 6; };
  // index bits: tile bits 0..2 are n0..n2; tile bit 3 + i takes n_{3+i} (rest) or j_i (block), so that the three
  // lowest thread bits always move the swizzled 16-byte slot (bank-conflict-free STS.128 / LDS.128)
  std::vector<int> rest_pool, block_pool;
  for (int t = 6; t < kTileBits; t++) (is_block(bits[t]) ? block_pool : rest_pool).push_back(t);
  for (int t = 0; t < kTileBits; t++) {
    p->pos[t] = bits[t];
    p->j_of[t] = p->n_of[t] = -1;
  }
  bool n_used[6] = {true, true, true, false, false, false}, j_used[6] = {false, false, false, false, false, false};
  for (int t = 0; t < 3; t++) p->n_of[t] = t;
  for (int i = 0; i < 3; i++) {
    const int t = 3 + i;
    if (is_block(bits[t])) { p->j_of[t] = i; j_used[i] = true; }
    else { p->n_of[t] = 3 + i; n_used[3 + i] = true; }
  }
  for (int t : rest_pool) {
    int k = 0;
    while (k < 6 && n_used[k]) k++;
    if (k == 6) return "internal: too many rest bits.";
    p->n_of[t] = k;
    n_used[k] = true;
  }
  for (int t : block_pool) {
    int k = 0;
    while (k < 6 && j_used[k]) k++;
    if (k == 6) return "internal: too many block bits.";
    p->j_of[t] = k;
    j_used[k] = true;
  }
  // kernel index bit j_k <-> which index bit of the caller's W
  for (int t = 0; t < kTileBits; t++)
    if (p->j_of[t] >= 0) {
      int b = 0;
      while (block[b] != bits[t]) b++;
      w_bit_of_jbit[p->j_of[t]] = b;
    }
  // tile number -> base over the remaining positions
  std::vector<int> other;
  for (int q = 0; q < n_qubits; q++)
    if (std::find(bits.begin(), bits.end(), q) == bits.end()) other.push_back(q);
  p->tile.nseg = 0;
  for (size_t i = 0; i < other.size();) {
    size_t j = i + 1;
    while (j < other.size() && other[j] == other[j - 1] + 1) j++;
    if (p->tile.nseg == 8) return "tile bit set too fragmented.";
    p->tile.src[p->tile.nseg] = (unsigned char)i;
    p->tile.dst[p->tile.nseg] = (unsigned char)other[i];
    p->tile.width[p->tile.nseg] = (unsigned char)(j - i);
    p->tile.nseg++;
    i = j;
  }
  p->ntiles = 1ull << (n_qubits - kTileBits);
  return nullptr;
}
}  // namespace tcb
