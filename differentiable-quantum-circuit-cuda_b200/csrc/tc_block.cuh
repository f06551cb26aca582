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
// shrink of ~5e-7 per block -- more than 1e-5 after the ~70 block applications (forward + un-compute) a depth-100
// amplitude passes through.  The kernels therefore work on exact slices: every f32 value x of a tile is split, on the
// grid g0 = 2^(E-8) (2^E > the largest |x| the fill warp has seen, see warp_max_exp / fill_grid_seed), into three
// BF16 numbers of 9 significant bits each
//
//     x = p0 + p1 + p2 + r,   p_i = k_i * g0 * 2^(-9 i),  |k_i| <= 256,  |r| <= 2^(E-27)
//
// (three magic-number roundings on the packed FP32 pipe, 8 add / fma.rn.f32x2 per PAIR of values; the high 16 bits
// of each f32 slice ARE the bf16: |k| <= 256 is exact in its 8 significant bits).  W is sliced the same way on the
// host (grid 2^-8 for unitaries).  The leading products p0(W) * p0(X) are integers of at most 2^16 on a common grid;
// their sum over K = 128 stays below 2^24, so the tensor core adds them EXACTLY in its own TMEM accumulator A0 whatever
// its rounding mode.  The lower-order products -- orders 1 and 2 by default (6 products in all), order 3 as well with
// `products` = 8 -- go to a second accumulator A1 whose truncation errors are 2^-9 smaller than the result's last bit.
// D = A0 + A1 is one rounded f32 addition in the epilogue.  Error per block: the quantisation of the inputs, unbiased
// (measured against a double-precision host evaluation: 6e-8 (8 products) / 2.7e-7 (6) of the largest amplitude).
//
// Operand layouts (what the UMMA descriptors describe):
//   * W slice i : A operand in TENSOR MEMORY (lane = row mu, 32-bit column = two bf16 along K: 64 columns per slice),
//     copied once per launch from the global image of make_w_image with tcgen05.st.
//   * X slice j : B operand in shared memory, MN-major (n contiguous), SWIZZLE_128B, bf16: row kappa = 128 B = 64 n.
//     The block-gradient products of tc_rev.cuh read the SAME slices along their rows (K-major descriptors).
//   * output staging (f32), two half-buffers with the SAME row / chunk structure as an X slice, so that the
//     drain is the mirror image of the fill.
// The fill goes through registers (LDG.128 -> slices -> STS.128): the slicing needs the CUDA cores
// anyway, and the interleaved (re, im) HBM layout of the reference is de-interleaved on the way.
//
// Kernels: k_tc_block_fwd (state <- W state: forward blocks, and W^dagger / W^T for the three-sweep reverse step),
// k_tc_block_grad (stand-alone block gradient of the three-sweep reverse step, option tc_rev = 0) and, in tc_rev.cuh,
// k_tc_block_rev (the fused reverse step that the executor uses by default).  Warp roles of all three (416 threads,
// one CTA per SM, persistent over tiles):
//   warps 0-7  fill  : HBM -> registers (the next tile's loads in flight) -> slicing grid -> slices -> smem stage, arrive `full`
//   warp  8    MMA   : converged warp, ONE elect.sync lane issues the tcgen05.mma of a tile, tcgen05.commit -> mbarrier
//   warps 9-12 drain : tcgen05.ld A0, A1 -> D -> smem staging -> registers -> HBM
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
constexpr int kSliceBytesW = kDim * kDim * 2;          // 32 KiB per W slice (global image; lives in TMEM in the kernel)
constexpr int kSliceBytesX = kDim * kN * 2;            // 16 KiB per X slice
constexpr int kStageBytes = kSlices * kSliceBytesX;    // 48 KiB per pipeline stage
constexpr int kStages = 4;                             // shared-memory stages (fill may run 3 tiles ahead of the drain)
constexpr int kAccStages = 2;                          // TMEM accumulator stages
constexpr int kImageW = kSlices * kSliceBytesW;        // 96 KiB
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 512 /*barriers*/;
constexpr int kFillWarps = 8, kDrainWarps = 4;
constexpr int kFillThreads = kFillWarps * 32, kDrainThreads = kDrainWarps * 32;
constexpr int kThreads = kFillThreads + 32 + kDrainThreads;   // 416
constexpr int kTmemCols = 512;     // W slices: 3 x 64 columns (bf16 pairs); accumulators: columns 256.. : 2 stages x (A0, A1) x 64
constexpr int kTmemAcc = 256;
#ifndef TC_PREFETCH
#define TC_PREFETCH 0
#endif
// tiles of L2 prefetch distance per CTA (one cp.async.bulk.prefetch.L2 per 256-byte run).  0 = off, the default: the
// fill's register pipeline already has a tile in flight half a period ahead of its use; prefetching on top of it made
// the loads return at once but the kernels slower (30 q reverse step 10.17 / 10.54 / 11.11 ms at distance 0 / 4 / 8,
// forward block 0.954 / 1.068 ms at 28 q: profiles/r2_tc_rev_bench_v7.txt)
constexpr int kPrefetch = TC_PREFETCH;

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
  // block index j (0..5) or of the rest index n (0..5).
  int pos[kTileBits];
  int j_of[kTileBits];   // -1 if the tile bit is a rest bit
  int n_of[kTileBits];   // -1 if the tile bit is a block bit
  // n0..n2 select the 8 amplitudes of one 16-byte bf16 chunk; the other nine tile bits, ascending, are the item bits
  // (thread / iteration), so that consecutive threads walk through consecutive memory.  `fast`: n0 is physical
  // position 0 and n1, n2 are the two HIGHEST rest bits -- an item is four 128-bit accesses, and a warp instruction
  // covers whole 256-byte runs (positions 1..4 are thread bits).  Otherwise (position 0 belongs to the block) n0..n2 are
  // the three highest rest bits and an item is eight 64-bit accesses, again contiguous across the warp.
  int item_tb[9];
  int elem_off[8];       // amplitude offset of element e of an item (deposit of e into the positions of n0..n2)
  int fast;
  Deposit tile;          // tile number -> amplitude base over the other n - 12 positions
  uint64_t tile_mask;    // the 12 positions of the tile as an index mask (TileWalk)
  uint64_t ntiles;
  const uint32_t* w_image; // 96 KiB: [slice][row mu][64 words], word k2 = (bf16 of column 2 k2) | (bf16 of column 2 k2 + 1) << 16
  int* error_flag;         // set to 1 by a watchdog if a barrier wait times out
  int products;            // 8: all slice products but p2 * p2;  6: also without p1 * p2, p2 * p1 (2^-24 relative each)
};

// Amplitude base of a CTA's tiles, advanced by a fixed stride WITHOUT the software bit deposit (a loop of 64-bit
// shifts over constant-bank tables that every thread paid several times per tile: 14 % of the warp stall samples of
// profiles/r2_tc_rev_28q_v2_ncu.txt).  With F = the tile's own positions, which the tile counter skips,
// deposit(a + b) = ((deposit(a) | F) + deposit(b)) & ~F: the carries run through the filled positions.
struct TileWalk {
  uint64_t cur, step, skip;
  __device__ __forceinline__ void init(const Params& p, uint64_t first, uint64_t stride) {
    cur = p.tile(first);
    step = p.tile(stride);
    skip = p.tile_mask;
  }
  __device__ __forceinline__ void advance() { cur = ((cur | skip) + step) & ~skip; }
};

// ------------------------------------------------------------------ byte images (host and device)
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
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred)
      :
      : "memory");
  return pred != 0;
}
// shared-memory matrix descriptors (see make_desc) as a base word plus byte offsets: SBO = 1024 everywhere
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ uint64_t desc_at(uint32_t lo, uint32_t hi, uint32_t byte_off) {
  return ((uint64_t)hi << 32) | (uint64_t)(lo + (byte_off >> 4));
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
// A operand from tensor memory (lanes = rows, 32-bit columns = pairs of bf16 along K)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 registers per thread -> 32 consecutive 32-bit columns of the thread's lane
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
// the same for two values at once on the packed FP32 pipe (add / fma .f32x2 of sm_100: 8 instructions per PAIR)
#ifndef TC_FADD2
#define TC_FADD2 1
#endif
__device__ __forceinline__ void slice3_pair(float xa, float xb, float m0, float m1, float m2, float& a0, float& a1, float& a2,
                                            float& b0, float& b1, float& b2) {
#if TC_FADD2
  unsigned long long x, M0, N0, M1, N1, M2, N2, neg1, t, q0, q1, q2, r1, r2;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(xa), "f"(xb));
  asm("mov.b64 %0, {%1, %1};" : "=l"(M0) : "f"(m0));
  asm("mov.b64 %0, {%1, %1};" : "=l"(N0) : "f"(-m0));
  asm("mov.b64 %0, {%1, %1};" : "=l"(M1) : "f"(m1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(N1) : "f"(-m1));
  asm("mov.b64 %0, {%1, %1};" : "=l"(M2) : "f"(m2));
  asm("mov.b64 %0, {%1, %1};" : "=l"(N2) : "f"(-m2));
  asm("mov.b64 %0, {%1, %1};" : "=l"(neg1) : "f"(-1.0f));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(x), "l"(M0));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(q0) : "l"(t), "l"(N0));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r1) : "l"(q0), "l"(neg1), "l"(x));     // x - p0 (exact)
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(r1), "l"(M1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(q1) : "l"(t), "l"(N1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r2) : "l"(q1), "l"(neg1), "l"(r1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(r2), "l"(M2));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(q2) : "l"(t), "l"(N2));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(b0) : "l"(q0));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a1), "=f"(b1) : "l"(q1));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(b2) : "l"(q2));
#else
  slice3(xa, m0, m1, m2, a0, a1, a2);
  slice3(xb, m0, m1, m2, b0, b1, b2);
#endif
}
// the high halves of two f32 words -> one word of two bf16 (lo = first)
__device__ __forceinline__ uint32_t pack_hi16(float a, float b) {
  return __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632);
}

// per-thread addressing of the NI items (8 consecutive amplitudes each) a fill / drain thread of a group of
// NT threads owns: item = it * NT + t, 9 bits <-> the item tile bits (Params::item_tb)
template <int NI, int NT>
struct ItemAddr {
  uint64_t goff[NI];   // amplitude offset inside the tile span
  uint32_t soff[NI];   // byte offset of the item's 16-byte chunk in row (c = 0, j) of an X slice
  __device__ __forceinline__ void init(const Params& p, int t) {
#pragma unroll
    for (int it = 0; it < NI; it++) {
      const int item = it * NT + t;
      uint64_t g = 0;
      int j = 0, n = 0;
#pragma unroll
      for (int b = 0; b < 9; b++) {
        if ((item >> b) & 1) {
          const int tb = p.item_tb[b];
          g |= 1ull << p.pos[tb];
          if (p.j_of[tb] >= 0) j |= 1 << p.j_of[tb]; else n |= 1 << p.n_of[tb];
        }
      }
      goff[it] = g;
      soff[it] = x_chunk_byte(j, n >> 3);
    }
  }
};

// the 8 amplitudes of an item: four 128-bit accesses (fast: elements 2h, 2h+1 are adjacent) or eight 64-bit ones
__device__ __forceinline__ void load_item(const float2* __restrict__ src, const Params& p, float4 (&v)[4]) {
  if (p.fast) {
#pragma unroll
    for (int h = 0; h < 4; h++) v[h] = __ldcs((const float4*)(src + p.elem_off[2 * h]));
  } else {
#pragma unroll
    for (int h = 0; h < 4; h++) {
      const float2 a = __ldcs(src + p.elem_off[2 * h]), b = __ldcs(src + p.elem_off[2 * h + 1]);
      v[h] = make_float4(a.x, a.y, b.x, b.y);
    }
  }
}
__device__ __forceinline__ void prefetch_item(const float2* src, const Params& p) {
#pragma unroll
  for (int h = 0; h < 4; h++) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(src + p.elem_off[2 * h]));
    if (!p.fast) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + p.elem_off[2 * h + 1]));
  }
}
// One instruction per 256-byte run: every tile holds the 5 lowest positions, so it is 128 contiguous runs of 32
// amplitudes.  (The per-item prefetches above cost a fill thread 8 instructions per 64 bytes.)
__device__ __forceinline__ void prefetch_run_l2(const float2* run_start) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], 256;" ::"l"(run_start) : "memory");
}
// amplitude offset of run r (0..127) of a tile: the bits of r on the tile's 7 upper positions
__device__ __forceinline__ uint64_t run_offset(const Params& p, int r) {
  uint64_t g = 0;
#pragma unroll
  for (int b = 0; b < 7; b++)
    if ((r >> b) & 1) g |= 1ull << p.pos[5 + b];
  return g;
}
__device__ __forceinline__ void store_item(float2* __restrict__ dst, const Params& p, const float4 (&v)[4]) {
  if (p.fast) {
#pragma unroll
    for (int h = 0; h < 4; h++) __stcs((float4*)(dst + p.elem_off[2 * h]), v[h]);
  } else {
#pragma unroll
    for (int h = 0; h < 4; h++) {
      __stcs(dst + p.elem_off[2 * h], make_float2(v[h].x, v[h].y));
      __stcs(dst + p.elem_off[2 * h + 1], make_float2(v[h].z, v[h].w));
    }
  }
}

// magic numbers of the 9-bit slicing (grids 2^(E - 8), 2^(E - 17), 2^(E - 26), 2^E > max), see k_tc_block_fwd
__device__ __forceinline__ void magic_of(uint32_t bexp, float& m0, float& m1, float& m2) {
  bexp = bexp < 32u ? 32u : (bexp > 230u ? 230u : bexp);
  const uint32_t m0b = ((bexp + 16u) << 23) | 0x400000u;
  m0 = __uint_as_float(m0b);
  m1 = __uint_as_float(m0b - (9u << 23));
  m2 = __uint_as_float(m0b - (18u << 23));
}

// Slicing grid of a fill warp WITHOUT a CTA-wide barrier per tile: the exponent of the largest |value| this warp holds
// (one REDUX), raised to the warp's own running maximum.  The fill warps agree ONCE, on the first tile of the CTA
// (fill_grid_seed: one named barrier per kernel), and from then on each warp only ever raises its own grid: the result is
// deterministic (no unsynchronised reads), and the slices are ordinary f32 / bf16 VALUES, so a warp that is a binade
// ahead of the others costs nothing but the slack of the exactness bound (K = 128 leading products of <= 2^16 grid
// units leave one binade below 2^24).  The absolute quantisation error stays 2^-27 of the largest amplitude the warp has
// seen.
__device__ __forceinline__ uint32_t warp_max_exp(const float4 (&v)[2][4]) {
  float mx = 0.f;
#pragma unroll
  for (int it = 0; it < 2; it++)
#pragma unroll
    for (int h = 0; h < 4; h++)
      mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[it][h].x), fabsf(v[it][h].y))), fmaxf(fabsf(v[it][h].z), fabsf(v[it][h].w)));
  return __reduce_max_sync(0xffffffffu, (__float_as_uint(mx) >> 23) & 0xffu);
}
// common starting grid of the fill warps: the maximum exponent of the CTA's first tile (e_shared: zero on entry)
__device__ __forceinline__ uint32_t fill_grid_seed(const float4 (&v)[2][4], uint32_t* e_shared, int lane) {
  const uint32_t e = warp_max_exp(v);
  if (lane == 0) atomicMax(e_shared, e);
  named_bar(1, kFillThreads);
  return *(volatile uint32_t*)e_shared;
}

// ------------------------------------------------------------------------------------ the kernel
// state: 2^n interleaved (re, im) f32 pairs, updated in place: every group of 64 amplitudes over the block
// qubits is multiplied by W.
//
// Warp roles (416 threads, one CTA per SM, persistent over tiles):
//   warps 0-7   fill  : HBM -> registers (the next tile's loads are issued as soon as this tile has been sliced) ->
//                       slicing grid (per warp, seeded once per kernel) -> slices -> smem stage (4 stages), arrive `full`
//   warp  8     MMA   : converged warp, proxy fence after `full`, one elected lane issues 48 / 64 tcgen05.mma per tile
//                       (A = W slices in TMEM, B = X slices in smem), tcgen05.commit -> `mma_done`
//   warps 9-12  drain : tcgen05.ld A0, A1 -> D = A0 + A1 -> smem staging (the stage's own, now dead, X slices) ->
//                       registers, arrive `tmem_empty` (2 accumulator stages) and `empty` -> HBM
__global__ void __launch_bounds__(kThreads, 1) k_tc_block_fwd(float2* __restrict__ state, const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms: 1024-byte aligned (pointer arithmetic on the array keeps the shared state space: a cast
  // through uintptr_t made every access a generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sm_x = smem;
  uint64_t* bars = (uint64_t*)(smem + kStages * kStageBytes);
  uint64_t* full = bars;                       // [kStages]    fill -> MMA    (256 arrivals)
  uint64_t* empty = bars + kStages;            // [kStages]    drain -> fill  (128 arrivals)
  uint64_t* mma_done = bars + 2 * kStages;     // [kStages]    MMA -> drain   (tcgen05.commit)
  uint64_t* tmem_empty = bars + 3 * kStages;   // [kAccStages] drain -> MMA   (128 arrivals)
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * kStages + kAccStages);
  uint32_t* e_shared = tmem_slot + 2;          // maximum exponent of the CTA's first tile (fill_grid_seed)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup
  if (threadIdx.x == 0) {
    *e_shared = 0;
    for (int s = 0; s < kStages; s++) {
      mbar_init(&full[s], kFillThreads);
      mbar_init(&empty[s], kDrainThreads);
      mbar_init(&mma_done[s], 1);
    }
    for (int s = 0; s < kAccStages; s++) mbar_init(&tmem_empty[s], kDrainThreads);
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

  if (warp < kFillWarps) {
    // =========================================================================== fill
    const int t = threadIdx.x;
    ItemAddr<2, kFillThreads> ia;
    ia.init(p, t);
    TileWalk wl, wp;   // base of the tile whose loads are issued next / of the tile prefetched into L2 next
    wl.init(p, blockIdx.x, gridDim.x);
    wp = wl;
    // L2 prefetch of the first tiles of this CTA: threads 0..127 own one 256-byte run each
    const uint64_t roff = run_offset(p, t & 127);
    for (int k = 0; k < kPrefetch; k++) {
      const uint64_t tl = blockIdx.x + (uint64_t)k * gridDim.x;
      if (tl < p.ntiles && t < 128) prefetch_run_l2(state + wp.cur + roff);
      wp.advance();
    }
    uint32_t it_count = 0, e_run = 0;
    // Software pipeline over registers: the loads of the NEXT tile are issued as soon as this tile's slices have been
    // taken, so that a whole tile (32 KiB, ~1500 cycles at this SM's share of the HBM bandwidth) is in flight while
    // the fill fences, waits for its stage and reduces the next maximum.  The L2 prefetch runs kPrefetch tiles ahead.
    float4 v[2][4];
    if ((uint64_t)blockIdx.x < p.ntiles) {
      const float2* src = state + wl.cur;
#pragma unroll
      for (int it = 0; it < 2; it++) load_item(src + ia.goff[it], p, v[it]);
      e_run = fill_grid_seed(v, e_shared, lane);
    }
    for (uint64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it_count++) {
      const int s = it_count % kStages;
      const uint32_t use = it_count / kStages;
      if (kPrefetch > 0) {
        if (tile + (uint64_t)kPrefetch * gridDim.x < p.ntiles && t < 128) prefetch_run_l2(state + wp.cur + roff);
        wp.advance();
      }
      // 9-bit slices: grids 2^(E - 8), 2^(E - 17), 2^(E - 26) with 2^E > max: |k_i| <= 256 is still exact in bf16
      // (8 significant bits) and the leading products stay exact: 128 * 2^16 = 2^23 < 2^24
      e_run = max(e_run, warp_max_exp(v));
      float m0, m1, m2;
      magic_of(e_run, m0, m1, m2);
      if (use > 0) mbar_wait(&empty[s], (use - 1) & 1, p.error_flag);
      uint8_t* stage = sm_x + s * kStageBytes;
#pragma unroll
      for (int it = 0; it < 2; it++) {
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
          for (int e = 0; e < 8; e += 2)
            slice3_pair(c ? im[e] : re[e], c ? im[e + 1] : re[e + 1], m0, m1, m2, q0[e], q1[e], q2[e], q0[e + 1], q1[e + 1], q2[e + 1]);
          const uint32_t off = ia.soff[it] + (uint32_t)c * 8192u;   // row (c, j) = row j + 64
          *(uint4*)(stage + 0 * kSliceBytesX + off) =
              make_uint4(pack_hi16(q0[0], q0[1]), pack_hi16(q0[2], q0[3]), pack_hi16(q0[4], q0[5]), pack_hi16(q0[6], q0[7]));
          *(uint4*)(stage + 1 * kSliceBytesX + off) =
              make_uint4(pack_hi16(q1[0], q1[1]), pack_hi16(q1[2], q1[3]), pack_hi16(q1[4], q1[5]), pack_hi16(q1[6], q1[7]));
          *(uint4*)(stage + 2 * kSliceBytesX + off) =
              make_uint4(pack_hi16(q2[0], q2[1]), pack_hi16(q2[2], q2[3]), pack_hi16(q2[4], q2[5]), pack_hi16(q2[6], q2[7]));
        }
      }
      wl.advance();
      if (tile + gridDim.x < p.ntiles) {
        const float2* src = state + wl.cur;
#pragma unroll
        for (int it = 0; it < 2; it++) load_item(src + ia.goff[it], p, v[it]);
      }
      // (the proxy fence is executed by the MMA warp after it has acquired `full`: here it would wait for the loads
      // that this thread has just issued)
      mbar_arrive(&full[s]);
    }
  } else if (warp == kFillWarps) {
    // =========================================================================== MMA issue
    // The whole warp runs the loop converged and one ELECTED lane issues: from `if (lane == 0)` ptxas wrapped every
    // UTCHMMA in a leader-election loop (50-80 cycles per instruction, more than the 32 the tensor pipe needs).
    constexpr uint32_t idesc = make_idesc(kDim, kN);
    const uint32_t x_addr = smem_u32(sm_x);
    const bool all8 = p.products >= 8;
    uint32_t it_count = 0;
    for (uint64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it_count++) {
      const int s = it_count % kStages, as = it_count % kAccStages;
      const uint32_t use = it_count / kStages, ause = it_count / kAccStages;
      mbar_wait(&full[s], use & 1, p.error_flag);
      fence_async_smem();   // generic-proxy stores of the fill warps (acquired through `full`) -> async-proxy reads of the MMAs
      if (ause > 0) mbar_wait(&tmem_empty[as], (ause - 1) & 1, p.error_flag);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t acc0 = tmem_base + kTmemAcc + (uint32_t)(as * 2 * kN), acc1 = acc0 + kN;
        const uint32_t x_lo = desc_lo(x_addr + s * kStageBytes, 1024);
        // (W slice, X slice) products: (0,0) alone into A0 (exact), the lower-order ones into A1
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
              const uint64_t bd = desc_at(x_lo, kDescHi, px * kSliceBytesX + ks * 2048);
              const uint32_t accumulate = lead ? (ks > 0) : !(first1 && ks == 0);
              umma_bf16_ts(lead ? acc0 : acc1, tmem_base + (uint32_t)(pw * 64 + ks * 8), bd, idesc, accumulate);
            }
            if (!lead) first1 = false;
          }
        }
        umma_commit(&mma_done[s]);
      }
      __syncwarp();
    }
  } else {
    // =========================================================================== drain
    const int t128 = threadIdx.x - (kFillThreads + 32);
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may access
    const int mu = q4 * 32 + lane;           // output row (c', i)
    ItemAddr<4, kDrainThreads> ia;
    ia.init(p, t128);
    TileWalk wd;
    wd.init(p, blockIdx.x, gridDim.x);
    uint32_t it_count = 0;
    for (uint64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it_count++, wd.advance()) {
      const int s = it_count % kStages, as = it_count % kAccStages;
      const uint32_t use = it_count / kStages;
      mbar_wait(&mma_done[s], use & 1, p.error_flag);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16) + kTmemAcc + (uint32_t)(as * 2 * kN);
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
      mbar_arrive(&tmem_empty[as]);               // the accumulators of this stage may be overwritten
      named_bar(2, kDrainThreads);
      float2* dst = state + wd.cur;
      // staging -> registers; the stage is free for the fill as soon as every drain thread has read its rows -- the
      // global stores follow from the registers
      float4 r[4][4];
#pragma unroll
      for (int it = 0; it < 4; it++) {
        r[it][0] = *(const float4*)(stage + ia.soff[it]);
        r[it][1] = *(const float4*)(stage + kSliceBytesX + ia.soff[it]);
        r[it][2] = *(const float4*)(stage + ia.soff[it] + 8192u);
        r[it][3] = *(const float4*)(stage + kSliceBytesX + ia.soff[it] + 8192u);
      }
      named_bar(2, kDrainThreads);
      mbar_arrive(&empty[s]);
#pragma unroll
      for (int it = 0; it < 4; it++) {
        const float4 r0 = r[it][0], r1 = r[it][1], i0 = r[it][2], i1 = r[it][3];
        const float4 o[4] = {make_float4(r0.x, i0.x, r0.y, i0.y), make_float4(r0.z, i0.z, r0.w, i0.w),
                             make_float4(r1.x, i1.x, r1.y, i1.y), make_float4(r1.z, i1.z, r1.w, i1.w)};
        store_item(dst + ia.goff[it], p, o);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kFillWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// ------------------------------------------------------------------------------------ gradient kernel
// Block gradient of the reverse pass: with a = the state BEFORE the block (already un-computed) and b = the
// adjoint AFTER the block (not yet pulled back),
//
//     P[mu, nu] = sum over every group of 64 amplitudes and every rest index n of  B~[mu, n] * A~[nu, n]
//
// (B~, A~ the real-ified 128-row forms), from which G_W[i, j] = sum b[i] a[j] (no conjugation, the convention of
// src/primitives.cu:323-329) is  Re = P[i, j] - P[64 + i, 64 + j],  Im = P[i, 64 + j] + P[64 + i, j].
// One tcgen05.mma chain per tile: M = 128 (rows of b), N = 128 (rows of a), K = 64 (n), both operands K-major --
// the X-slice layout of the forward kernel read along its rows.  9-bit slices as in the forward kernel; the leading
// product p0(b) * p0(a) accumulates exactly in A0 as long as the grids of consecutive tiles agree (a running maximum
// of the tile maxima keeps them equal after the first tiles of a CTA) and the sum stays below 2^24 grid units
// (64 * 2^16 = 2^22 per tile in the worst case; the window is flushed every kFlush = 4 tiles); the five
// lower-order products go to A1.  A flush adds A0 + A1 into the
// CTA's private f32 partial in global memory (single writer per element: deterministic); the partials are summed
// in double by k_tc_grad_reduce.
constexpr int kGradStageBytes = 2 * kStageBytes;   // b slices, then a slices: 96 KiB
constexpr int kGradStages = 2;
constexpr int kGradSmemBytes = kGradStages * kGradStageBytes + 1024 + 512;
constexpr int kFlush = 4;
constexpr int kGradPrefetch = 5;   // tiles of L2 prefetch distance (this kernel loads a tile at the start of its iteration)

struct GradParams {
  Params geo;          // pos / j_of / n_of / tile / ntiles / error_flag (w_image, products unused)
  float* partials;     // [gridDim.x][128][128], zero on entry
};

template <int SEL>   // fill one buffer's slices of a stage from registers
__device__ __forceinline__ void fill_slices(const float4 (&v)[2][4], const uint32_t (&soff)[2], float m0, float m1, float m2,
                                            uint8_t* slices) {
#pragma unroll
  for (int it = 0; it < 2; it++) {
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
      const uint32_t off = soff[it] + (uint32_t)c * 8192u;
      *(uint4*)(slices + 0 * kSliceBytesX + off) =
          make_uint4(pack_hi16(q0[0], q0[1]), pack_hi16(q0[2], q0[3]), pack_hi16(q0[4], q0[5]), pack_hi16(q0[6], q0[7]));
      *(uint4*)(slices + 1 * kSliceBytesX + off) =
          make_uint4(pack_hi16(q1[0], q1[1]), pack_hi16(q1[2], q1[3]), pack_hi16(q1[4], q1[5]), pack_hi16(q1[6], q1[7]));
      *(uint4*)(slices + 2 * kSliceBytesX + off) =
          make_uint4(pack_hi16(q2[0], q2[1]), pack_hi16(q2[2], q2[3]), pack_hi16(q2[4], q2[5]), pack_hi16(q2[6], q2[7]));
    }
  }
}

__device__ __forceinline__ float tile_max8(const float4 (&v)[2][4], float* mxbuf, int warp, int lane, int bar_id) {
  float mx = 0.f;
#pragma unroll
  for (int it = 0; it < 2; it++)
#pragma unroll
    for (int h = 0; h < 4; h++)
      mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v[it][h].x), fabsf(v[it][h].y))), fmaxf(fabsf(v[it][h].z), fabsf(v[it][h].w)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) mxbuf[warp] = mx;
  named_bar(bar_id, kFillThreads);
  return fmaxf(fmaxf(fmaxf(mxbuf[0], mxbuf[1]), fmaxf(mxbuf[2], mxbuf[3])), fmaxf(fmaxf(mxbuf[4], mxbuf[5]), fmaxf(mxbuf[6], mxbuf[7])));
}

__global__ void __launch_bounds__(kThreads, 1)
    k_tc_block_grad(const float2* __restrict__ a_state, const float2* __restrict__ b_state, const __grid_constant__ GradParams gp) {
  const Params& p = gp.geo;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = (uint64_t*)(smem + kGradStages * kGradStageBytes);
  uint64_t* full = bars;            // [2] fill -> MMA  (256 arrivals)
  uint64_t* empty = bars + 2;       // [2] MMA -> fill  (tcgen05.commit: the MMAs reading the stage are done)
  uint64_t* acc_done = bars + 4;    // [2] MMA -> drain (tcgen05.commit at the end of a flush window)
  uint64_t* acc_empty = bars + 6;   // [2] drain -> MMA (128 arrivals)
  uint32_t* tmem_slot = (uint32_t*)(bars + 8);
  float* sm_max = (float*)(tmem_slot + 2);   // [4][8]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; s++) {
      mbar_init(&full[s], kFillThreads);
      mbar_init(&empty[s], 1);
      mbar_init(&acc_done[s], 1);
      mbar_init(&acc_empty[s], kDrainThreads);
    }
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
  // tiles of this CTA and flush windows
  const uint64_t my_tiles = p.ntiles > (uint64_t)blockIdx.x ? (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp < kFillWarps) {
    // =========================================================================== fill (b slices, then a slices)
    const int t = threadIdx.x;
    ItemAddr<2, kFillThreads> ia;
    ia.init(p, t);
    auto load = [&](const float2* base, uint64_t tile, float4 (&v)[2][4]) {
      const float2* src = base + p.tile(tile);
#pragma unroll
      for (int it = 0; it < 2; it++) load_item(src + ia.goff[it], p, v[it]);
    };
    auto prefetch = [&](uint64_t tile) {
      const uint64_t base = p.tile(tile);
#pragma unroll
      for (int it = 0; it < 2; it++) {
        prefetch_item(a_state + base + ia.goff[it], p);
        prefetch_item(b_state + base + ia.goff[it], p);
      }
    };
    for (int k = 0; k < kGradPrefetch; k++)
      if (blockIdx.x + (uint64_t)k * gridDim.x < p.ntiles) prefetch(blockIdx.x + (uint64_t)k * gridDim.x);
    uint32_t e_run_a = 0, e_run_b = 0;
    uint32_t it_count = 0;
    for (uint64_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, it_count++) {
      const int s = it_count & 1;
      const uint32_t use = it_count >> 1;
      uint8_t* stage = smem + s * kGradStageBytes;
      float4 vb[2][4], va[2][4];
      load(b_state, tile, vb);      // out of L2 (prefetched kGradPrefetch tiles ago)
      load(a_state, tile, va);
      if (tile + (uint64_t)kGradPrefetch * gridDim.x < p.ntiles) prefetch(tile + (uint64_t)kGradPrefetch * gridDim.x);
      float mx = tile_max8(vb, sm_max + ((it_count & 1) * 2 + 0) * 8, warp, lane, 1);
      e_run_b = max(e_run_b, (__float_as_uint(mx) >> 23) & 0xffu);
      float m0, m1, m2;
      magic_of(e_run_b, m0, m1, m2);
      if (use > 0) mbar_wait(&empty[s], (use - 1) & 1, p.error_flag);
      fill_slices<0>(vb, ia.soff, m0, m1, m2, stage);
      mx = tile_max8(va, sm_max + ((it_count & 1) * 2 + 1) * 8, warp, lane, 1);
      e_run_a = max(e_run_a, (__float_as_uint(mx) >> 23) & 0xffu);
      magic_of(e_run_a, m0, m1, m2);
      fill_slices<1>(va, ia.soff, m0, m1, m2, stage + kStageBytes);
      fence_async_smem();
      mbar_arrive(&full[s]);
    }
  } else if (warp == kFillWarps) {
    // =========================================================================== MMA issue (converged warp, one elected lane)
    // D f32, A / B bf16, both K-major, M = N = 128
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t base = smem_u32(smem);
    for (uint32_t it = 0; it < (uint32_t)my_tiles; it++) {
      const int s = it & 1;
      const uint32_t use = it >> 1, win = it / kFlush, as = win & 1;
      const bool first = (it % kFlush) == 0, last = (it % kFlush) == kFlush - 1 || it + 1 == (uint32_t)my_tiles;
      mbar_wait(&full[s], use & 1, p.error_flag);
      if (first && win >= 2) mbar_wait(&acc_empty[as], ((win >> 1) - 1) & 1, p.error_flag);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t acc0 = tmem_base + as * 256u, acc1 = acc0 + 128u;
        const uint32_t b_lo = desc_lo(base + s * kGradStageBytes, 16), a_lo = desc_lo(base + s * kGradStageBytes + kStageBytes, 16);
        bool first1 = true;
#pragma unroll
        for (int pb = 0; pb < 3; pb++) {
#pragma unroll
          for (int pa = 0; pa < 3; pa++) {
            if (pb + pa >= 3) continue;      // orders 0, 1, 2: (0,0) | (0,1) (1,0) | (0,2) (1,1) (2,0)
            const bool lead = (pb == 0 && pa == 0);
#pragma unroll
            for (int ks = 0; ks < 4; ks++) {   // K = 16 columns n per instruction: 32 bytes along the 128-byte rows
              const uint32_t accumulate = lead ? !(first && ks == 0) : !(first && first1 && ks == 0);
              umma_bf16(lead ? acc0 : acc1, desc_at(b_lo, kDescHi, pb * kSliceBytesX + ks * 32),
                        desc_at(a_lo, kDescHi, pa * kSliceBytesX + ks * 32), idesc, accumulate);
            }
            if (!lead) first1 = false;
          }
        }
        umma_commit(&empty[s]);
        if (last) umma_commit(&acc_done[as]);
      }
      __syncwarp();
    }
  } else {
    // =========================================================================== drain: flush windows
    const int q4 = warp & 3, mu = q4 * 32 + lane;
    float* mine = gp.partials + ((size_t)blockIdx.x * kDim + mu) * kDim;
    const uint32_t nwin = (uint32_t)((my_tiles + kFlush - 1) / kFlush);
    for (uint32_t win = 0; win < nwin; win++) {
      const uint32_t as = win & 1;
      mbar_wait(&acc_done[as], (win >> 1) & 1, p.error_flag);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16) + as * 256u;
#pragma unroll 1
      for (int ch = 0; ch < 4; ch++) {
        float a0[32], a1[32];
        tmem_ld32(lane_addr + ch * 32, a0);
        tmem_ld32(lane_addr + 128 + ch * 32, a1);
        tmem_ld_wait();
        float4* dst = (float4*)(mine + ch * 32);
#pragma unroll
        for (int g = 0; g < 8; g++) {
          float4 o = dst[g];
          o.x += a0[4 * g] + a1[4 * g];
          o.y += a0[4 * g + 1] + a1[4 * g + 1];
          o.z += a0[4 * g + 2] + a1[4 * g + 2];
          o.w += a0[4 * g + 3] + a1[4 * g + 3];
          dst[g] = o;
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kFillWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

// out[k] (+)= sum over CTAs of partials[cta][k], k < 128 * 128, in double
__global__ void k_tc_grad_reduce(const float* __restrict__ partials, int ncta, double* __restrict__ out, int accumulate) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= kDim * kDim) return;
  double s = 0;
  for (int c = 0; c < ncta; c++) s += (double)partials[(size_t)c * kDim * kDim + k];
  out[k] = accumulate ? out[k] + s : s;
}

// ------------------------------------------------------------------------------------ host side
// bf16 slices of v on the grids 2^(E - 8 - 9 i): returns the three slice values (exactly representable in bf16)
inline void host_slice3(double v, int E, float out[3]) {
  double r = v;
  for (int i = 0; i < 3; i++) {
    const double g = std::ldexp(1.0, E - 8 - 9 * i);   // 9-bit slices, |k| <= 256 (see k_tc_block_fwd)
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

// W: 64 x 64 complex (row-major, double re / im pairs), index bit b of the block index = kernel index bit j_b.
// Real-ified R[(c', i), (c, j)]: [[Wr, -Wi], [Wi, Wr]].  Returns the 96 KiB image [slice][row][64 words] the kernel
// copies into tensor memory; the slicing grid 2^(E_w - 7) has 2^E_w >= max |entry| (E_w = 0 for unitaries).
// `odd_low`: development switch for the order of the two bf16 of a 32-bit TMEM column (default: even K index low).
inline std::vector<uint32_t> make_w_image(const double* w_re_im, bool odd_low = false) {
  double mx = 0;
  for (int i = 0; i < 64 * 64 * 2; i++) mx = std::max(mx, std::fabs(w_re_im[i]));
  int E = 0;
  while (std::ldexp(1.0, E) < mx) E++;
  std::vector<uint32_t> img(kImageW / 4, 0u);
  for (int m = 0; m < kDim; m++) {
    for (int k = 0; k < kDim; k++) {
      const int cp = m >> 6, i = m & 63, c = k >> 6, j = k & 63;
      const double wr = w_re_im[2 * (i * 64 + j)], wi = w_re_im[2 * (i * 64 + j) + 1];
      const double r = (cp == c) ? wr : (cp == 1 ? wi : -wi);
      float sl[3];
      host_slice3(r, E, sl);
      for (int s = 0; s < 3; s++) {
        const uint32_t b = host_bf16_bits(sl[s]);
        const bool high = ((k & 1) != 0) != odd_low;
        img[((size_t)s * kDim + m) * 64 + (k >> 1)] |= high ? (b << 16) : b;
      }
    }
  }
  return img;
}

// Geometry of a pass: `block` = the 6 physical positions of the block qubits in the order of the caller's index
// bits (index bit b <-> block[b]).  Fills pos / j_of / n_of / item_tb / elem_off / tile; w_bit_of_jbit[k] = the
// caller's index bit that the kernel's block-index bit k stands for.
inline const char* make_params(const int* block, int n_qubits, Params* p, int* w_bit_of_jbit /* [6] */) {
  std::vector<int> bits = {0, 1, 2, 3, 4};   // every tile holds the 5 lowest positions: 256-byte runs in HBM
  for (int b = 0; b < 6; b++) {
    if (block[b] < 0 || block[b] >= n_qubits) return "block qubit out of range.";
    for (int c = 0; c < b; c++)
      if (block[c] == block[b]) return "block qubits must be distinct.";
    if (std::find(bits.begin(), bits.end(), block[b]) == bits.end()) bits.push_back(block[b]);
  }
  for (int q = 5; (int)bits.size() < kTileBits && q < n_qubits; q++)
    if (std::find(bits.begin(), bits.end(), q) == bits.end()) bits.push_back(q);
  if ((int)bits.size() != kTileBits) return "register too small for a tensor-core pass.";
  std::sort(bits.begin(), bits.end());
  auto is_block = [&](int pos) { return std::find(block, block + 6, pos) != block + 6; };
  for (int t = 0; t < kTileBits; t++) {
    p->pos[t] = bits[t];
    p->j_of[t] = p->n_of[t] = -1;
  }
  // n0..n2: position 0 when it is a rest bit, then the highest rest tile bits (see Params)
  int nlow[3], nl = 0;
  if (!is_block(0)) nlow[nl++] = 0;
  {
    std::vector<int> high;
    for (int t = kTileBits - 1; t >= 1 && (int)high.size() < 3 - nl; t--)
      if (!is_block(bits[t])) high.push_back(t);
    if ((int)high.size() < 3 - nl) return "internal: fewer than three rest bits.";
    for (int k = (int)high.size() - 1; k >= 0; k--) nlow[nl++] = high[k];   // ascending
  }
  for (int k = 0; k < 3; k++) p->n_of[nlow[k]] = k;
  p->fast = !is_block(0) ? 1 : 0;
  for (int e = 0; e < 8; e++) {
    long off = 0;
    for (int k = 0; k < 3; k++)
      if ((e >> k) & 1) off |= 1l << bits[nlow[k]];
    if (off > 0x7fffffffl) return "tile bit set too spread out.";
    p->elem_off[e] = (int)off;
  }
  // item bits: the other nine tile bits, ascending.  Item bit i < 3 takes n_{3+i} (rest) or j_i (block), so that the
  // three lowest thread bits always move the swizzled 16-byte slot (bank-conflict-free STS.128 / LDS.128)
  int ni = 0;
  for (int t = 0; t < kTileBits; t++)
    if (t != nlow[0] && t != nlow[1] && t != nlow[2]) p->item_tb[ni++] = t;
  bool n_used[6] = {true, true, true, false, false, false}, j_used[6] = {false, false, false, false, false, false};
  for (int i = 0; i < 3; i++) {
    const int t = p->item_tb[i];
    if (is_block(bits[t])) { p->j_of[t] = i; j_used[i] = true; }
    else { p->n_of[t] = 3 + i; n_used[3 + i] = true; }
  }
  for (int i = 3; i < 9; i++) {
    const int t = p->item_tb[i];
    bool* used = is_block(bits[t]) ? j_used : n_used;
    int k = 0;
    while (k < 6 && used[k]) k++;
    if (k == 6) return "internal: index bits exhausted.";
    (is_block(bits[t]) ? p->j_of[t] : p->n_of[t]) = k;
    used[k] = true;
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
  p->tile_mask = 0;
  for (int t = 0; t < kTileBits; t++) p->tile_mask |= 1ull << bits[t];
  p->ntiles = 1ull << (n_qubits - kTileBits);
  return nullptr;
}

}  // namespace tcb
