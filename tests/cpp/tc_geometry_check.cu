// CPU check (host code only, compiled with nvcc because tc_block.cuh also holds the kernels) of the geometry of a
// tensor-core block pass, csrc/tc_block.cuh: make_params / ItemAddr / TileWalk, for random registers and block positions:
//  * the 12 tile positions hold the 5 lowest positions and the 6 block qubits; every tile bit is a block-index bit or a
//    rest-index bit, each used once;
//  * the 512 items x 8 elements of a fill / drain thread group address every amplitude of the tile exactly once, and
//    land on 64 x 64 distinct (block index, rest index) slots of an X slice;
//  * the tile bases never touch a tile position, are distinct, and the TileWalk recurrence
//    ((cur | F) + deposit(stride)) & ~F equals deposit(tile + stride).
//   nvcc -std=c++17 -I <csrc> tc_geometry_check.cu -o tc_geometry_check && ./tc_geometry_check
#include <cstdio>
#include <random>
#include <set>

#include "tc_block.cuh"

static std::mt19937_64 rng(11);

static int fail(const char* what, int n, const int* block) {
  printf("FAIL: %s (n = %d, block = %d,%d,%d,%d,%d,%d)\n", what, n, block[0], block[1], block[2], block[3], block[4], block[5]);
  return 1;
}

int main() {
  int cases = 0;
  for (int trial = 0; trial < 400; trial++) {
    const int n = 14 + (int)(rng() % 21);   // 14 .. 34
    int block[6];
    for (int b = 0; b < 6;) {
      const int q = trial % 4 == 0 ? (int)(rng() % 8) : (int)(rng() % n);   // every 4th trial: crowd the low positions
      bool dup = false;
      for (int c = 0; c < b; c++) dup |= block[c] == q;
      if (!dup && q < n) block[b++] = q;
    }
    tcb::Params p;
    int wbit[6];
    const char* err = tcb::make_params(block, n, &p, wbit);
    if (err) {
      // only legitimate refusals: block bits too far apart for 32-bit element offsets, or a fragmented complement
      continue;
    }
    cases++;
    // 1. tile positions
    uint64_t mask = 0;
    for (int t = 0; t < 12; t++) {
      if (t > 0 && p.pos[t] <= p.pos[t - 1]) return fail("tile positions not ascending", n, block);
      mask |= 1ull << p.pos[t];
    }
    if (mask != p.tile_mask) return fail("tile_mask", n, block);
    for (int q = 0; q < 5; q++)
      if (!((mask >> q) & 1)) return fail("low positions missing from the tile", n, block);
    for (int b = 0; b < 6; b++)
      if (!((mask >> block[b]) & 1)) return fail("block qubit missing from the tile", n, block);
    // 2. roles
    int jseen = 0, nseen = 0;
    for (int t = 0; t < 12; t++) {
      if ((p.j_of[t] >= 0) == (p.n_of[t] >= 0)) return fail("tile bit with no / two roles", n, block);
      bool is_block = false;
      for (int b = 0; b < 6; b++) is_block |= block[b] == p.pos[t];
      if (is_block != (p.j_of[t] >= 0)) return fail("role does not match the block", n, block);
      if (p.j_of[t] >= 0) jseen |= 1 << p.j_of[t]; else nseen |= 1 << p.n_of[t];
    }
    if (jseen != 63 || nseen != 63) return fail("index bits not a permutation", n, block);
    int wseen = 0;
    for (int k = 0; k < 6; k++) wseen |= 1 << wbit[k];
    if (wseen != 63) return fail("w_bit_of_jbit not a permutation", n, block);
    for (int t = 0; t < 12; t++)
      if (p.j_of[t] >= 0 && block[wbit[p.j_of[t]]] != p.pos[t]) return fail("w_bit_of_jbit", n, block);
    // 3. + 4. items (mirror of ItemAddr::init and of the fill's element order)
    std::set<uint64_t> amps;
    std::set<int> slots;
    int n_elem_bits[3] = {-1, -1, -1};
    for (int t = 0; t < 12; t++)
      if (p.n_of[t] >= 0 && p.n_of[t] < 3) n_elem_bits[p.n_of[t]] = t;
    for (int k = 0; k < 3; k++)
      if (n_elem_bits[k] < 0) return fail("n0..n2 not among the tile bits", n, block);
    for (int e = 0; e < 8; e++) {
      long want = 0;
      for (int k = 0; k < 3; k++)
        if ((e >> k) & 1) want |= 1l << p.pos[n_elem_bits[k]];
      if (want != p.elem_off[e]) return fail("elem_off", n, block);
    }
    if (p.fast != (p.pos[n_elem_bits[0]] == 0 ? 1 : 0)) return fail("fast flag", n, block);
    for (int item = 0; item < 512; item++) {
      uint64_t g = 0;
      int j = 0, nn = 0;
      for (int b = 0; b < 9; b++)
        if ((item >> b) & 1) {
          const int tb = p.item_tb[b];
          g |= 1ull << p.pos[tb];
          if (p.j_of[tb] >= 0) j |= 1 << p.j_of[tb]; else nn |= 1 << p.n_of[tb];
        }
      if (nn & 7) return fail("an item bit is one of n0..n2", n, block);
      const uint32_t chunk = tcb::x_chunk_byte(j, nn >> 3);
      if (chunk >= (uint32_t)tcb::kSliceBytesX / 2 || (chunk & 15)) return fail("chunk offset", n, block);
      for (int e = 0; e < 8; e++) {
        amps.insert(g + (uint64_t)p.elem_off[e]);
        slots.insert((int)chunk * 8 / 16 + e);   // 16-byte chunk of 8 bf16: slot = chunk index * 8 + element
      }
    }
    if (amps.size() != 4096) return fail("items do not cover the tile once", n, block);
    for (uint64_t a : amps)
      if (a & ~mask) return fail("item offset outside the tile positions", n, block);
    if (slots.size() != 4096) return fail("shared-memory slots collide", n, block);
    // 5. tile bases and the walk
    if (p.ntiles != 1ull << (n - 12)) return fail("ntiles", n, block);
    std::set<uint64_t> bases;
    const uint64_t sample = p.ntiles < 4096 ? p.ntiles : 4096;
    for (uint64_t k = 0; k < sample; k++) {
      const uint64_t t = p.ntiles <= 4096 ? k : rng() % p.ntiles;
      const uint64_t base = p.tile(t);
      if (base & mask) return fail("tile base touches a tile position", n, block);
      if (base >> n) return fail("tile base outside the register", n, block);
      bases.insert(base);
      const uint64_t stride = 1 + rng() % 200;
      if (t + stride < p.ntiles) {
        const uint64_t walked = ((base | mask) + p.tile(stride)) & ~mask;
        if (walked != p.tile(t + stride)) return fail("TileWalk recurrence", n, block);
      }
    }
    if (p.ntiles <= 4096 && bases.size() != p.ntiles) return fail("tile bases collide", n, block);
  }
  printf("%d geometries checked\n", cases);
  if (cases < 200) { printf("FAIL: too few geometries accepted\n"); return 1; }
  printf("OK\n");
  return 0;
}
