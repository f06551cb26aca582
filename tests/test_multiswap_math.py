"""CPU: index arithmetic of the merged multi-qubit exchange (csrc/circuit.cuh: k_peer_multiswap, exchange_multi) re-executed
in NumPy on a small sharded register: k simultaneous (global qubit <-> local position) swaps, every pair of elements
handled by exactly one of its two owners, must equal the k single swaps carried out one after the other
(new[lpos = b, rank bit = c] = old[lpos = c, rank bit = b], Circuit::exchange)."""
import itertools

import numpy as np
import pytest


def ins0(i, pos):
    low = i & ((1 << pos) - 1)
    return ((i >> pos) << (pos + 1)) | low


def single_swap(shards, gbit, lpos):
    world = len(shards)
    new = [s.copy() for s in shards]
    idx = np.arange(shards[0].size)
    for r in range(world):
        c = (r >> gbit) & 1
        partner = r ^ (1 << gbit)
        sel = ((idx >> lpos) & 1) == (1 - c)            # my half lpos == 1 - c  <-  partner's half lpos == c
        src = (idx[sel] & ~(1 << lpos)) | (c << lpos)
        new[r][idx[sel]] = shards[partner][src]
    return new


def multiswap_as_the_kernel_does_it(shards, gbits, lposs, lv):
    """One 'launch' per rank, in place on all shards, with the kernel's work decomposition (vectors of 2^lv amplitudes)."""
    world, n_loc = len(shards), int(np.log2(shards[0].size))
    k = len(gbits)
    bufs = [s.reshape(-1, 1 << lv).copy() for s in shards]     # vector-indexed views
    touched = [np.zeros(b.shape[0], dtype=int) for b in bufs]
    half_log2 = n_loc - lv - k - 1
    per = 1 << half_log2
    for rank in range(world):
        c = sum(((rank >> gbits[i]) & 1) << i for i in range(k))
        pos_pair = [lp - lv for lp in lposs]
        pos_sorted = sorted(pos_pair)
        dep_c = sum(((c >> i) & 1) << pos_pair[i] for i in range(k))
        peer = {}
        for b in range(1 << k):
            if b == c:
                continue
            p = rank
            for i in range(k):
                p = (p & ~(1 << gbits[i])) | (((b >> i) & 1) << gbits[i])
            peer[b] = p
        for w in range(per * ((1 << k) - 1)):
            j = w >> half_log2
            b = c ^ (j + 1)
            x = (w & (per - 1)) + (0 if c < b else per)
            for i in range(k):
                x = ins0(x, pos_sorted[i])
            dep_b = sum(((b >> i) & 1) << pos_pair[i] for i in range(k))
            mi, pi = x | dep_b, x | dep_c
            va, vb = bufs[rank][mi].copy(), bufs[peer[b]][pi].copy()
            bufs[rank][mi] = vb
            bufs[peer[b]][pi] = va
            touched[rank][mi] += 1
            touched[peer[b]][pi] += 1
    return [b.reshape(-1) for b in bufs], touched


@pytest.mark.parametrize("lv", [0, 1])
@pytest.mark.parametrize("world,gbits,lposs", [
    (4, (0, 1), (3, 2)), (4, (1, 0), (2, 5)), (8, (0, 1, 2), (4, 5, 6)), (8, (2, 0, 1), (6, 1, 3)), (8, (0, 2), (5, 2)),
    (8, (1, 2), (1, 4)),
])
def test_merged_exchange_equals_the_single_swaps_in_sequence(world, gbits, lposs, lv):
    if any(lp < lv for lp in lposs):
        pytest.skip("positions below the vector width take the single-swap kernel")
    n_loc = 7
    rng = np.random.default_rng(1)
    shards = [rng.integers(0, 1 << 30, size=1 << n_loc) + (r << 40) for r in range(world)]
    want = shards
    for g, lp in zip(gbits, lposs):
        want = single_swap(want, g, lp)
    got, touched = multiswap_as_the_kernel_does_it(shards, gbits, lposs, lv)
    for r in range(world):
        np.testing.assert_array_equal(got[r], want[r])
        c = sum(((r >> gbits[i]) & 1) << i for i in range(len(gbits)))
        # every vector outside this rank's own 2^-k (local bits == c) moved exactly once
        vec = np.arange(touched[r].size)
        bits = sum((((vec >> (lp - lv)) & 1) << i) for i, lp in enumerate(lposs))
        np.testing.assert_array_equal(touched[r], (bits != c).astype(int))


def test_swaps_on_disjoint_pairs_commute():
    rng = np.random.default_rng(2)
    shards = [rng.integers(0, 1 << 30, size=1 << 6) + (r << 40) for r in range(8)]
    pairs = [(0, 3), (1, 4), (2, 5)]
    ref = None
    for order in itertools.permutations(pairs):
        cur = shards
        for g, lp in order:
            cur = single_swap(cur, g, lp)
        if ref is None:
            ref = cur
        for a, b in zip(cur, ref):
            np.testing.assert_array_equal(a, b)
