"""Element-by-element re-execution of the reference kernels' index arithmetic.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pure-Python loops: small n
only.  Each function walks the same work items (`tid` over pairs / quads), forms
the same base index with the reference's INSERT_ZERO bit insertion
(src/primitives.cu:104-105) and reads the gate with the same flat index as the
kernel it cites.  tests/test_oracle.py checks oracle.statevector (the einsum
form) against these for every (pos2, pos1), which pins the index conventions to
the CUDA source rather than to my reading of the Rust test einsums.
"""
from __future__ import annotations

import numpy as np

_SIZE_MAX = (1 << 64) - 1


def insert_zero(mask: int, offset: int) -> int:
    """INSERT_ZERO(mask, offset) (src/primitives.cu:104-105)."""
    return (((mask & offset) << 1) | ((~mask & _SIZE_MAX) & offset)) & _SIZE_MAX


def _masks(pos2: int, pos1: int):
    # src/primitives.cu:306-311 / 581-586: max_mask = MIN(mask1, mask2) etc.
    mask1 = (_SIZE_MAX << pos1) & _SIZE_MAX
    mask2 = (_SIZE_MAX << pos2) & _SIZE_MAX
    return min(mask1, mask2), max(mask1, mask2)  # (max_mask, min_mask)


def _quad_base(tid: int, pos2: int, pos1: int) -> int:
    max_mask, min_mask = _masks(pos2, pos1)
    btid = insert_zero(min_mask, tid)
    return insert_zero(max_mask, btid)


def q1gate(state, gate, pos):
    """src/primitives.cu:513-532."""
    n = len(state).bit_length() - 1
    out = np.array(state, copy=True)
    mask = (_SIZE_MAX << pos) & _SIZE_MAX
    stride = 1 << pos
    for tid in range(1 << (n - 1)):
        btid = insert_zero(mask, tid)
        tmp = [0j, 0j]
        for q in range(2):
            for p in range(2):
                tmp[p] += gate[2 * p + q] * state[stride * q + btid]
        out[btid] = tmp[0]
        out[btid + stride] = tmp[1]
    return out


def q2gate(state, gate, pos2, pos1):
    """src/primitives.cu:573-606."""
    n = len(state).bit_length() - 1
    out = np.array(state, copy=True)
    s1, s2 = 1 << pos1, 1 << pos2
    for tid in range(1 << (n - 2)):
        btid = _quad_base(tid, pos2, pos1)
        tmp = [0j] * 4
        for q1 in range(2):
            for q2 in range(2):
                for p1 in range(2):
                    for p2 in range(2):
                        tmp[2 * q2 + q1] += (
                            gate[8 * q2 + 4 * q1 + 2 * p2 + p1]
                            * state[s2 * p2 + s1 * p1 + btid]
                        )
        out[btid] = tmp[0]
        out[btid + s1] = tmp[1]
        out[btid + s2] = tmp[2]
        out[btid + s1 + s2] = tmp[3]
    return out


def q2gate_diag(state, gate, pos2, pos1):
    """src/primitives.cu:649-672."""
    n = len(state).bit_length() - 1
    out = np.array(state, copy=True)
    s1, s2 = 1 << pos1, 1 << pos2
    for tid in range(1 << (n - 2)):
        btid = _quad_base(tid, pos2, pos1)
        out[btid] = gate[0] * state[btid]
        out[btid + s1] = gate[1] * state[btid + s1]
        out[btid + s2] = gate[2] * state[btid + s2]
        out[btid + s1 + s2] = gate[3] * state[btid + s1 + s2]
    return out


def q1density(state, pos):
    """src/primitives.cu:689-739."""
    n = len(state).bit_length() - 1
    mask = (_SIZE_MAX << pos) & _SIZE_MAX
    stride = 1 << pos
    rho = np.zeros(4, dtype=np.complex128)
    for tid in range(1 << (n - 1)):
        btid = insert_zero(mask, tid)
        for q in range(2):
            for p in range(2):
                rho[2 * p + q] += state[p * stride + btid] * np.conj(state[q * stride + btid])
    return rho


def q2density(state, pos2, pos1):
    """src/primitives.cu:779-837."""
    n = len(state).bit_length() - 1
    s1, s2 = 1 << pos1, 1 << pos2
    rho = np.zeros(16, dtype=np.complex128)
    for tid in range(1 << (n - 2)):
        btid = _quad_base(tid, pos2, pos1)
        for q1 in range(2):
            for q2 in range(2):
                for p1 in range(2):
                    for p2 in range(2):
                        rho[8 * p2 + 4 * p1 + 2 * q2 + q1] += state[
                            p2 * s2 + p1 * s1 + btid
                        ] * np.conj(state[q2 * s2 + q1 * s1 + btid])
    return rho


def q1grad(fwd, bwd, pos):
    """src/primitives.cu:202-253."""
    n = len(fwd).bit_length() - 1
    mask = (_SIZE_MAX << pos) & _SIZE_MAX
    stride = 1 << pos
    g = np.zeros(4, dtype=np.complex128)
    for tid in range(1 << (n - 1)):
        btid = insert_zero(mask, tid)
        for q in range(2):
            for p in range(2):
                g[2 * p + q] += bwd[p * stride + btid] * fwd[q * stride + btid]
    return g


def q2grad(fwd, bwd, pos2, pos1):
    """src/primitives.cu:295-354."""
    n = len(fwd).bit_length() - 1
    s1, s2 = 1 << pos1, 1 << pos2
    g = np.zeros(16, dtype=np.complex128)
    for tid in range(1 << (n - 2)):
        btid = _quad_base(tid, pos2, pos1)
        for q1 in range(2):
            for q2 in range(2):
                for p1 in range(2):
                    for p2 in range(2):
                        g[8 * p2 + 4 * p1 + 2 * q2 + q1] += (
                            bwd[p2 * s2 + p1 * s1 + btid] * fwd[q2 * s2 + q1 * s1 + btid]
                        )
    return g


def q2grad_diag(fwd, bwd, pos2, pos1):
    """src/primitives.cu:398-452."""
    n = len(fwd).bit_length() - 1
    s1, s2 = 1 << pos1, 1 << pos2
    g = np.zeros(4, dtype=np.complex128)
    for tid in range(1 << (n - 2)):
        btid = _quad_base(tid, pos2, pos1)
        for q in range(2):
            for p in range(2):
                g[2 * p + q] += bwd[p * s2 + q * s1 + btid] * fwd[p * s2 + q * s1 + btid]
    return g
