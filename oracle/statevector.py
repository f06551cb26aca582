"""NumPy restatement of the arithmetic in /root/reference/src/primitives.cu.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Layout (reference: src/qdc/circuit.py:29-30, src/primitives.cu:104-105):
a state is a flat array of 2**n complex numbers; qubit k is bit k of the linear
index (qubit 0 is the innermost / least significant).  Every function below is
written as an einsum over the view ``(2**(n-hi-1), 2, 2**(hi-lo-1), 2, 2**lo)``
exactly like the CPU oracles the reference's own Rust tests use
(src/quantized_tensor.rs:287-398); those einsum strings are the index-convention
specification of the path.

All functions are out-of-place and dtype-preserving (complex64 or complex128).
"""
from __future__ import annotations

import numpy as np


def qubits_of(state: np.ndarray) -> int:
    """log2 of the state length (src/quantized_tensor.rs:44-52)."""
    size = state.shape[0]
    assert size & (size - 1) == 0 and size > 0, "State size is not a power of 2."
    return size.bit_length() - 1


def standard_state(n: int, dtype=np.complex128) -> np.ndarray:
    """|0...0> (src/primitives.cu:176-199, `set2standard`)."""
    psi = np.zeros(1 << n, dtype=dtype)
    psi[0] = 1
    return psi


def _view1(state, pos):
    n = qubits_of(state)
    return state.reshape(1 << (n - pos - 1), 2, 1 << pos)


def _view2(state, pos2, pos1):
    n = qubits_of(state)
    hi, lo = (pos2, pos1) if pos2 > pos1 else (pos1, pos2)
    return state.reshape(1 << (n - hi - 1), 2, 1 << (hi - lo - 1), 2, 1 << lo)


# ---------------------------------------------------------------- gates
def q1gate(state: np.ndarray, gate, pos: int) -> np.ndarray:
    """psi'[..p..] = sum_q g[2p+q] psi[..q..] at bit `pos`.

    Reference kernel: src/primitives.cu:513-532 (`_q1gate`); test oracle
    src/quantized_tensor.rs:287-294 (einsum "iqk,jq->ijk").
    """
    g = np.asarray(gate, dtype=state.dtype).reshape(2, 2)
    return np.einsum("iqk,jq->ijk", _view1(state, pos), g).reshape(-1)


def q2gate(state: np.ndarray, gate, pos2: int, pos1: int) -> np.ndarray:
    """psi'[q2,q1] = sum g[8 q2 + 4 q1 + 2 p2 + p1] psi[p2,p1].

    `pos2` carries the more significant gate index whichever physical bit is
    larger.  Reference kernel: src/primitives.cu:573-606 (`_q2gate`); test
    oracle src/quantized_tensor.rs:296-308.
    """
    assert pos2 != pos1
    g = np.asarray(gate, dtype=state.dtype).reshape(2, 2, 2, 2)
    v = _view2(state, pos2, pos1)
    if pos2 > pos1:
        out = np.einsum("iqkpm,jlqp->ijklm", v, g)
    else:
        out = np.einsum("iqkpm,ljpq->ijklm", v, g)
    return out.reshape(-1)


def q2gate_diag(state: np.ndarray, gate, pos2: int, pos1: int) -> np.ndarray:
    """psi[p2,p1] *= d[2 p2 + p1].

    Reference kernel: src/primitives.cu:649-672 (`_q2gate_diag`); test oracle
    src/quantized_tensor.rs:310-322.
    """
    assert pos2 != pos1
    d = np.asarray(gate, dtype=state.dtype).reshape(2, 2)
    v = _view2(state, pos2, pos1)
    if pos2 > pos1:
        out = np.einsum("ijklm,jl->ijklm", v, d)
    else:
        out = np.einsum("ijklm,lj->ijklm", v, d)
    return out.reshape(-1)


def inverse(gate, size: int) -> np.ndarray:
    """Dense inverse used for the NonU un-compute.

    The reference calls cuBLAS `cublas{C,Z}matinvBatched` (CUDA toolkit 12.9,
    not vendored) at src/primitives.cu:114-138; its published contract is the
    plain matrix inverse, restated here with LAPACK via NumPy in complex128.
    """
    g = np.asarray(gate).reshape(size, size)
    return np.linalg.inv(g.astype(np.complex128)).astype(g.dtype).reshape(-1)


def q1gate_inv(state, gate, pos):
    """src/primitives.cu:547-570 (`q1gate_inv`)."""
    return q1gate(state, inverse(gate, 2), pos)


def q2gate_inv(state, gate, pos2, pos1):
    """src/primitives.cu:622-646 (`q2gate_inv`)."""
    return q2gate(state, inverse(gate, 4), pos2, pos1)


# ------------------------------------------------------------ densities
def q1density(state: np.ndarray, pos: int) -> np.ndarray:
    """rho[2p+q] = sum psi[p] conj(psi[q]); flat length 4.

    Reference: src/primitives.cu:689-739; oracle src/quantized_tensor.rs:324-332.
    """
    v = _view1(state, pos)
    return np.einsum("iqj,ipj->qp", v, v.conj()).reshape(-1)


def q2density(state: np.ndarray, pos2: int, pos1: int) -> np.ndarray:
    """rho[8 p2 + 4 p1 + 2 q2 + q1] = sum psi[p2,p1] conj(psi[q2,q1]); length 16.

    Reference: src/primitives.cu:779-837; oracle src/quantized_tensor.rs:334-347.
    """
    assert pos2 != pos1
    v = _view2(state, pos2, pos1)
    if pos2 > pos1:
        out = np.einsum("iqkpm,irksm->qprs", v, v.conj())
    else:
        out = np.einsum("iqkpm,irksm->pqsr", v, v.conj())
    return out.reshape(-1)


# ------------------------------------------------------------ gradients
def q1grad(fwd: np.ndarray, bwd: np.ndarray, pos: int) -> np.ndarray:
    """G[2p+q] = sum bwd[p] fwd[q]  (no conjugation); length 4.

    Reference: src/primitives.cu:202-253; oracle src/quantized_tensor.rs:349-358.
    """
    return np.einsum("iqj,ipj->qp", _view1(bwd, pos), _view1(fwd, pos)).reshape(-1)


def q2grad(fwd: np.ndarray, bwd: np.ndarray, pos2: int, pos1: int) -> np.ndarray:
    """G[8 p2 + 4 p1 + 2 q2 + q1] = sum bwd[p2,p1] fwd[q2,q1]; length 16.

    Reference: src/primitives.cu:295-354; oracle src/quantized_tensor.rs:360-374.
    """
    assert pos2 != pos1
    b, f = _view2(bwd, pos2, pos1), _view2(fwd, pos2, pos1)
    if pos2 > pos1:
        out = np.einsum("iqkpm,irksm->qprs", b, f)
    else:
        out = np.einsum("iqkpm,irksm->pqsr", b, f)
    return out.reshape(-1)


def q2grad_diag(fwd: np.ndarray, bwd: np.ndarray, pos2: int, pos1: int) -> np.ndarray:
    """G[2p+q] = sum bwd[p,q] fwd[p,q] on the (p=bit pos2, q=bit pos1) lattice.

    Reference: src/primitives.cu:398-452; oracle src/quantized_tensor.rs:376-390.
    """
    assert pos2 != pos1
    b, f = _view2(bwd, pos2, pos1), _view2(fwd, pos2, pos1)
    if pos2 > pos1:
        out = np.einsum("iqkpm,iqkpm->qp", b, f)
    else:
        out = np.einsum("iqkpm,iqkpm->pq", b, f)
    return out.reshape(-1)


# ---------------------------------------------------------- elementwise
def conj_and_double(state: np.ndarray) -> np.ndarray:
    """dst = 2 conj(src)  (src/primitives.cu:904-929)."""
    return 2 * state.conj()


def add(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """dst + src  (src/primitives.cu:931-953)."""
    return dst + src


# ------------------------------------------- host-side gate transforms
def q1_tr(gate):
    """Transpose of a flat 2x2 (src/quantized_tensor.rs:110-114: swap(1,2))."""
    return np.asarray(gate).reshape(2, 2).T.reshape(-1).copy()


def q1_conj_tr(gate):
    """Conjugate transpose of a flat 2x2 (src/quantized_tensor.rs:115-119)."""
    return np.asarray(gate).reshape(2, 2).conj().T.reshape(-1).copy()


def q2_tr(gate):
    """Transpose of a flat 4x4 (src/quantized_tensor.rs:134-139: six swaps)."""
    return np.asarray(gate).reshape(4, 4).T.reshape(-1).copy()


def q2_conj_tr(gate):
    """Conjugate transpose of a flat 4x4 (src/quantized_tensor.rs:140-145)."""
    return np.asarray(gate).reshape(4, 4).conj().T.reshape(-1).copy()


# ------------------------------------------- faster forms (CPU baseline)
def q2gate_fast(state: np.ndarray, gate, pos2: int, pos1: int) -> np.ndarray:
    """Same result as `q2gate`, written as 16 scaled-slice accumulations.

    Used only for timing the CPU baseline (bench.py); tested against `q2gate`.
    """
    g = np.asarray(gate, dtype=state.dtype).reshape(2, 2, 2, 2)
    v = _view2(state, pos2, pos1)
    out = np.empty_like(v)
    swap = pos2 < pos1  # then the hi physical bit carries gate index 1
    for h in range(2):
        for l in range(2):
            acc = None
            for hp in range(2):
                for lp in range(2):
                    c = g[l, h, lp, hp] if swap else g[h, l, hp, lp]
                    term = c * v[:, hp, :, lp, :]
                    acc = term if acc is None else acc + term
            out[:, h, :, l, :] = acc
    return out.reshape(-1)
