"""CPU tests that PIN the oracle (oracle/) before anything is checked against it.

 1. einsum restatement (oracle.statevector) == literal re-execution of the
    reference kernels' index arithmetic (oracle.literal) for every position.
 2. The reference's own known answers: GHZ amplitudes / densities
    (src/primitives.cu:961-1033, src/quantized_tensor.rs:488-506,
    src/test_ghz.py:9-60) through the circuit VM restatement.
 3. The 8th-order finite-difference identity of src/test_autodiff.py:133-165
    at rel 1e-9 on a circuit with every instruction kind.
 4. Golden vectors produced by the reference's own CUDA build on a B200
    (tests/golden/, see tests/golden/README.md) when present.
"""
import itertools
import os

import numpy as np
import pytest

from oracle import literal as lit
from oracle import statevector as sv
from oracle.circuit import OracleCircuit, vjp
from oracle import circuit as oc

from conftest import haar_unitary


@pytest.mark.parametrize("n", [3, 5])
def test_einsum_matches_literal_index_arithmetic(n):
    rng = np.random.default_rng(n)
    psi = rng.random(1 << n) + 1j * rng.random(1 << n)
    bw = rng.random(1 << n) + 1j * rng.random(1 << n)
    g1 = rng.random(4) + 1j * rng.random(4)
    g2 = rng.random(16) + 1j * rng.random(16)
    d = rng.random(4) + 1j * rng.random(4)
    for p in range(n):
        np.testing.assert_allclose(sv.q1gate(psi, g1, p), lit.q1gate(psi, g1, p), rtol=1e-13)
        np.testing.assert_allclose(sv.q1density(psi, p), lit.q1density(psi, p), rtol=1e-13)
        np.testing.assert_allclose(sv.q1grad(psi, bw, p), lit.q1grad(psi, bw, p), rtol=1e-13)
    for p2, p1 in itertools.permutations(range(n), 2):
        np.testing.assert_allclose(sv.q2gate(psi, g2, p2, p1), lit.q2gate(psi, g2, p2, p1), rtol=1e-13)
        np.testing.assert_allclose(sv.q2gate_fast(psi, g2, p2, p1), lit.q2gate(psi, g2, p2, p1), rtol=1e-13)
        np.testing.assert_allclose(sv.q2gate_diag(psi, d, p2, p1), lit.q2gate_diag(psi, d, p2, p1), rtol=1e-13)
        np.testing.assert_allclose(sv.q2density(psi, p2, p1), lit.q2density(psi, p2, p1), rtol=1e-13)
        np.testing.assert_allclose(sv.q2grad(psi, bw, p2, p1), lit.q2grad(psi, bw, p2, p1), rtol=1e-13)
        np.testing.assert_allclose(sv.q2grad_diag(psi, bw, p2, p1), lit.q2grad_diag(psi, bw, p2, p1), rtol=1e-13)


def test_host_transforms_match_rust_swaps():
    rng = np.random.default_rng(0)
    g = rng.random(16) + 1j * rng.random(16)
    t = g.copy()
    for a, b in ((1, 4), (2, 8), (6, 9), (3, 12), (7, 13), (11, 14)):  # src/quantized_tensor.rs:136-137
        t[a], t[b] = t[b], t[a]
    np.testing.assert_array_equal(sv.q2_tr(g), t)
    np.testing.assert_array_equal(sv.q2_conj_tr(g), t.conj())
    g1 = rng.random(4) + 1j * rng.random(4)
    t1 = g1.copy()
    t1[1], t1[2] = t1[2], t1[1]
    np.testing.assert_array_equal(sv.q1_tr(g1), t1)


def test_ghz_c_level_known_answer():
    """src/primitives.cu:961-1033 incl. the H . CZ_diag . H decomposition of the last CNOT."""
    n = 12
    s = 1 / np.sqrt(2)
    had = np.array([s, s, s, -s], dtype=np.complex128)
    cnot = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0], dtype=np.complex128)
    cz = np.array([1, 1, 1, -1], dtype=np.complex128)
    psi = sv.standard_state(n)
    psi = sv.q1gate(psi, had, 0)
    for i in range(n - 2):
        psi = sv.q2gate(psi, cnot, i, i + 1)
    psi = sv.q1gate(psi, had, n - 1)
    psi = sv.q2gate_diag(psi, cz, n - 2, n - 1)
    psi = sv.q1gate(psi, had, n - 1)
    expect = np.zeros(1 << n, dtype=np.complex128)
    expect[0] = expect[-1] = s
    np.testing.assert_allclose(psi, expect, atol=1e-12)
    for i in range(n):
        np.testing.assert_allclose(sv.q1density(psi, i), [0.5, 0, 0, 0.5], atol=1e-12)
    for i in range(n - 1):
        e = np.zeros(16)
        e[0] = e[15] = 0.5
        np.testing.assert_allclose(sv.q2density(psi, i, i + 1), e, atol=1e-12)


def build_ghz_python_circuit(c, n):
    """Instruction sequence of src/test_ghz.py:16-30."""
    c.add_q1_const_gate(0)
    for i in range(n - 1):
        c.get_q2_dens_op_with_grad(i, i + 1)
    for i in range(n):
        c.get_q1_dens_op_with_grad(i)
    for i in range(n - 1):
        c.add_q2_const_gate(i, i + 1)
    for i in range(n):
        c.get_q1_dens_op(i)
    for i in range(n - 1):
        c.get_q2_dens_op(i, i + 1)


def check_ghz_python_outputs(all_dm, diff_dm, n, atol):
    """Assertions of src/test_ghz.py:34-60."""
    assert len(all_dm) == 2 * n + 2 * (n - 1)
    assert len(diff_dm) == n + (n - 1)
    for lhs, rhs in zip(all_dm[: n + (n - 1)], diff_dm):
        np.testing.assert_allclose(lhs, rhs, atol=atol)
    s = 1 / np.sqrt(2)
    first_psi = np.tensordot(np.array([s, s]), np.array([1.0, 0.0]), axes=0).reshape(4)
    np.testing.assert_allclose(all_dm[0], np.outer(first_psi, first_psi.conj()), atol=atol)
    second = np.zeros((4, 4)); second[0, 0] = 1
    for d in all_dm[1:(n - 1)]:
        np.testing.assert_allclose(d, second, atol=atol)
    np.testing.assert_allclose(all_dm[n - 1], [[0.5, 0.5], [0.5, 0.5]], atol=atol)
    for d in all_dm[n:(2 * n - 1)]:
        np.testing.assert_allclose(d, [[1, 0], [0, 0]], atol=atol)
    for d in all_dm[(2 * n - 1):(3 * n - 1)]:
        np.testing.assert_allclose(d, [[0.5, 0], [0, 0.5]], atol=atol)
    tq = np.zeros((4, 4)); tq[0, 0] = tq[3, 3] = 0.5
    for d in all_dm[(3 * n - 1):]:
        np.testing.assert_allclose(d, tq, atol=atol)


def test_ghz_python_api_known_answer():
    n = 12  # BASELINE.json configs[0]; the file itself uses 21
    c = OracleCircuit(n)
    build_ghz_python_circuit(c, n)
    s = 1 / np.sqrt(2)
    had = np.array([s, s, s, -s], dtype=np.complex128)
    cnot = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0], dtype=np.complex128)
    gates = [had] + (n - 1) * [cnot]
    check_ghz_python_outputs(c.run(gates, []), c.forward(gates, []), n, 1e-12)


def build_autodiff_circuit(c, n, layers):
    """Instruction pattern of src/test_autodiff.py:51-81 (incl. the leaked loop
    variable `for i in range(i)` at :76-77)."""
    for _ in range(layers):
        for i in range(n):
            c.get_q1_dens_op_with_grad(i)
        for i in range(0, n - 1, 2):
            c.get_q2_dens_op_with_grad(i + 1, i)
        for i in range(n):
            c.add_q1_var_gate(i)
        for i in range(0, n - 1, 2):
            c.add_q2_var_gate(i + 1, i)
        for i in range(0, n - 1, 2):
            c.add_q2_var_gate_diag(i + 1, i)
        for i in range(n):
            c.add_q1_const_gate(i)
        for i in range(1, n - 1, 2):
            c.add_q2_const_gate(i + 1, i)
        for i in range(1, n - 1, 2):
            c.add_q2_const_gate_diag(i + 1, i)
        for i in range(n):
            c.add_q1_var_gate_nonu(i)
        for i in range(0, n - 1, 2):
            c.add_q2_var_gate_nonu(i + 1, i)
        for i in range(n):
            c.add_q1_const_gate_nonu(i)
        for i in range(1, n - 1, 2):
            c.add_q2_const_gate_nonu(i + 1, i)
        last = list(range(1, n - 1, 2))[-1] if n > 2 else 0
        for i in range(last):
            c.get_q1_dens_op(i)
    for i in range(n):
        c.get_q1_dens_op(i)
    for i in range(0, n - 1, 2):
        c.get_q2_dens_op(i + 1, i)


def autodiff_gates(rng, n, layers, dtype=np.complex128):
    """Gate lists of src/test_autodiff.py:96-118 (NumPy RNG instead of JAX keys)."""
    cnot = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0], dtype=dtype)
    rc = lambda k: (rng.normal(size=k * k) + 1j * rng.normal(size=k * k))  # noqa: E731

    def one_list(pairs):
        # the file uses int((n-1)/2) for both lists, valid for its odd n = 15; for
        # even n the variable (even-offset) and constant (odd-offset) pair counts differ
        gl = []
        for _ in range(layers):
            gl += [haar_unitary(rng, 2) for _ in range(n)]
            gl += pairs * [cnot]
            gl += [np.exp(1j * rng.normal(size=4)) for _ in range(pairs)]
            gl += [0.01 * rc(2) + haar_unitary(rng, 2) for _ in range(n)]
            gl += [0.01 * rc(4) + haar_unitary(rng, 4) for _ in range(pairs)]
        return [np.asarray(g, dtype=dtype) for g in gl]

    return one_list(len(range(1, n - 1, 2))), one_list(len(range(0, n - 1, 2)))  # (const, var)


def autodiff_layout(n, layers, pairs):
    sizes = []
    for _ in range(layers):
        sizes += [4] * n + [16] * pairs + [4] * pairs + [4] * n + [16] * pairs
    return sizes


def autodiff_var_layout(n, layers):
    return autodiff_layout(n, layers, len(range(0, n - 1, 2)))


def autodiff_const_layout(n, layers):
    return autodiff_layout(n, layers, len(range(1, n - 1, 2)))


def tsallis_loss_and_cotangents(dens):
    """mean(1 - tr rho^2), src/test_autodiff.py:87-92, and its JAX cotangents."""
    N = len(dens)
    loss = sum((1 - np.einsum("ij,ji->", d, d)).real for d in dens) / N
    cts = [(-2.0 / N) * d.T for d in dens]
    return loss, cts


def test_backward_matches_8th_order_finite_difference():
    """src/test_autodiff.py:133-165 on the restated VM: rel 1e-9 (complex128)."""
    n, layers, eta = 7, 3, 1e-6
    rng = np.random.default_rng(42)
    c = OracleCircuit(n)
    build_autodiff_circuit(c, n, layers)
    const, var = autodiff_gates(rng, n, layers)
    assert [v.size for v in var] == autodiff_var_layout(n, layers)
    pert = [rng.normal(size=v.size) + 1j * rng.normal(size=v.size) for v in var]

    def loss_at(s):
        return tsallis_loss_and_cotangents(c.forward(const, [v + s * eta * p for v, p in zip(var, pert)]))[0]

    fd = (loss_at(-4) / 280 - loss_at(4) / 280 - 4 * loss_at(-3) / 105 + 4 * loss_at(3) / 105
          + loss_at(-2) / 5 - loss_at(2) / 5 - 4 * loss_at(-1) / 5 + 4 * loss_at(1) / 5) / eta
    dens = c.forward(const, var)
    _, cts = tsallis_loss_and_cotangents(dens)
    grads = vjp(c, var, const, cts)
    ds = sum((g @ p).real for g, p in zip(grads, pert))
    assert abs(ds - fd) / min(abs(ds), abs(fd)) < 1e-9
    # after backward the working state is the initial state again (every gate un-computed)
    np.testing.assert_allclose(c.state, c.initial_state, atol=1e-9)


def test_zero_gradient_for_gates_after_last_seed():
    """src/circuit.rs:327-332: var gates met before any seed get zeros but are un-computed."""
    n = 4
    rng = np.random.default_rng(1)
    c = OracleCircuit(n)
    c.add_q1_var_gate(0)
    c.get_q1_dens_op_with_grad(0)
    c.add_q2_var_gate(1, 0)
    c.add_q2_var_gate_diag(2, 3)
    var = [haar_unitary(rng, 2), haar_unitary(rng, 4), np.exp(1j * rng.normal(size=4))]
    dens = c.forward([], var)
    grads = c.backward([np.eye(2, dtype=np.complex128)], [], var)
    assert [g.size for g in grads] == [4, 16, 4]
    assert np.all(grads[1] == 0) and np.all(grads[2] == 0) and np.any(grads[0] != 0)
    assert len(dens) == 1


def test_count_mismatch_errors():
    c = OracleCircuit(3)
    c.add_q1_const_gate(0)
    with pytest.raises(ValueError, match="less than required"):
        c.run([], [])
    with pytest.raises(ValueError, match="more than required"):
        c.run([np.eye(2).reshape(-1)] * 2, [])


GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_cuda_b200.npz")


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="golden vectors from the reference CUDA build not generated yet")
def test_oracle_matches_reference_cuda_golden_vectors():
    """Outputs of /root/reference/src/primitives.cu (unmodified, sm_100a) run on a
    B200 by tests/golden/make_golden.py; the oracle must reproduce them."""
    from golden.replay_golden import check_oracle_against_golden
    check_oracle_against_golden(GOLDEN)


def test_bench_torch_cpu_port_matches_the_oracle():
    """bench.py's multi-core CPU baseline applies gates as batched 4x4 matmuls: same result as the oracle."""
    import importlib
    import torch
    from oracle import statevector as sv
    bench = importlib.import_module("bench")
    n = 7
    rng = np.random.default_rng(3)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    g = bench.haar(rng, 4)
    for lo in (0, 2, 5):
        want = sv.q2gate(psi, g, lo + 1, lo)
        got = bench._torch_apply_adjacent(torch.tensor(psi), torch.tensor(g.reshape(4, 4)), lo, n).numpy()
        np.testing.assert_allclose(got, want, atol=1e-12)
        grad = torch.einsum("apc,aqc->pq", torch.tensor(want).view(1 << (n - lo - 2), 4, 1 << lo),
                            torch.tensor(psi).view(1 << (n - lo - 2), 4, 1 << lo)).numpy().reshape(-1)
        np.testing.assert_allclose(grad, sv.q2grad(psi, want, lo + 1, lo), atol=1e-11)
