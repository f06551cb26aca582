"""GPU parity tests of the 18 legacy C symbols, written after the reference's
own Rust unit tests (src/quantized_tensor.rs:400-609): same construction
(random NON-unitary gates, UNNORMALISED random states), same element-wise
relative comparison (src/test_utils.rs:21-42), same tolerances -- but every
position / every ordered pair instead of 20 random pairs, both precisions, and
the oracle is oracle.statevector (complex128) instead of ndarray_einsum_beta.
Everything goes through the C ABI via the QuantizedTensor mirror.
"""
import itertools

import numpy as np
import pytest

from oracle import statevector as sv
from conftest import TOL, cmp_complex_slices, random_nonunitary, random_state_unnormalized

pytestmark = pytest.mark.gpu

DTYPES = [np.complex64, np.complex128]
N = 17  # qubits_number of the Rust tests


def _pairs(n, rng, count=None):
    pairs = list(itertools.permutations(range(n), 2))
    if count is None:
        return pairs
    idx = rng.choice(len(pairs), size=count, replace=False)
    return [pairs[i] for i in idx]


@pytest.mark.parametrize("dtype", DTYPES)
def test_q1gate(pkg, dtype):
    """src/quantized_tensor.rs:401-412"""
    rng = np.random.default_rng(1)
    state = random_state_unnormalized(rng, N, dtype)
    vm = pkg.QuantizedTensor.new_from_host(state)
    ref = state.astype(np.complex128)
    for pos in range(N):
        g = random_nonunitary(rng, 4, dtype)
        ref = sv.q1gate(ref, g.astype(np.complex128), pos)
        vm.apply_q1_gate(g, pos)
        cmp_complex_slices(ref, vm.get_cpu_state_copy(), TOL[np.dtype(dtype)] * 5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_q1gate_inv(pkg, dtype):
    """src/quantized_tensor.rs:415-427 (tolerance 1e-2 as in the reference)"""
    rng = np.random.default_rng(2)
    state = random_state_unnormalized(rng, N, dtype)
    vm = pkg.QuantizedTensor.new_from_host(state)
    for pos in range(N):
        g = random_nonunitary(rng, 4, dtype)
        vm.apply_q1_gate(g, pos)
        vm.apply_q1_gate_inv(g, pos)
        out = vm.get_cpu_state_copy()
        cmp_complex_slices(state, out, 1e-2)
        state = out


@pytest.mark.parametrize("dtype", DTYPES)
def test_q2gate_all_pairs(pkg, dtype):
    """src/quantized_tensor.rs:430-446, every ordered (pos2, pos1)."""
    rng = np.random.default_rng(3)
    n = 11
    for pos2, pos1 in _pairs(n, rng):
        state = random_state_unnormalized(rng, n, dtype)
        vm = pkg.QuantizedTensor.new_from_host(state)
        g = random_nonunitary(rng, 16, dtype)
        vm.apply_q2_gate(g, pos2, pos1)
        ref = sv.q2gate(state.astype(np.complex128), g.astype(np.complex128), pos2, pos1)
        cmp_complex_slices(ref, vm.get_cpu_state_copy(), TOL[np.dtype(dtype)])


@pytest.mark.parametrize("dtype", DTYPES)
def test_q2gate_chain(pkg, dtype):
    """src/quantized_tensor.rs:430-446 as written: 20 random pairs applied in sequence at n = 17."""
    rng = np.random.default_rng(4)
    state = random_state_unnormalized(rng, N, dtype)
    vm = pkg.QuantizedTensor.new_from_host(state)
    ref = state.astype(np.complex128)
    for pos2, pos1 in _pairs(N, rng, 20):
        g = random_nonunitary(rng, 16, dtype)
        ref = sv.q2gate(ref, g.astype(np.complex128), pos2, pos1)
        vm.apply_q2_gate(g, pos2, pos1)
        cmp_complex_slices(ref, vm.get_cpu_state_copy(), TOL[np.dtype(dtype)] * 5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_q2gate_inv(pkg, dtype):
    """src/quantized_tensor.rs:449-466"""
    rng = np.random.default_rng(5)
    state = random_state_unnormalized(rng, N, dtype)
    vm = pkg.QuantizedTensor.new_from_host(state)
    for pos2, pos1 in _pairs(N, rng, 20):
        g = random_nonunitary(rng, 16, dtype)
        vm.apply_q2_gate(g, pos2, pos1)
        vm.apply_q2_gate_inv(g, pos2, pos1)
        out = vm.get_cpu_state_copy()
        cmp_complex_slices(state, out, 1e-2)
        state = out


@pytest.mark.parametrize("dtype", DTYPES)
def test_q2gate_inv_matches_numpy_inverse(pkg, dtype):
    rng = np.random.default_rng(6)
    n = 10
    state = random_state_unnormalized(rng, n, dtype)
    for pos2, pos1 in _pairs(n, rng, 8):
        g = (np.eye(4).reshape(-1) + 0.3 * random_nonunitary(rng, 16, np.complex128)).astype(dtype)
        vm = pkg.QuantizedTensor.new_from_host(state)
        vm.apply_q2_gate_inv(g, pos2, pos1)
        ref = sv.q2gate_inv(state.astype(np.complex128), g.astype(np.complex128), pos2, pos1)
        cmp_complex_slices(ref, vm.get_cpu_state_copy(), TOL[np.dtype(dtype)] * 10)


@pytest.mark.parametrize("dtype", DTYPES)
def test_singular_inverse_reports_reference_error(pkg, dtype):
    """src/primitives.cu:128-132: "U(%d, %d) is zero." """
    vm = pkg.QuantizedTensor.new_standard(4, precision="f32" if dtype == np.complex64 else "f64")
    with pytest.raises(pkg.QdcError, match=r"U\(\d, \d\) is zero\."):
        vm.apply_q1_gate_inv(np.zeros(4, dtype=dtype), 0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_q2gate_diag(pkg, dtype):
    """src/quantized_tensor.rs:469-485, every ordered pair."""
    rng = np.random.default_rng(7)
    n = 11
    for pos2, pos1 in _pairs(n, rng):
        state = random_state_unnormalized(rng, n, dtype)
        vm = pkg.QuantizedTensor.new_from_host(state)
        d = random_nonunitary(rng, 4, dtype)
        vm.apply_q2_gate_diag(d, pos2, pos1)
        ref = sv.q2gate_diag(state.astype(np.complex128), d.astype(np.complex128), pos2, pos1)
        cmp_complex_slices(ref, vm.get_cpu_state_copy(), TOL[np.dtype(dtype)])


@pytest.mark.parametrize("dtype", DTYPES)
def test_ghz(pkg, dtype):
    """src/quantized_tensor.rs:488-506 (n = 21)"""
    n = 21
    vm = pkg.QuantizedTensor.new_standard(n, precision="f32" if dtype == np.complex64 else "f64")
    vm.apply_q1_gate(pkg.common_gates.get_hadamard(dtype), 0)
    cnot = pkg.common_gates.get_cnot(dtype)
    for i in range(n - 1):
        vm.apply_q2_gate(cnot, i, i + 1)
    state = vm.get_cpu_state_copy()
    s = 1 / np.sqrt(2)
    assert abs(state[0] - s) < 1e-5 and abs(state[-1] - s) < 1e-5
    assert np.all(np.abs(state[1:-1]) < 1e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_ghz_c_level_with_cz_decomposition(pkg, dtype):
    """src/primitives.cu:961-1033: last CNOT as H . q2gate_diag(CZ) . H; all densities."""
    n = 21
    vm = pkg.QuantizedTensor.new_standard(n, precision="f32" if dtype == np.complex64 else "f64")
    had, cnot = pkg.common_gates.get_hadamard(dtype), pkg.common_gates.get_cnot(dtype)
    cz = np.array([1, 1, 1, -1], dtype=dtype)
    vm.apply_q1_gate(had, 0)
    for i in range(n - 2):
        vm.apply_q2_gate(cnot, i, i + 1)
    vm.apply_q1_gate(had, n - 1)
    vm.apply_q2_gate_diag(cz, n - 2, n - 1)
    vm.apply_q1_gate(had, n - 1)
    state = vm.get_cpu_state_copy()
    s = 1 / np.sqrt(2)
    assert abs(state[0] - s) < 1e-5 and abs(state[-1] - s) < 1e-5
    assert np.all(np.abs(state[1:-1]) < 1e-5)
    for i in range(n):
        np.testing.assert_allclose(vm.get_q1_density(i), [0.5, 0, 0, 0.5], atol=1e-5)
    e = np.zeros(16); e[0] = e[15] = 0.5
    for i in range(n - 1):
        np.testing.assert_allclose(vm.get_q2_density(i, i + 1), e, atol=1e-5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_q1density(pkg, dtype):
    """src/quantized_tensor.rs:509-518"""
    rng = np.random.default_rng(8)
    state = random_state_unnormalized(rng, N, dtype)
    vm = pkg.QuantizedTensor.new_from_host(state)
    for i in range(N):
        cmp_complex_slices(sv.q1density(state.astype(np.complex128), i), vm.get_q1_density(i), TOL[np.dtype(dtype)])


@pytest.mark.parametrize("dtype", DTYPES)
def test_q2density(pkg, dtype):
    """src/quantized_tensor.rs:521-535, every ordered pair at n = 11 plus 20 random at n = 17."""
    rng = np.random.default_rng(9)
    for n, count in ((11, None), (N, 20)):
        state = random_state_unnormalized(rng, n, dtype)
        vm = pkg.QuantizedTensor.new_from_host(state)
        for pos2, pos1 in _pairs(n, rng, count):
            cmp_complex_slices(sv.q2density(state.astype(np.complex128), pos2, pos1),
                               vm.get_q2_density(pos2, pos1), TOL[np.dtype(dtype)])


@pytest.mark.parametrize("dtype", DTYPES)
def test_q1grad(pkg, dtype):
    """src/quantized_tensor.rs:538-549"""
    rng = np.random.default_rng(10)
    fwd = random_state_unnormalized(rng, N, dtype)
    bwd = random_state_unnormalized(rng, N, dtype)
    f, b = pkg.QuantizedTensor.new_from_host(fwd), pkg.QuantizedTensor.new_from_host(bwd)
    for pos in range(N):
        ref = sv.q1grad(fwd.astype(np.complex128), bwd.astype(np.complex128), pos)
        cmp_complex_slices(ref, pkg.get_q1_grad(f, b, pos), TOL[np.dtype(dtype)])


@pytest.mark.parametrize("dtype", DTYPES)
def test_q2grad(pkg, dtype):
    """src/quantized_tensor.rs:552-568"""
    rng = np.random.default_rng(11)
    for n, count in ((11, None), (N, 20)):
        fwd = random_state_unnormalized(rng, n, dtype)
        bwd = random_state_unnormalized(rng, n, dtype)
        f, b = pkg.QuantizedTensor.new_from_host(fwd), pkg.QuantizedTensor.new_from_host(bwd)
        for pos2, pos1 in _pairs(n, rng, count):
            ref = sv.q2grad(fwd.astype(np.complex128), bwd.astype(np.complex128), pos2, pos1)
            cmp_complex_slices(ref, pkg.get_q2_grad(f, b, pos2, pos1), TOL[np.dtype(dtype)])


@pytest.mark.parametrize("dtype", DTYPES)
def test_q2grad_diag(pkg, dtype):
    """src/quantized_tensor.rs:571-587"""
    rng = np.random.default_rng(12)
    for n, count in ((11, None), (N, 20)):
        fwd = random_state_unnormalized(rng, n, dtype)
        bwd = random_state_unnormalized(rng, n, dtype)
        f, b = pkg.QuantizedTensor.new_from_host(fwd), pkg.QuantizedTensor.new_from_host(bwd)
        for pos2, pos1 in _pairs(n, rng, count):
            ref = sv.q2grad_diag(fwd.astype(np.complex128), bwd.astype(np.complex128), pos2, pos1)
            cmp_complex_slices(ref, pkg.get_q2_grad_diag(f, b, pos2, pos1), TOL[np.dtype(dtype)])


@pytest.mark.parametrize("dtype", DTYPES)
def test_conj_half_add_clone(pkg, dtype):
    """src/quantized_tensor.rs:590-609 (+ Clone, :229-236)"""
    rng = np.random.default_rng(13)
    src = random_state_unnormalized(rng, N, dtype)
    dst = random_state_unnormalized(rng, N, dtype)
    vs, vd = pkg.QuantizedTensor.new_from_host(src), pkg.QuantizedTensor.new_from_host(dst)
    np.testing.assert_array_equal(vs.conj_and_double().get_cpu_state_copy(), 2 * src.conj())
    np.testing.assert_array_equal(vs.clone().get_cpu_state_copy(), src)
    vd.add(vs)
    np.testing.assert_array_equal(vd.get_cpu_state_copy(), dst + src)


@pytest.mark.parametrize("dtype", DTYPES)
def test_accumulate_semantics(pkg, dtype):
    """Result buffers are `+=` (src/primitives.cu:281-288, 765-772)."""
    rng = np.random.default_rng(14)
    n = 9
    state = random_state_unnormalized(rng, n, dtype)
    vm = pkg.QuantizedTensor.new_from_host(state)
    lib = pkg.get_lib("f32" if dtype == np.complex64 else "f64")
    out = np.full(16, 3 + 4j, dtype=dtype)
    lib.call("get_q2density", vm._ptr, out.ctypes.data, 5, 2, n)
    ref = sv.q2density(state.astype(np.complex128), 5, 2) + (3 + 4j)
    cmp_complex_slices(ref, out, TOL[np.dtype(dtype)])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1, 2, 3, 4])
def test_tiny_states(pkg, dtype, n):
    """Edge sizes: fewer work items than one warp / one vector."""
    rng = np.random.default_rng(15 + n)
    state = random_state_unnormalized(rng, n, dtype)
    vm = pkg.QuantizedTensor.new_from_host(state)
    ref = state.astype(np.complex128)
    for pos in range(n):
        g = random_nonunitary(rng, 4, dtype)
        vm.apply_q1_gate(g, pos)
        ref = sv.q1gate(ref, g.astype(np.complex128), pos)
    for pos2, pos1 in itertools.permutations(range(n), 2):
        g = random_nonunitary(rng, 16, dtype)
        d = random_nonunitary(rng, 4, dtype)
        vm.apply_q2_gate(g, pos2, pos1)
        vm.apply_q2_gate_diag(d, pos2, pos1)
        ref = sv.q2gate_diag(sv.q2gate(ref, g.astype(np.complex128), pos2, pos1), d.astype(np.complex128), pos2, pos1)
        ref = ref / np.abs(ref).max()
        vm.set_from_host(ref.astype(dtype))
        ref = ref.astype(dtype).astype(np.complex128)
        cmp_complex_slices(sv.q2density(ref, pos2, pos1), vm.get_q2_density(pos2, pos1), TOL[np.dtype(dtype)])
    cmp_complex_slices(ref, vm.get_cpu_state_copy(), TOL[np.dtype(dtype)] * 5)


@pytest.mark.parametrize("dtype", DTYPES)
def test_fused_reverse_step_equals_three_kernel_recipe(pkg, dtype):
    """qdc_reverse_step (one 4*S pass) == un-compute, gradient, adjoint of
    src/circuit.rs:320-333 / 348-363 / 380-392 done with the oracle."""
    import ctypes as C
    rng = np.random.default_rng(16)
    n = 12
    prec = "f32" if dtype == np.complex64 else "f64"
    lib = pkg.get_lib(prec)
    tol = TOL[np.dtype(dtype)]
    from conftest import haar_unitary
    for kind, klen, nonu in ((8, 4, 0), (9, 4, 1), (1, 16, 0), (3, 16, 1), (5, 4, 0)):
        for pos2, pos1 in [(0, 1), (1, 0), (3, 7), (11, 0), (5, 11), (10, 9)]:
            fwd = random_state_unnormalized(rng, n, dtype)
            bwd = random_state_unnormalized(rng, n, dtype)
            k = 2 if klen == 4 else 4
            if kind == 5:
                gate = np.exp(1j * rng.normal(size=4)).astype(dtype)
            else:
                gate = haar_unitary(rng, k)
                if nonu:
                    gate = gate + 0.01 * (rng.normal(size=k * k) + 1j * rng.normal(size=k * k))
                gate = gate.astype(dtype)
            f, b = pkg.QuantizedTensor.new_from_host(fwd), pkg.QuantizedTensor.new_from_host(bwd)
            grad = np.zeros(klen, dtype=dtype)
            lib.call("qdc_reverse_step", f._ptr, b._ptr, gate.ctypes.data, grad.ctypes.data, kind, nonu,
                     pos2, pos1, n)
            F, B, G = fwd.astype(np.complex128), bwd.astype(np.complex128), gate.astype(np.complex128)
            if kind in (8, 9):
                F2 = sv.q1gate_inv(F, G, pos2) if nonu else sv.q1gate(F, sv.q1_conj_tr(G), pos2)
                g_ref = sv.q1grad(F2, B, pos2)
                B2 = sv.q1gate(B, sv.q1_tr(G), pos2)
            elif kind in (1, 3):
                F2 = sv.q2gate_inv(F, G, pos2, pos1) if nonu else sv.q2gate(F, sv.q2_conj_tr(G), pos2, pos1)
                g_ref = sv.q2grad(F2, B, pos2, pos1)
                B2 = sv.q2gate(B, sv.q2_tr(G), pos2, pos1)
            else:
                F2 = sv.q2gate_diag(F, G.conj(), pos2, pos1)
                g_ref = sv.q2grad_diag(F2, B, pos2, pos1)
                B2 = sv.q2gate_diag(B, G, pos2, pos1)
            cmp_complex_slices(F2, f.get_cpu_state_copy(), tol * 5)
            cmp_complex_slices(B2, b.get_cpu_state_copy(), tol * 5)
            cmp_complex_slices(g_ref, grad, tol * 5)
