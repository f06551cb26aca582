"""GPU: the product against the reference's own CUDA build -- live (both
libraries driven with identical inputs in one process) and through the
committed golden vectors."""
import os

import numpy as np
import pytest

from oracle import ref_replay as rr
from conftest import TOL, cmp_complex_slices, random_nonunitary, random_state_unnormalized
from test_oracle import GOLDEN, autodiff_gates, build_autodiff_circuit, tsallis_loss_and_cotangents

pytestmark = pytest.mark.gpu
DTYPES = [np.complex64, np.complex128]


def prec(dtype):
    return "f32" if np.dtype(dtype) == np.complex64 else "f64"


needs_ref = pytest.mark.skipif(not rr.ref_available("f32"), reason="oracle/_ref not built (make -C oracle)")


@needs_ref
@pytest.mark.parametrize("dtype", DTYPES)
def test_primitives_match_reference_library_live(pkg, dtype):
    rng = np.random.default_rng(5)
    n = 14
    lib = rr.RefLib(prec(dtype))
    tol = TOL[np.dtype(dtype)]
    state = random_state_unnormalized(rng, n, dtype)
    bwd = random_state_unnormalized(rng, n, dtype)
    for pos2, pos1 in [(0, 1), (1, 0), (13, 0), (0, 13), (6, 7), (12, 3), (2, 9)]:
        g2 = random_nonunitary(rng, 16, dtype); g1 = random_nonunitary(rng, 4, dtype)
        d = random_nonunitary(rng, 4, dtype)
        a, b = pkg.QuantizedTensor.new_from_host(state), rr.RefTensor.new_from_host(lib, state)
        ab, bb = pkg.QuantizedTensor.new_from_host(bwd), rr.RefTensor.new_from_host(lib, bwd)
        cmp_complex_slices(b.get_q2_density(pos2, pos1), a.get_q2_density(pos2, pos1), tol)
        cmp_complex_slices(b.get_q1_density(pos2), a.get_q1_density(pos2), tol)
        cmp_complex_slices(rr.get_q2_grad(b, bb, pos2, pos1), pkg.get_q2_grad(a, ab, pos2, pos1), tol)
        cmp_complex_slices(rr.get_q2_grad_diag(b, bb, pos2, pos1), pkg.get_q2_grad_diag(a, ab, pos2, pos1), tol)
        cmp_complex_slices(rr.get_q1_grad(b, bb, pos1), pkg.get_q1_grad(a, ab, pos1), tol)
        a.apply_q2_gate(g2, pos2, pos1); b.apply_q2_gate(g2, pos2, pos1)
        a.apply_q1_gate(g1, pos1); b.apply_q1_gate(g1, pos1)
        a.apply_q2_gate_diag(d, pos2, pos1); b.apply_q2_gate_diag(d, pos2, pos1)
        cmp_complex_slices(b.get_cpu_state_copy(), a.get_cpu_state_copy(), tol * 5)
        a.apply_q2_gate_inv(g2, pos2, pos1); b.apply_q2_gate_inv(g2, pos2, pos1)
        cmp_complex_slices(b.get_cpu_state_copy(), a.get_cpu_state_copy(), 1e-2 if dtype == np.complex64 else 1e-9)


@needs_ref
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("fuse", [0, 1, 2])
def test_circuit_matches_reference_replay_live(pkg, dtype, fuse):
    """BASELINE.json configs[1] shape: layered circuit with every instruction kind;
    densities and gradients vs the reference CUDA library driven by the replay of
    src/circuit.rs (<= 1e-5 f32, <= 1e-12 f64, relative to the largest entry)."""
    from quantum_differentiable_circuit import Circuit
    n, layers = 13, 2
    rng = np.random.default_rng(42)
    const, var = autodiff_gates(rng, n, layers, dtype)
    c = Circuit(n, precision=prec(dtype)); c.set_option("fuse", fuse); c.set_option("tile_bits", 11)
    r = rr.RefCircuit(n, prec(dtype))
    build_autodiff_circuit(c, n, layers); build_autodiff_circuit(r, n, layers)
    tol = TOL[np.dtype(dtype)]
    run_c, run_r = c.run(const, var), r.run(const, var)
    for a, b in zip(run_c, run_r):
        assert np.abs(a - b).max() < tol
    dens_c, dens_r = c.forward(const, var), r.forward(const, var)
    for a, b in zip(dens_c, dens_r):
        assert np.abs(a - b).max() < tol
    _, cts = tsallis_loss_and_cotangents([d.astype(np.complex128) for d in dens_r])
    cts = [ct.astype(dtype) for ct in cts]
    g_c = c.backward([ct.conj() for ct in cts], const, var)
    g_r = r.backward([ct.conj() for ct in cts], const, var)
    scale = max(np.abs(g).max() for g in g_r)
    assert max(np.abs(a - b).max() for a, b in zip(g_c, g_r)) / scale < tol


@pytest.mark.skipif(not os.path.exists(GOLDEN), reason="golden vectors not generated yet")
@pytest.mark.parametrize("dtype", DTYPES)
def test_product_matches_golden_vectors(pkg, dtype):
    from golden.replay_golden import _rel, _split, circuit_layout
    from quantum_differentiable_circuit import Circuit
    z = np.load(GOLDEN)
    p = prec(dtype)
    tol = 1e-5 if p == "f32" else 1e-12
    st, bw = z[f"{p}/state"], z[f"{p}/bwd"]
    g1, g2, d = z[f"{p}/g1"], z[f"{p}/g2"], z[f"{p}/d"]
    fb = pkg.QuantizedTensor.new_from_host(bw)
    for key in z.files:
        parts = key.split("/")
        if parts[0] != p or len(parts) != 3 or parts[1] == "circ":
            continue
        op, pos = parts[1], [int(x) for x in parts[2].split("_")]
        t = pkg.QuantizedTensor.new_from_host(st)
        if op == "q1gate": t.apply_q1_gate(g1, *pos); got = t.get_cpu_state_copy()
        elif op == "q1gate_inv": t.apply_q1_gate_inv(g1, *pos); got = t.get_cpu_state_copy()
        elif op == "q1density": got = t.get_q1_density(*pos)
        elif op == "q1grad": got = pkg.get_q1_grad(t, fb, *pos)
        elif op == "q2gate": t.apply_q2_gate(g2, *pos); got = t.get_cpu_state_copy()
        elif op == "q2gate_inv": t.apply_q2_gate_inv(g2, *pos); got = t.get_cpu_state_copy()
        elif op == "q2gate_diag": t.apply_q2_gate_diag(d, *pos); got = t.get_cpu_state_copy()
        elif op == "q2density": got = t.get_q2_density(*pos)
        elif op == "q2grad": got = pkg.get_q2_grad(t, fb, *pos)
        elif op == "q2grad_diag": got = pkg.get_q2_grad_diag(t, fb, *pos)
        else: raise KeyError(op)
        tt = 1e-2 if op.endswith("_inv") and p == "f32" else tol * (50 if op.endswith("_inv") else 1)
        assert _rel(got, z[key]) < tt, (key, _rel(got, z[key]))
    n, layers = 8, 2
    cs, vs = circuit_layout(n, layers)
    const, var = _split(z[f"{p}/circ/const"], cs), _split(z[f"{p}/circ/var"], vs)
    c = Circuit(n, precision=p)
    build_autodiff_circuit(c, n, layers)
    run = np.concatenate([x.reshape(-1) for x in c.run(const, var)])
    fwd = np.concatenate([x.reshape(-1) for x in c.forward(const, var)])
    assert np.abs(run - z[f"{p}/circ/run"]).max() < tol
    assert np.abs(fwd - z[f"{p}/circ/forward"]).max() < tol
    sizes = [4 if k == 13 else 16 for k in c._kinds if k in (12, 13)]
    cts = [x.reshape(2, 2) if x.size == 4 else x.reshape(4, 4) for x in _split(z[f"{p}/circ/cts"], sizes)]
    grads = np.concatenate(c.backward([x.conj() for x in cts], const, var))
    ref_g = z[f"{p}/circ/grads"]
    assert np.abs(grads - ref_g).max() / np.abs(ref_g).max() < tol
