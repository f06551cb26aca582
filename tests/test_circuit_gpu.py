"""GPU parity tests of the circuit path (Circuit.run / forward / backward and
qdc.AutoGradCircuit) against the oracle's restatement of src/circuit.rs, the
reference's own Python tests (src/test_ghz.py, src/test_autodiff.py) and the
VQSE workload (example_vqse_ising.py)."""
import numpy as np
import pytest

from oracle.circuit import OracleCircuit, vjp
from conftest import TOL, haar_unitary
from test_oracle import (autodiff_gates, build_autodiff_circuit, build_ghz_python_circuit,
                         check_ghz_python_outputs, tsallis_loss_and_cotangents)

pytestmark = pytest.mark.gpu

DTYPES = [np.complex64, np.complex128]


def prec(dtype):
    return "f32" if np.dtype(dtype) == np.complex64 else "f64"


def assert_close_list(got, ref, tol):
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        r = np.asarray(r)
        scale = max(np.abs(r).max(), 1e-30)
        assert np.abs(np.asarray(g) - r).max() / scale < tol, (np.abs(np.asarray(g) - r).max() / scale, tol)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [12, 21])
def test_ghz_python_api(pkg, dtype, n):
    """src/test_ghz.py:9-60 (n = 21 in the file, 12 in BASELINE.json configs[0])."""
    from qdc import AutoGradCircuit
    c = AutoGradCircuit(n, precision=prec(dtype))
    build_ghz_python_circuit(c, n)
    simple_run, autodiff_run = c.build()
    cnot = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0], dtype=dtype)
    had = (np.array([1, 1, 1, -1]) / np.sqrt(2)).astype(dtype)
    gates = [had] + (n - 1) * [cnot]
    all_dm = simple_run([], gates)
    diff_dm = autodiff_run([], gates)
    check_ghz_python_outputs(all_dm, diff_dm, n, 1e-6 if dtype == np.complex64 else 1e-13)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("fuse", [0, 1, 2])
def test_every_instruction_kind_matches_oracle(pkg, dtype, fuse):
    """Circuit of src/test_autodiff.py:51-81 (all 14 instruction kinds): densities
    of run/forward and all gate gradients vs the oracle VM."""
    from quantum_differentiable_circuit import Circuit
    n, layers = (13, 2) if fuse else (9, 2)
    rng = np.random.default_rng(42)
    c = Circuit(n, precision=prec(dtype))
    c.set_option("fuse", fuse)
    c.set_option("tile_bits", 11)      # several tiles per pass at this small n
    o = OracleCircuit(n)
    init = np.zeros(1 << n, dtype=dtype); init[0] = 1
    c.set_state_from_vector(init)
    build_autodiff_circuit(c, n, layers)
    build_autodiff_circuit(o, n, layers)
    const, var = autodiff_gates(rng, n, layers, dtype)
    tol = TOL[np.dtype(dtype)]  # north_star: 1e-5 (f32) / 1e-12 (f64) relative to the largest entry
    assert_close_list(c.run(const, var), o.run(const, var), tol)
    dens = c.forward(const, var)
    dens_o = o.forward(const, var)
    assert_close_list(dens, dens_o, tol)
    _, cts = tsallis_loss_and_cotangents(dens_o)
    grads_o = vjp(o, var, const, cts)
    grads = c.backward([np.asarray(ct, dtype=dtype).conj() for ct in cts], const, var)
    assert [g.size for g in grads] == [g.size for g in grads_o]
    gscale = max(np.abs(g).max() for g in grads_o)
    for g, go in zip(grads, grads_o):
        assert np.abs(g - go).max() / gscale < tol
    # the working state is back at the initial state (every gate un-computed)
    assert np.abs(c.get_cpu_state_copy() - init).max() < tol
    assert c.last_profile() == {}  # profiling is opt-in


@pytest.mark.parametrize("fuse", [0, 1, 2])
def test_autodiff_finite_difference_f64(pkg, fuse):
    """src/test_autodiff.py:12-165 as written: n = 15, 10 layers, complex128,
    8th-order central difference with eta = 1e-6, rel 1e-9."""
    from qdc import AutoGradCircuit
    n, layers, eta = 15, 10, 1e-6
    rng = np.random.default_rng(42)
    c = AutoGradCircuit(n, precision="f64")
    c.circuit.set_option("fuse", fuse)
    c.circuit.set_option("tile_bits", 11)
    init = np.zeros(1 << n, dtype=np.complex128); init[0] = 1
    c.set_state_from_vector(init)
    build_autodiff_circuit(c, n, layers)
    const, var = autodiff_gates(rng, n, layers)
    pert = [rng.normal(size=v.size) + 1j * rng.normal(size=v.size) for v in var]
    _, fwd_circ = c.build()

    def loss_at(s):
        return tsallis_loss_and_cotangents(fwd_circ([v + s * eta * p for v, p in zip(var, pert)], const))[0]

    fd = (loss_at(-4) / 280 - loss_at(4) / 280 - 4 * loss_at(-3) / 105 + 4 * loss_at(3) / 105
          + loss_at(-2) / 5 - loss_at(2) / 5 - 4 * loss_at(-1) / 5 + 4 * loss_at(1) / 5) / eta
    dens = fwd_circ(var, const)
    _, cts = tsallis_loss_and_cotangents(dens)
    grads = fwd_circ.vjp(var, const, cts)
    ds = sum((g @ p).real for g, p in zip(grads, pert))
    assert abs(ds - fd) / min(abs(ds), abs(fd)) < 1e-9


def test_torch_autograd_glue(pkg):
    """autodiff_run on torch tensors: d loss / d gate through torch.autograd equals the raw vjp."""
    import torch
    from qdc import AutoGradCircuit
    n = 6
    rng = np.random.default_rng(3)
    c = AutoGradCircuit(n, precision="f64")
    for i in range(n):
        c.add_q1_var_gate(i)
    for i in range(0, n - 1, 2):
        c.add_q2_var_gate(i + 1, i)
    for i in range(n):
        c.get_q1_dens_op_with_grad(i)
    var = [haar_unitary(rng, 2) for _ in range(n)] + [haar_unitary(rng, 4) for _ in range(n // 2)]
    _, run = c.build()
    tv = [torch.tensor(v, requires_grad=True) for v in var]
    dens = run(tv, [])
    loss = sum((1 - torch.einsum("ij,ji->", d, d)).real for d in dens) / len(dens)
    loss.backward()
    dens_np = run.forward(var, [])
    _, cts = tsallis_loss_and_cotangents(dens_np)
    grads = run.vjp(var, [], cts)
    for t, g in zip(tv, grads):
        np.testing.assert_allclose(t.grad.numpy(), g.conj(), rtol=1e-10, atol=1e-12)


def test_torch_glue_with_device_resident_parameters(pkg):
    """the variational parameters live on the GPU, the gates of example_vqse_ising.py:15-28
    are built from them with torch ops on the device, the whole gate list crosses to the host in ONE transfer
    per call, and d energy / d parameter through torch.autograd equals a central finite difference."""
    import torch
    from qdc import AutoGradCircuit
    n, layers = 8, 2
    c = AutoGradCircuit(n, precision="f64")
    build_vqse(c, n, layers)
    _, run = c.build()
    h = torch.tensor(tfim_h(np.complex128), device="cuda")

    def energy(params):
        gates = []
        for l in range(layers):
            g, b = params[2 * l], params[2 * l + 1]
            zz = torch.exp(1j * torch.stack([-g, g, g, -g])).to(torch.complex128)
            x = torch.stack([torch.cos(b), -1j * torch.sin(b), -1j * torch.sin(b), torch.cos(b)]).to(torch.complex128)
            gates += n * [zz] + n * [x]
        dens = run(gates, [])
        assert all(d.is_cuda for d in dens)
        return sum(torch.einsum("ij,ji->", d, h).real for d in dens)

    p0 = torch.tensor(np.random.default_rng(5).normal(size=2 * layers), device="cuda", requires_grad=True)
    e = energy(p0)
    e.backward()
    grad = p0.grad.detach().cpu().numpy()
    fd = np.zeros_like(grad)
    eps = 1e-5
    with torch.no_grad():
        for k in range(grad.size):
            d = torch.zeros_like(p0); d[k] = eps
            fd[k] = float(energy(p0 + d) - energy(p0 - d)) / (2 * eps)
    np.testing.assert_allclose(grad, fd, rtol=1e-6, atol=1e-8)


def build_vqse(c, n, layers):
    """example_vqse_ising.py:66-79"""
    for _ in range(layers):
        for i in range(n - 1):
            c.add_q2_var_gate_diag(i, i + 1)
        c.add_q2_var_gate_diag(0, n - 1)
        for i in range(n):
            c.add_q1_var_gate(i)
    for i in range(n - 1):
        c.get_q2_dens_op_with_grad(i, i + 1)
    c.get_q2_dens_op_with_grad(0, n - 1)


def vqse_gates(params, n, dtype):
    """example_vqse_ising.py:15-28, 42-49"""
    gates = []
    for i in range(0, len(params), 2):
        g, b = params[i], params[i + 1]
        zz = np.array([np.exp(-1j * g), np.exp(1j * g), np.exp(1j * g), np.exp(-1j * g)], dtype=dtype)
        x = np.array([np.cos(b), -1j * np.sin(b), -1j * np.sin(b), np.cos(b)], dtype=dtype)
        gates += n * [zz] + n * [x]
    return gates


def tfim_h(dtype, field=1.0):
    """example_vqse_ising.py:86-93"""
    sz = np.diag([1.0, -1.0]); sx = np.array([[0, 1.0], [1.0, 0]]); eye = np.eye(2)
    k = lambda a, b: np.tensordot(a, b, axes=0).transpose(0, 2, 1, 3).reshape(4, 4)  # noqa: E731
    return (-k(sz, sz) - 0.5 * field * (k(sx, eye) + k(eye, sx))).astype(dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("fuse", [0, 1, 2])
def test_vqse_step_matches_oracle(pkg, dtype, fuse):
    """One value-and-grad of the TFIM ansatz (example_vqse_ising.py:52-113) at n = 10."""
    from quantum_differentiable_circuit import Circuit
    n, layers = (13, 3) if fuse else (10, 3)
    rng = np.random.default_rng(42)
    params = rng.normal(size=2 * layers)
    c = Circuit(n, precision=prec(dtype)); o = OracleCircuit(n)
    c.set_option("fuse", fuse)
    c.set_option("tile_bits", 11)
    init = (np.ones(1 << n) / np.sqrt(1 << n)).astype(dtype)
    c.set_state_from_vector(init); o.set_state_from_vector(init)
    build_vqse(c, n, layers); build_vqse(o, n, layers)
    gates = vqse_gates(params, n, dtype)
    h = tfim_h(dtype)
    dens, dens_o = c.forward([], gates), o.forward([], gates)
    tol = TOL[np.dtype(dtype)]  # north_star: 1e-5 (f32) / 1e-12 (f64) relative to the largest entry
    assert_close_list(dens, dens_o, tol)
    e, e_o = sum(np.einsum("ij,ji", d, h).real for d in dens), sum(np.einsum("ij,ji", d, h).real for d in dens_o)
    assert abs(e - e_o) / abs(e_o) < tol
    cts = [h.T.copy() for _ in dens]            # d(sum tr(rho h))/d rho, JAX cotangent
    grads = c.backward([ct.conj() for ct in cts], [], gates)
    grads_o = vjp(o, gates, [], cts)
    gscale = max(np.abs(g).max() for g in grads_o)
    for g, go in zip(grads, grads_o):
        assert np.abs(g - go).max() / gscale < tol


@pytest.mark.parametrize("dtype", DTYPES)
def test_zero_gradient_before_first_seed_and_errors(pkg, dtype):
    """src/circuit.rs:327-332 (zeros for var gates after the last diff density)
    and the panic conditions of :171-209, :425-427."""
    from quantum_differentiable_circuit import Circuit
    rng = np.random.default_rng(1)
    c = Circuit(4, precision=prec(dtype))
    with pytest.raises(pkg.QdcError, match="The circuit is empty."):
        c.run([], [])
    c.add_q1_var_gate(0)
    c.get_q1_dens_op_with_grad(0)
    c.add_q2_var_gate(1, 0)
    c.add_q2_var_gate_diag(2, 3)
    var = [haar_unitary(rng, 2, dtype), haar_unitary(rng, 4, dtype), np.exp(1j * rng.normal(size=4)).astype(dtype)]
    dens = c.forward([], var)
    assert len(dens) == 1 and dens[0].shape == (2, 2)
    grads = c.backward([np.eye(2, dtype=dtype)], [], var)
    assert [g.size for g in grads] == [4, 16, 4]
    assert np.all(grads[1] == 0) and np.all(grads[2] == 0) and np.any(grads[0] != 0)
    with pytest.raises(pkg.QdcError, match="less than required"):
        c.forward([], var[:2])
    with pytest.raises(pkg.QdcError, match="more than required"):
        c.forward([var[0]], var)
    with pytest.raises(pkg.QdcError, match="Incorrect len"):
        c.forward([], [var[1], var[1], var[2]])
    with pytest.raises(TypeError):
        c.forward([], [v.astype(np.complex128 if dtype == np.complex64 else np.complex64) for v in var])
    with pytest.raises(pkg.QdcError, match="must be different"):
        c.add_q2_var_gate(1, 1)
    with pytest.raises(pkg.QdcError, match="out of the bound"):
        c.add_q1_var_gate(4)
    with pytest.raises(pkg.QdcError, match="does not match"):
        c.set_state_from_vector(np.zeros(8, dtype=dtype))


@pytest.mark.parametrize("fuse", [0, 1, 2])
def test_f32_matches_f64_on_autodiff_20q(pkg, fuse):
    """BASELINE.json configs[1]: 20-qubit layered circuit, f32 gradients vs the f64 build (<= 1e-5)."""
    from quantum_differentiable_circuit import Circuit
    n, layers = 20, 2
    rng = np.random.default_rng(7)
    const, var = autodiff_gates(rng, n, layers)
    res = {}
    for dtype in DTYPES:
        c = Circuit(n, precision=prec(dtype))
        c.set_option("fuse", fuse)
        build_autodiff_circuit(c, n, layers)
        cg, vg = [g.astype(dtype) for g in const], [g.astype(dtype) for g in var]
        dens = c.forward(cg, vg)
        _, cts = tsallis_loss_and_cotangents([d.astype(np.complex128) for d in dens])
        grads = c.backward([ct.conj().astype(dtype) for ct in cts], cg, vg)
        res[np.dtype(dtype)] = (dens, grads)
    d32, g32 = res[np.dtype(np.complex64)]
    d64, g64 = res[np.dtype(np.complex128)]
    assert_close_list(d32, d64, 1e-5)
    gscale = max(np.abs(g).max() for g in g64)
    assert max(np.abs(a - b).max() for a, b in zip(g32, g64)) / gscale < 1e-5


def test_large_index_ghz_31q(pkg):
    """n = 31 (f32, 16 GiB): beyond the reference's `1 << n` int limit
    (SURVEY.md App. B).  GHZ through the legacy symbols, checked by densities."""
    n = 31
    vm = pkg.QuantizedTensor.new_standard(n, precision="f32")
    vm.apply_q1_gate(pkg.common_gates.get_hadamard(), 0)
    cnot = pkg.common_gates.get_cnot()
    for i in range(n - 1):
        vm.apply_q2_gate(cnot, i, i + 1)
    e = np.zeros(16); e[0] = e[15] = 0.5
    for p2, p1 in [(0, n - 1), (n - 1, 0), (n - 2, n - 1), (15, 30)]:
        np.testing.assert_allclose(vm.get_q2_density(p2, p1), e, atol=1e-5)
    for p in (0, 16, n - 1):
        np.testing.assert_allclose(vm.get_q1_density(p), [0.5, 0, 0, 0.5], atol=1e-5)
    vm.drop()


@pytest.mark.parametrize("dtype", DTYPES)
def test_fused_brickwork_equals_per_gate_executor(pkg, dtype):
    """Tiled multi-gate passes vs one pass per gate on the benchmark's circuit
    family (brickwork, Haar gates, Hermitian cotangents) at n = 22, default tile
    geometry: identical densities and gradients up to rounding."""
    import importlib
    bench = importlib.import_module("bench")
    from quantum_differentiable_circuit import Circuit
    n, depth = 22, 12
    var, cts = bench.brickwork_inputs(n, depth, dtype)
    out = {}
    # (fuse, soa): soa = 0 selects the interleaved-layout f32 tile kernels, the f32 default being the
    # pair-lane kernels (tile_soa_kernels.cuh); the key 3 is fuse = 1 with soa = 0
    # keys 4-6: register-blocked forward forced on (rb_policy 1; the default picks it by gate mix) with the
    # window-grown and the first-fit tiling, and with tiny passes (tiles that do not use up their bit budget)
    variants = ((0, 0, 1, {}), (1, 1, 1, {}), (2, 2, 1, {}), (3, 1, 0, {}),
                (4, 2, 1, {"rb_policy": 1}), (5, 2, 0, {"rb_policy": 1, "tile_strategy": 0}),
                (6, 2, 1, {"rb_policy": 1, "max_tile_gates": 2}))
    for key, fuse, soa, extra in variants:
        c = Circuit(n, precision=prec(dtype))
        c.set_option("fuse", fuse)
        c.set_option("soa", soa)
        for k, v in extra.items():
            c.set_option(k, v)
        bench.build_brickwork(c, n, depth)
        dens = c.forward([], var)
        grads = c.backward([x.conj() for x in cts], [], var)
        out[key] = (dens, grads, c.last_stats()["hbm_passes"])
    tol = TOL[np.dtype(dtype)]  # north_star: 1e-5 (f32) / 1e-12 (f64) relative to the largest entry
    gscale = max(np.abs(g).max() for g in out[0][1])
    for fuse in (1, 2, 3, 4, 5, 6):
        assert_close_list(out[fuse][0], out[0][0], tol)
        assert max(np.abs(a - b).max() for a, b in zip(out[fuse][1], out[0][1])) / gscale < tol
        if fuse != 6:
            assert out[fuse][2] < out[0][2] / 3, "fusion must cut the number of HBM sweeps"


@pytest.mark.parametrize("dtype", DTYPES)
def test_batched_densities_and_seeds_match_oracle_and_per_density_sweeps(pkg, dtype):
    """all densities / seeds of one program point in shared tiled sweeps
    (tile_dens_kernels.cuh).  A program point with more densities than one tile holds (q1 and q2,
    both qubit orders, on bit 0, differentiable mixed with plain) between two gate layers."""
    from quantum_differentiable_circuit import Circuit
    n = 14
    rng = np.random.default_rng(11)

    def build(c):
        for i in range(0, n - 1, 2):
            c.add_q2_var_gate(i + 1, i)
        for i in range(n - 1):                     # 13 q2 densities, alternating qubit order
            (c.get_q2_dens_op_with_grad if i % 3 else c.get_q2_dens_op)(*((i, i + 1) if i % 2 else (i + 1, i)))
        for i in range(0, n, 3):                   # q1 densities incl. bit 0
            c.get_q1_dens_op_with_grad(i)
        c.get_q2_dens_op_with_grad(n - 1, 0)       # a far pair
        for i in range(1, n - 1, 2):
            c.add_q2_var_gate(i, i + 1)
        c.add_q1_var_gate(0)
        for i in range(0, n - 1, 2):               # a second program point
            c.get_q2_dens_op_with_grad(i, i + 1)

    o = OracleCircuit(n)
    build(o)
    var = []
    for kind, *_ in o.instructions:
        if kind == 1:
            var.append(haar_unitary(rng, 4, dtype).reshape(-1))
        elif kind == 8:
            var.append(haar_unitary(rng, 2, dtype).reshape(-1))
    dens_o = o.forward([], var)
    cts = []
    for d in dens_o:
        a = rng.normal(size=d.shape) + 1j * rng.normal(size=d.shape)
        cts.append(((a + a.conj().T) / 2).astype(dtype))
    grads_o = vjp(o, var, [], cts)
    run_o = o.run([], var)
    tol = TOL[np.dtype(dtype)]  # north_star: 1e-5 (f32) / 1e-12 (f64) relative to the largest entry
    gscale = max(np.abs(g).max() for g in grads_o)
    launches = {}
    for batch in (1, 0):
        c = Circuit(n, precision=prec(dtype))
        c.set_option("tile_bits", 11)
        c.set_option("batch_dens", batch)
        build(c)
        assert_close_list(c.run([], var), run_o, tol)
        dens = c.forward([], var)
        launches[batch] = c.last_stats()["hbm_passes"]
        assert_close_list(dens, dens_o, tol)
        grads = c.backward([ct.conj() for ct in cts], [], var)
        launches[batch] += c.last_stats()["hbm_passes"]
        assert max(np.abs(a - b).max() for a, b in zip(grads, grads_o)) / gscale < tol
    assert launches[1] < launches[0] - 20, launches


@pytest.mark.parametrize("dtype", DTYPES)
def test_vqse_default_tile_geometry_equals_per_gate_executor(pkg, dtype):
    """The VQSE ansatz (diagonal ZZ ring + X rotations, ring densities incl. the far pair (0, n-1)) at n = 21
    with the DEFAULT tile geometry (2^12 / 2^11 tiles): select-free diagonal paths, one-qubit gates, lane-mixing
    gates on qubit 0, batched densities / seeds -- against the per-instruction executor."""
    from quantum_differentiable_circuit import Circuit
    n, layers = 21, 2
    rng = np.random.default_rng(7)
    gates = vqse_gates(rng.normal(size=2 * layers), n, dtype)
    h = tfim_h(dtype)
    init = (np.ones(1 << n) / np.sqrt(1 << n)).astype(dtype)
    out = {}
    for fuse in (0, 1, 2):
        c = Circuit(n, precision=prec(dtype))
        c.set_option("fuse", fuse)
        c.set_state_from_vector(init)
        build_vqse(c, n, layers)
        dens = c.forward([], gates)
        grads = c.backward([h.T.copy().conj() for _ in dens], [], gates)
        out[fuse] = (dens, grads)
    tol = TOL[np.dtype(dtype)]  # north_star: 1e-5 (f32) / 1e-12 (f64) relative to the largest entry
    gscale = max(np.abs(g).max() for g in out[0][1])
    for fuse in (1, 2):
        assert_close_list(out[fuse][0], out[0][0], tol)
        assert max(np.abs(a - b).max() for a, b in zip(out[fuse][1], out[0][1])) / gscale < tol
