"""State checkpoint format (include/qdc_circuit.h: qdc_circuit_save_state / load_state;
SURVEY.md 8(f) item 4).  CPU: the pure-NumPy reader / writer / assembler.  GPU: the
library's streamed save and load against get_cpu_state_copy and the oracle."""
import importlib
import os

import numpy as np
import pytest

pkg = importlib.import_module("differentiable-quantum-circuit-cuda_b200")
sio = pkg.state_io


def _random_state(n, dtype, seed=3):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return (v / np.linalg.norm(v)).astype(dtype)


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
def test_single_file_round_trip(tmp_path, dtype):
    psi = _random_state(9, dtype)
    path = tmp_path / "s.qdc"
    sio.write_shard(path, psi, 9)
    assert os.path.getsize(path) == sio.HEADER_BYTES + psi.nbytes
    h, data = sio.read_shard(path)
    assert h == {"real_bytes": psi.dtype.itemsize // 2, "n": 9, "n_loc": 9, "rank": 0, "world": 1,
                 "map": list(range(9))}
    assert np.array_equal(np.asarray(data), psi)
    assert np.array_equal(sio.assemble([path]), psi)


def test_assemble_sharded_with_permuted_layout(tmp_path):
    """Shards written in a permuted physical layout re-assemble to the canonical vector."""
    n, world = 7, 4
    psi = _random_state(n, np.complex128)
    rng = np.random.default_rng(0)
    qmap = list(rng.permutation(n))          # logical q -> physical position
    # physical tensor: axis of position p is n-1-p and carries logical qubit q with qmap[q] == p
    inv = [qmap.index(p) for p in range(n)]  # physical position -> logical qubit
    t = psi.reshape([2] * n).transpose([n - 1 - inv[n - 1 - ax] for ax in range(n)])
    phys = np.ascontiguousarray(t).reshape(-1)
    shard = 1 << (n - 2)
    paths = []
    for r in range(world):
        p = tmp_path / f"s.rank{r}of{world}"
        sio.write_shard(p, phys[r * shard:(r + 1) * shard], n, rank=r, world=world, qubit_map=qmap)
        paths.append(p)
    assert np.array_equal(sio.assemble(paths[::-1]), psi)
    with pytest.raises(ValueError):
        sio.assemble(paths[:3])


def test_rejects_foreign_files(tmp_path):
    p = tmp_path / "junk"
    p.write_bytes(b"x" * 200)
    with pytest.raises(ValueError):
        sio.read_header(p)
    p.write_bytes(b"x" * 10)
    with pytest.raises(ValueError):
        sio.read_header(p)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_save_and_load_through_the_library(tmp_path, precision):
    from quantum_differentiable_circuit import Circuit
    from oracle import statevector as sv
    from conftest import haar_unitary
    n = 21   # 16 / 32 MiB: several 32 MiB staging chunks only for f64 -> also run a chunk-straddling size below
    dtype = np.complex64 if precision == "f32" else np.complex128
    psi = _random_state(n, dtype)
    rng = np.random.default_rng(1)
    g = haar_unitary(rng, 4).astype(dtype).reshape(-1)
    src = tmp_path / "init.qdc"
    sio.write_shard(src, psi, n)
    c = Circuit(n, precision=precision)
    c.add_q2_var_gate(7, 2)
    c.get_q2_dens_op(7, 2)
    c.load_state(str(src))
    dens = c.run([], [g])
    want = sv.q2gate(psi.astype(np.complex128), g.astype(np.complex128), 7, 2)
    tol = 1e-5 if precision == "f32" else 1e-12
    assert np.abs(c.get_cpu_state_copy() - want).max() < tol
    assert abs(np.trace(dens[0]) - 1) < 10 * tol
    dst = tmp_path / "out.qdc"
    c.save_state(str(dst))
    h, data = sio.read_shard(dst)
    assert h["n"] == n and h["map"] == list(range(n)) and c.state_layout() == list(range(n))
    assert np.array_equal(np.asarray(data), c.get_cpu_state_copy())
    # wrong precision / wrong size are refused with a message
    other = Circuit(n, precision="f64" if precision == "f32" else "f32")
    with pytest.raises(pkg.QdcError):
        other.load_state(str(dst))
    with pytest.raises(pkg.QdcError):
        Circuit(n - 1, precision=precision).load_state(str(dst))


@pytest.mark.gpu
def test_streamed_io_crosses_staging_chunks(tmp_path):
    """2^23 complex128 = 128 MiB = four 32 MiB staging chunks each way."""
    from quantum_differentiable_circuit import Circuit
    n = 23
    psi = _random_state(n, np.complex128, seed=9)
    src = tmp_path / "big.qdc"
    sio.write_shard(src, psi, n)
    c = Circuit(n, precision="f64")
    c.get_q1_dens_op(0)
    c.load_state(str(src))
    c.run([], [])
    dst = tmp_path / "big_out.qdc"
    c.save_state(str(dst))
    assert np.array_equal(np.asarray(sio.read_shard(dst)[1]), psi)
