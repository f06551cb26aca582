import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PKG_NAME = "differentiable-quantum-circuit-cuda_b200"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (registers `qdc` and `quantum_differentiable_circuit`)."""
    return importlib.import_module(PKG_NAME)


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(autouse=True)
def _require_gpu_for_gpu_tests(request):
    if request.node.get_closest_marker("gpu") and not has_cuda():
        pytest.fail("gpu-marked test selected but no CUDA device is visible (no CPU fallback exists)")


# ---- shared helpers ------------------------------------------------------
def cmp_complex_slices(lhs, rhs, tol):
    """Element-wise relative comparison of the reference's Rust tests
    (src/test_utils.rs:21-42): |a-b| / max(|a|,|b|) < tol, skipped where both are 0."""
    lhs = np.asarray(lhs).reshape(-1)
    rhs = np.asarray(rhs).reshape(-1)
    assert lhs.shape == rhs.shape
    m = np.maximum(np.abs(lhs), np.abs(rhs))
    nz = m != 0
    rel = np.abs(lhs[nz] - rhs[nz]) / m[nz]
    bad = np.nonzero(rel >= tol)[0]
    assert bad.size == 0, f"{bad.size} elements differ by more than {tol}: worst {rel.max():.3e}"


def random_state_unnormalized(rng, n, dtype):
    """src/quantized_tensor.rs:280-285: entries U[0,1) + i U[0,1)."""
    return (rng.random(1 << n) + 1j * rng.random(1 << n)).astype(dtype)


def random_nonunitary(rng, size, dtype):
    """src/quantized_tensor.rs:256-278."""
    return (rng.random(size) + 1j * rng.random(size)).astype(dtype)


def haar_unitary(rng, k, dtype=np.complex128):
    z = rng.normal(size=(k, k)) + 1j * rng.normal(size=(k, k))
    q, r = np.linalg.qr(z)
    q = q * (np.diag(r) / np.abs(np.diag(r)))
    return q.reshape(-1).astype(dtype)


TOL = {np.dtype(np.complex64): 1e-5, np.dtype(np.complex128): 1e-12}
