"""CPU: host mathematics of the tensor-core fused blocks (csrc/tc_host.hpp) -- the block matrix of a window and
the chain rule from the block gradient G_W = sum adjoint (x) state to the gradients of the member gates, checked
against the gate-by-gate reverse pass of src/circuit.rs:320-392 on the same vectors (tests/cpp/tc_host_check.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_chain_rule_equals_gate_by_gate_reverse_pass(tmp_path):
    exe = str(tmp_path / "tc_host_check")
    csrc = os.path.join(ROOT, "differentiable-quantum-circuit-cuda_b200", "csrc")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", csrc, os.path.join(ROOT, "tests", "cpp", "tc_host_check.cpp"),
                    "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr


def test_block_pass_geometry_covers_every_tile_once(tmp_path):
    """csrc/tc_block.cuh: make_params / ItemAddr / TileWalk for 400 random registers (14..34 qubits) and block positions
    (tests/cpp/tc_geometry_check.cu; host code only, compiled with nvcc because the header also holds the kernels)."""
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        import pytest
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "tc_geometry_check")
    csrc = os.path.join(ROOT, "differentiable-quantum-circuit-cuda_b200", "csrc")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-I", csrc,
                    os.path.join(ROOT, "tests", "cpp", "tc_geometry_check.cu"), "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
