"""Generate golden vectors by running the REFERENCE's own CUDA implementation
(oracle/_ref/libprimitives_ref{32,64}.so == /root/reference/src/primitives.cu
compiled unmodified for sm_100a) on a B200.

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/ref_cuda_b200.npz'
    cp gpurun_out/ref_cuda_b200.npz tests/golden/

Inputs are seeded here; inputs AND the reference's outputs are stored so the
CPU suite (tests/test_oracle.py) can pin the oracle to the reference without a
GPU, and the GPU suite can compare the product to the very same vectors.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_replay as rr  # noqa: E402
from test_oracle import autodiff_gates, build_autodiff_circuit, tsallis_loss_and_cotangents  # noqa: E402

N_PRIM = 10
PAIRS = [(0, 1), (1, 0), (2, 7), (9, 0), (4, 9), (8, 3), (5, 6)]
POSITIONS = [0, 1, 4, 9]


def main(out_path):
    data = {}
    for precision, dtype in (("f32", np.complex64), ("f64", np.complex128)):
        lib = rr.RefLib(precision)
        rng = np.random.default_rng(2024)
        state = (rng.random(1 << N_PRIM) + 1j * rng.random(1 << N_PRIM)).astype(dtype)
        bwd = (rng.random(1 << N_PRIM) + 1j * rng.random(1 << N_PRIM)).astype(dtype)
        g1 = (rng.random(4) + 1j * rng.random(4)).astype(dtype)
        g2 = (rng.random(16) + 1j * rng.random(16)).astype(dtype)
        d = (rng.random(4) + 1j * rng.random(4)).astype(dtype)
        p = precision
        data.update({f"{p}/state": state, f"{p}/bwd": bwd, f"{p}/g1": g1, f"{p}/g2": g2, f"{p}/d": d})
        fb = rr.RefTensor.new_from_host(lib, bwd)
        for pos in POSITIONS:
            t = rr.RefTensor.new_from_host(lib, state); t.apply_q1_gate(g1, pos)
            data[f"{p}/q1gate/{pos}"] = t.get_cpu_state_copy()
            t = rr.RefTensor.new_from_host(lib, state); t.apply_q1_gate_inv(g1, pos)
            data[f"{p}/q1gate_inv/{pos}"] = t.get_cpu_state_copy()
            t = rr.RefTensor.new_from_host(lib, state)
            data[f"{p}/q1density/{pos}"] = t.get_q1_density(pos)
            data[f"{p}/q1grad/{pos}"] = rr.get_q1_grad(t, fb, pos)
        for p2, p1 in PAIRS:
            t = rr.RefTensor.new_from_host(lib, state); t.apply_q2_gate(g2, p2, p1)
            data[f"{p}/q2gate/{p2}_{p1}"] = t.get_cpu_state_copy()
            t = rr.RefTensor.new_from_host(lib, state); t.apply_q2_gate_inv(g2, p2, p1)
            data[f"{p}/q2gate_inv/{p2}_{p1}"] = t.get_cpu_state_copy()
            t = rr.RefTensor.new_from_host(lib, state); t.apply_q2_gate_diag(d, p2, p1)
            data[f"{p}/q2gate_diag/{p2}_{p1}"] = t.get_cpu_state_copy()
            t = rr.RefTensor.new_from_host(lib, state)
            data[f"{p}/q2density/{p2}_{p1}"] = t.get_q2_density(p2, p1)
            data[f"{p}/q2grad/{p2}_{p1}"] = rr.get_q2_grad(t, fb, p2, p1)
            data[f"{p}/q2grad_diag/{p2}_{p1}"] = rr.get_q2_grad_diag(t, fb, p2, p1)
        t = rr.RefTensor.new_from_host(lib, state)
        data[f"{p}/conj_and_double"] = t.conj_and_double().get_cpu_state_copy()
        # circuit-level: every instruction kind (src/test_autodiff.py pattern), n = 8, 2 layers
        n, layers = 8, 2
        rng = np.random.default_rng(42)
        const, var = autodiff_gates(rng, n, layers, dtype)
        c = rr.RefCircuit(n, precision)
        build_autodiff_circuit(c, n, layers)
        run_d = c.run(const, var)
        fwd_d = c.forward(const, var)
        _, cts = tsallis_loss_and_cotangents([x.astype(np.complex128) for x in fwd_d])
        cts = [ct.astype(dtype) for ct in cts]
        grads = c.backward([ct.conj() for ct in cts], const, var)
        data[f"{p}/circ/const"] = np.concatenate(const)
        data[f"{p}/circ/var"] = np.concatenate(var)
        data[f"{p}/circ/run"] = np.concatenate([x.reshape(-1) for x in run_d])
        data[f"{p}/circ/forward"] = np.concatenate([x.reshape(-1) for x in fwd_d])
        data[f"{p}/circ/cts"] = np.concatenate([x.reshape(-1) for x in cts])
        data[f"{p}/circ/grads"] = np.concatenate(grads)
        data[f"{p}/circ/final_state"] = c.get_cpu_state_copy()
    np.savez_compressed(out_path, **data)
    print("wrote", out_path, len(data), "arrays")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "ref_cuda_b200.npz"))
