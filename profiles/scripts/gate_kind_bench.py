"""tile_bwd / tile_fwd time of single-gate-kind circuits, pair-lane (soa=1) vs interleaved (soa=0) f32 kernels."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
importlib.import_module("differentiable-quantum-circuit-cuda_b200")
from quantum_differentiable_circuit import Circuit

n, layers = 26, 12
rng = np.random.default_rng(0)


def haar(k):
    z = rng.normal(size=(k, k)) + 1j * rng.normal(size=(k, k))
    return np.linalg.qr(z)[0].astype(np.complex64).reshape(-1)


for kind in ("q1", "diag", "q2", "q1_low", "q1_high"):
    for soa in (1, 0):
        c = Circuit(n, precision="f32")
        c.set_option("soa", soa); c.set_option("profile", 1)
        var = []
        for _ in range(layers):
            if kind == "q1":
                for i in range(n): c.add_q1_var_gate(i); var.append(haar(2))
            elif kind == "q1_low":
                for i in range(4): c.add_q1_var_gate(i); var.append(haar(2))
            elif kind == "q1_high":
                for i in range(4, 12): c.add_q1_var_gate(i); var.append(haar(2))
            elif kind == "diag":
                for i in range(n - 1): c.add_q2_var_gate_diag(i, i + 1); var.append(np.exp(1j * rng.normal(size=4)).astype(np.complex64))
            else:
                for i in range(_ % 2, n - 1, 2): c.add_q2_var_gate(i + 1, i); var.append(haar(4))
        c.get_q2_dens_op_with_grad(1, 0)
        ct = np.eye(4, dtype=np.complex64)
        for it in range(2):
            c.forward([], var); pf = c.last_profile()
            c.backward([ct], [], var); pb = c.last_profile()
        print(kind, "soa", soa, "gates", len(var), "tile_fwd %.1f ms (%d)" % (pf.get("tile_fwd", {}).get("ms", 0), pf.get("tile_fwd", {}).get("launches", 0)),
              "tile_bwd %.1f ms (%d)" % (pb.get("tile_bwd", {}).get("ms", 0), pb.get("tile_bwd", {}).get("launches", 0)), flush=True)
