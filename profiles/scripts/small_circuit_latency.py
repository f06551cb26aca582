"""Wall-clock vs device time of one forward + backward on small registers (launch / host-bound regime):
the test_autodiff.py pattern (src/test_autodiff.py:49-118) at n qubits, 10 layers, this engine vs the
reference's CUDA library replaying circuit.rs."""
import importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
importlib.import_module("differentiable-quantum-circuit-cuda_b200")
import torch
from quantum_differentiable_circuit import Circuit
from test_oracle import autodiff_gates, build_autodiff_circuit
from oracle import ref_replay as rr

for n in (15, 20, 24):
    layers = 10
    rng = np.random.default_rng(42)
    const, var = autodiff_gates(rng, n, layers, np.complex64)
    for name in ("ours fuse=2", "ours fuse=0", "reference"):
        if name == "reference":
            if not rr.ref_available("f32", big=False):
                continue
            c = rr.RefCircuit(n, "f32")
        else:
            c = Circuit(n, precision="f32")
            c.set_option("fuse", 2 if "2" in name else 0)
        build_autodiff_circuit(c, n, layers)
        dens = c.forward(const, var)
        cts = [np.eye(d.shape[0], dtype=np.complex64) for d in dens]
        c.backward(cts, const, var)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            c.forward(const, var)
            c.backward(cts, const, var)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"n={n} {name:12s}: {dt*1e3:9.2f} ms per fwd+bwd ({len(var)} var + {len(const)} const gates, {len(dens)} diff densities)", flush=True)
