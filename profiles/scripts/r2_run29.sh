# round 2, GPU call 29: the reference arm in the final tree; small-register latency of the final library
cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err; echo "ref exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_ref_final.json"))
print({k:d.get(k) for k in ("impl","value","ms_per_step","unavailable")}, d.get("config",{}).get("same_config"))
PY
timeout 120 python - <<'PY' > gpurun_out/r2_small_circuit_latency.txt 2>&1
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
importlib.import_module("differentiable-quantum-circuit-cuda_b200")
import torch
from quantum_differentiable_circuit import Circuit
from test_oracle import autodiff_gates, build_autodiff_circuit
for n in (15, 20, 24):
    rng = np.random.default_rng(42)
    const, var = autodiff_gates(rng, n, 10, np.complex64)
    c = Circuit(n, precision="f32")
    build_autodiff_circuit(c, n, 10)
    dens = c.forward(const, var)
    cts = [np.eye(d.shape[0], dtype=np.complex64) for d in dens]
    c.backward(cts, const, var)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        c.forward(const, var); c.backward(cts, const, var)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    t1 = time.perf_counter()
    for _ in range(5):
        c._flatten(const, 1); c._flatten(var, 1); c._flatten(cts, 2); c._flatten(const, 1); c._flatten(var, 1)
    fl = (time.perf_counter() - t1) / 5
    print(f"n={n}: {dt*1e3:8.2f} ms per fwd+bwd ({len(var)} var + {len(const)} const gates, {len(dens)} diff densities); Python marshalling of the gate lists alone: {fl*1e3:.2f} ms")
PY
cat gpurun_out/r2_small_circuit_latency.txt
