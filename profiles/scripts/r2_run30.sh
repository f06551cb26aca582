# round 2, GPU call 30: 20 q autodiff step (BASELINE configs[1]) under different tile geometries
cd $GRAFT_REPO_ROOT
timeout 150 python - <<'PY' > gpurun_out/r2_small_circuit_options.txt 2>&1
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
importlib.import_module("differentiable-quantum-circuit-cuda_b200")
import torch
from quantum_differentiable_circuit import Circuit
from test_oracle import autodiff_gates, build_autodiff_circuit
for n in (18, 20, 22):
    rng = np.random.default_rng(42)
    const, var = autodiff_gates(rng, n, 10, np.complex64)
    for name, opts in (("default", ()), ("tile_bits=11", (("tile_bits", 11),)), ("tile_bits=11,low=3", (("tile_bits", 11), ("low_bits", 3))),
                       ("tile_bits=13", (("tile_bits", 13),)), ("fuse=1", (("fuse", 1),)), ("profile", (("profile", 1),))):
        c = Circuit(n, precision="f32")
        try:
            for k, v in opts:
                c.set_option(k, v)
            build_autodiff_circuit(c, n, 10)
            dens = c.forward(const, var)
            cts = [np.eye(d.shape[0], dtype=np.complex64) for d in dens]
            c.backward(cts, const, var)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                c.forward(const, var); sf = c.last_stats(); pf = c.last_profile()
                c.backward(cts, const, var); sb = c.last_stats(); pb = c.last_profile()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 5
            extra = ""
            if name == "profile":
                extra = " device ms by category fwd " + str({k: round(v["ms"], 2) for k, v in pf.items()}) + " bwd " + str({k: round(v["ms"], 2) for k, v in pb.items()})
            print(f"n={n} {name:20s}: {dt*1e3:8.2f} ms per fwd+bwd, launches {sf['kernel_launches']} + {sb['kernel_launches']}{extra}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"n={n} {name}: {type(e).__name__}: {e}", flush=True)
        del c
PY
cat gpurun_out/r2_small_circuit_options.txt
