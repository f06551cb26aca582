"""f32 gradient accuracy of the VQSE ansatz with and without host-side gate fusion, against the f64 build."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("differentiable-quantum-circuit-cuda_b200")
import bench
n, layers = 24, 26
res = {}
for name, cls, prec, dtype in (("f64", pkg.Circuit, "f64", np.complex128), ("f32", pkg.Circuit, "f32", np.complex64),
                               ("f32 fused", pkg.FusedCircuit, "f32", np.complex64),
                               ("f64 fused", pkg.FusedCircuit, "f64", np.complex128)):
    c = cls(n, precision=prec)
    bench.build_vqse(c, n, layers)
    var, cts = bench.vqse_inputs(n, layers, dtype)
    c.set_state_from_vector((np.ones(1 << n) / np.sqrt(float(1 << n))).astype(dtype))
    dens = c.forward([], var)
    grads = c.backward([x.conj() for x in cts], [], var)
    res[name] = (dens, grads)
ref_d, ref_g = res["f64"]
scale = max(np.abs(g).max() for g in ref_g)
for name in ("f32", "f32 fused", "f64 fused"):
    d, g = res[name]
    print(name, "density err %.2e" % max(np.abs(a - b).max() for a, b in zip(d, ref_d)),
          "gradient err / max|grad| %.2e" % (max(np.abs(a - b).max() for a, b in zip(g, ref_g)) / scale),
          "grad norm %.7f" % np.sqrt(sum(np.vdot(x, x).real for x in g)))
