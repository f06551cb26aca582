import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
importlib.import_module("differentiable-quantum-circuit-cuda_b200")
from quantum_differentiable_circuit import Circuit
import bench
for prec, dtype in (("f32", np.complex64), ("f64", np.complex128)):
    n, depth = 24, 40
    var, cts = bench.brickwork_inputs(n, depth, dtype)
    ref = None
    for name, opts in (("fuse0", {"fuse": 0}),
                       ("aos rb1 strat1", {"fuse": 2, "soa": 0, "rb_policy": 1, "tile_strategy": 1}),
                       ("aos rb1 strat0", {"fuse": 2, "soa": 0, "rb_policy": 1, "tile_strategy": 0}),
                       ("aos rb1 strat1 T11", {"fuse": 2, "soa": 0, "rb_policy": 1, "tile_strategy": 1, "tile_bits": 11}),
                       ("aos rb1 strat1 maxg4", {"fuse": 2, "soa": 0, "rb_policy": 1, "tile_strategy": 1, "max_tile_gates": 4}),
                       ("aos rb1 strat1 maxg2", {"fuse": 2, "soa": 0, "rb_policy": 1, "tile_strategy": 1, "max_tile_gates": 2})):
        c = Circuit(n, precision=prec)
        for k, v in opts.items():
            c.set_option(k, v)
        bench.build_brickwork(c, n, depth)
        dens = c.forward([], var)
        st = c.last_stats()
        if ref is None:
            ref = dens
        ed = max(np.abs(a - b).max() for a, b in zip(dens, ref))
        print(prec, name, "dens err %.2e" % ed, st["kernel_launches"], flush=True)
