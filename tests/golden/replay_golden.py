"""Replays tests/golden/ref_cuda_b200.npz (outputs of the reference CUDA build
on a B200, made by make_golden.py) against (a) the oracle on CPU and (b) the
product on GPU."""
import numpy as np

from oracle import statevector as sv
from oracle.circuit import OracleCircuit

N_PRIM = 10


def _rel(a, b):
    a, b = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
    m = np.maximum(np.abs(a), np.abs(b))
    nz = m != 0
    return float((np.abs(a[nz] - b[nz]) / m[nz]).max()) if nz.any() else 0.0


def _split(flat, sizes):
    out, o = [], 0
    for s in sizes:
        out.append(flat[o:o + s]); o += s
    assert o == flat.size
    return out


def circuit_layout(n, layers):
    from test_oracle import autodiff_const_layout, autodiff_var_layout
    return autodiff_const_layout(n, layers), autodiff_var_layout(n, layers)


def check_oracle_against_golden(path):
    from test_oracle import build_autodiff_circuit
    z = np.load(path)
    for p, tol in (("f32", 2e-5), ("f64", 1e-12)):
        st = z[f"{p}/state"].astype(np.complex128); bw = z[f"{p}/bwd"].astype(np.complex128)
        g1 = z[f"{p}/g1"].astype(np.complex128); g2 = z[f"{p}/g2"].astype(np.complex128)
        d = z[f"{p}/d"].astype(np.complex128)
        for key in z.files:
            parts = key.split("/")
            if parts[0] != p or len(parts) != 3 or parts[1] == "circ":
                continue
            op, where = parts[1], parts[2]
            pos = [int(x) for x in where.split("_")]
            ref = z[key]
            if op == "q1gate": got = sv.q1gate(st, g1, *pos)
            elif op == "q1gate_inv": got = sv.q1gate_inv(st, g1, *pos)
            elif op == "q1density": got = sv.q1density(st, *pos)
            elif op == "q1grad": got = sv.q1grad(st, bw, *pos)
            elif op == "q2gate": got = sv.q2gate(st, g2, *pos)
            elif op == "q2gate_inv": got = sv.q2gate_inv(st, g2, *pos)
            elif op == "q2gate_diag": got = sv.q2gate_diag(st, d, *pos)
            elif op == "q2density": got = sv.q2density(st, *pos)
            elif op == "q2grad": got = sv.q2grad(st, bw, *pos)
            elif op == "q2grad_diag": got = sv.q2grad_diag(st, bw, *pos)
            else: raise KeyError(op)
            t = 1e-2 if op.endswith("_inv") and p == "f32" else tol * (50 if op.endswith("_inv") else 1)
            assert _rel(got, ref) < t, (key, _rel(got, ref))
        assert _rel(sv.conj_and_double(st), z[f"{p}/conj_and_double"]) == 0.0
        # circuit level
        n, layers = 8, 2
        cs, vs = circuit_layout(n, layers)
        const = _split(z[f"{p}/circ/const"].astype(np.complex128), cs)
        var = _split(z[f"{p}/circ/var"].astype(np.complex128), vs)
        o = OracleCircuit(n)
        build_autodiff_circuit(o, n, layers)
        run = np.concatenate([x.reshape(-1) for x in o.run(const, var)])
        fwd = np.concatenate([x.reshape(-1) for x in o.forward(const, var)])
        ctol = tol * 20
        assert np.abs(run - z[f"{p}/circ/run"]).max() < ctol
        assert np.abs(fwd - z[f"{p}/circ/forward"]).max() < ctol
        nd = sum(1 for i in o.instructions if i[0] in (12, 13))
        sizes = [4 if i[0] == 13 else 16 for i in o.instructions if i[0] in (12, 13)]
        cts = [c.reshape(2, 2) if c.size == 4 else c.reshape(4, 4)
               for c in _split(z[f"{p}/circ/cts"].astype(np.complex128), sizes)]
        assert len(cts) == nd
        grads = np.concatenate(o.backward([c.conj() for c in cts], const, var))
        ref_g = z[f"{p}/circ/grads"]
        assert np.abs(grads - ref_g).max() / np.abs(ref_g).max() < ctol, np.abs(grads - ref_g).max() / np.abs(ref_g).max()
