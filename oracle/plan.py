"""Interpreter for scheduler plans (the int64 encoding of `qdc_schedule`,
include/qdc_circuit.h) on top of the NumPy oracle.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  It lets the CPU suite check
the product's host-side scheduling logic without a GPU:

  * `run_plan_global`  executes a plan on ONE array holding all ranks' shards
    (physical layout: index bit p <-> physical position p, the top bits being
    the rank), checking on the way that every dense gate / density only touches
    local positions; forward densities and reverse-mode gradients must equal
    those of the program-order oracle VM (oracle.circuit.OracleCircuit).
  * `ShardedPlanRunner` executes the same plan with one NumPy shard per rank and
    an `exchange(buf, partner)` callback (torch.distributed/gloo in the tests),
    mirroring the half-shard swap of the CUDA executor.
"""
from __future__ import annotations

import numpy as np

from . import circuit as oc
from . import statevector as sv

ST_GATE, ST_DENS, ST_SWAP, ST_TILE = range(4)


def decode_plan(enc):
    enc = [int(x) for x in enc]
    steps, i = [], 0
    while enc[i] != -1:
        t, inst, p2, p1, gbit, lpos, count, nbits = enc[i:i + 8]
        i += 8
        st = {"type": t, "inst": inst, "p2": p2, "p1": p1, "gbit": gbit, "lpos": lpos}
        if t == ST_TILE:
            gates = []
            for _ in range(count):
                gates.append({"type": ST_GATE, "inst": enc[i], "p2": enc[i + 1], "p1": enc[i + 2]})
                i += 3
            st["gates"] = gates
            st["bits"] = enc[i:i + nbits]
            i += nbits
        steps.append(st)
    n = enc[i + 1]
    final_map = enc[i + 2:i + 2 + n]
    return steps, final_map


def flat_steps(steps):
    """Tile passes expanded to their gate steps, in execution order."""
    out = []
    for st in steps:
        if st["type"] == ST_TILE:
            bits = set(st["bits"])
            for g in st["gates"]:
                assert g["p2"] in bits and (g["p1"] < 0 or g["p1"] in bits), "gate outside its tile"
                out.append(g)
        else:
            out.append(st)
    return out


def swap_bits(state, a, b):
    n = sv.qubits_of(state)
    t = state.reshape([2] * n)
    return np.swapaxes(t, n - 1 - a, n - 1 - b).reshape(-1).copy()


def _apply_gate(state, kind, gate, p2, p1, inverse=False, transpose=False):
    if kind in oc._Q1_GATES:
        if inverse:
            g = sv.inverse(gate, 2) if kind in oc._NONU else sv.q1_conj_tr(gate)
        else:
            g = sv.q1_tr(gate) if transpose else gate
        return sv.q1gate(state, g, p2)
    if kind in oc._Q2_GATES:
        if inverse:
            g = sv.inverse(gate, 4) if kind in oc._NONU else sv.q2_conj_tr(gate)
        else:
            g = sv.q2_tr(gate) if transpose else gate
        return sv.q2gate(state, g, p2, p1)
    g = np.conj(gate) if inverse else gate
    return sv.q2gate_diag(state, g, p2, p1)


def bind(instructions, const_gates, var_gates):
    gates, ci, vi = {}, 0, 0
    for i, inst in enumerate(instructions):
        k = inst[0]
        if k >= oc.Q2_DENS:
            continue
        if k in oc._VAR:
            gates[i] = np.asarray(var_gates[vi]).reshape(-1); vi += 1
        else:
            gates[i] = np.asarray(const_gates[ci]).reshape(-1); ci += 1
    assert ci == len(const_gates) and vi == len(var_gates)
    return gates


def run_plan_global(steps, instructions, n, n_loc, const_gates, var_gates, cotangents_conj=None,
                    initial=None, dtype=np.complex128):
    """Forward (+ optional backward) of a plan on the global physical array.
    Returns (densities in program order, grads in program order or None, final state)."""
    gates = bind(instructions, const_gates, var_gates)
    state = sv.standard_state(n, dtype) if initial is None else np.asarray(initial, dtype=dtype).copy()
    fs = flat_steps(steps)
    dens = {}
    for st in fs:
        if st["type"] == ST_SWAP:
            state = swap_bits(state, n_loc + st["gbit"], st["lpos"])
            continue
        kind = instructions[st["inst"]][0]
        if st["type"] == ST_DENS:
            assert st["p2"] < n_loc and st["p1"] < n_loc, "density on a global position"
            dens[st["inst"]] = (sv.q1density(state, st["p2"]).reshape(2, 2) if st["p1"] < 0
                                else sv.q2density(state, st["p2"], st["p1"]).reshape(4, 4))
            continue
        if kind not in oc._DIAG_GATES:
            assert st["p2"] < n_loc and st["p1"] < n_loc, "dense gate on a global position"
        state = _apply_gate(state, kind, gates[st["inst"]], st["p2"], st["p1"])
    dens_list = [dens[i] for i in sorted(dens)]
    if cotangents_conj is None:
        return dens_list, None, state
    diff = [i for i, inst in enumerate(instructions) if inst[0] in (oc.DIFF_Q1_DENS, oc.DIFF_Q2_DENS)]
    ct = {i: np.asarray(c).reshape(-1) for i, c in zip(diff, cotangents_conj)}
    bwd, grads = None, {}
    for st in reversed(fs):
        if st["type"] == ST_SWAP:
            state = swap_bits(state, n_loc + st["gbit"], st["lpos"])
            if bwd is not None:
                bwd = swap_bits(bwd, n_loc + st["gbit"], st["lpos"])
            continue
        i = st["inst"]
        kind = instructions[i][0]
        if st["type"] == ST_DENS:
            if i not in ct:
                continue
            add = sv.conj_and_double(state)
            add = (sv.q1gate(add, sv.q1_tr(ct[i]), st["p2"]) if st["p1"] < 0
                   else sv.q2gate(add, sv.q2_tr(ct[i]), st["p2"], st["p1"]))
            bwd = add if bwd is None else bwd + add
            continue
        g = gates[i]
        state = _apply_gate(state, kind, g, st["p2"], st["p1"], inverse=True)
        if bwd is not None:
            if kind in oc._VAR:
                if kind in oc._Q1_GATES:
                    grads[i] = sv.q1grad(state, bwd, st["p2"])
                elif kind in oc._Q2_GATES:
                    grads[i] = sv.q2grad(state, bwd, st["p2"], st["p1"])
                else:
                    grads[i] = sv.q2grad_diag(state, bwd, st["p2"], st["p1"])
            bwd = _apply_gate(bwd, kind, g, st["p2"], st["p1"], transpose=True)
        elif kind in oc._VAR:
            grads[i] = np.zeros(g.size, dtype=dtype)
    return dens_list, [grads[i] for i in sorted(grads)], state


class ShardedPlanRunner:
    """One rank's view: local NumPy shard + an exchange callback, mirroring
    Circuit::exchange / effective_diag / scatter_diag_grad of the CUDA executor."""

    def __init__(self, n, rank, world, exchange, allreduce, dtype=np.complex128):
        self.n, self.rank, self.world = n, rank, world
        self.g = world.bit_length() - 1
        self.n_loc = n - self.g
        self.exchange, self.allreduce, self.dtype = exchange, allreduce, dtype

    def _swap(self, buf, gbit, lpos):
        c = (self.rank >> gbit) & 1
        partner = self.rank ^ (1 << gbit)
        v = buf.reshape(1 << (self.n_loc - lpos - 1), 2, 1 << lpos)
        out = np.ascontiguousarray(v[:, 1 - c, :])
        got = self.exchange(out, partner)
        v[:, 1 - c, :] = got
        return v.reshape(-1)

    def _rank_bit(self, p):
        return (self.rank >> (p - self.n_loc)) & 1

    def _eff_diag(self, d, p2, p1):
        g2, g1 = p2 >= self.n_loc, p1 >= self.n_loc
        if not g2 and not g1:
            return d, p2, p1
        e = np.zeros(4, dtype=d.dtype)
        if g2 and g1:
            e[0] = e[3] = d[2 * self._rank_bit(p2) + self._rank_bit(p1)]
            return e, 0, 0
        if g2:
            c2 = self._rank_bit(p2)
            e[0], e[3] = d[2 * c2], d[2 * c2 + 1]
            return e, p1, p1
        c1 = self._rank_bit(p1)
        e[0], e[3] = d[c1], d[2 + c1]
        return e, p2, p2

    @staticmethod
    def _diag_sel(state, d, s2, s1):
        idx = np.arange(state.size)
        j = 2 * ((idx >> s2) & 1) + ((idx >> s1) & 1)
        return j, d[j]

    def run(self, steps, instructions, const_gates, var_gates, cotangents_conj):
        gates = bind(instructions, const_gates, var_gates)
        state = np.zeros(1 << self.n_loc, dtype=self.dtype)
        if self.rank == 0:
            state[0] = 1
        fs = flat_steps(steps)
        dens = {}
        for st in fs:
            if st["type"] == ST_SWAP:
                state = self._swap(state, st["gbit"], st["lpos"])
                continue
            i = st["inst"]
            kind = instructions[i][0]
            if st["type"] == ST_DENS:
                dens[i] = (sv.q1density(state, st["p2"]) if st["p1"] < 0 else sv.q2density(state, st["p2"], st["p1"]))
                continue
            if kind in oc._DIAG_GATES:
                e, s2, s1 = self._eff_diag(gates[i], st["p2"], st["p1"])
                state = state * self._diag_sel(state, e, s2, s1)[1]
            else:
                state = _apply_gate(state, kind, gates[i], st["p2"], st["p1"])
        order = sorted(dens)
        flat = self.allreduce(np.concatenate([dens[i] for i in order])) if order else np.zeros(0)
        dens_list, o = [], 0
        for i in order:
            m = dens[i].size
            dens_list.append(flat[o:o + m].reshape((2, 2) if m == 4 else (4, 4))); o += m
        diff = [i for i, inst in enumerate(instructions) if inst[0] in (oc.DIFF_Q1_DENS, oc.DIFF_Q2_DENS)]
        ct = {i: np.asarray(c).reshape(-1) for i, c in zip(diff, cotangents_conj)}
        bwd, grads = None, {}
        for st in reversed(fs):
            if st["type"] == ST_SWAP:
                state = self._swap(state, st["gbit"], st["lpos"])
                if bwd is not None:
                    bwd = self._swap(bwd, st["gbit"], st["lpos"])
                continue
            i = st["inst"]
            kind = instructions[i][0]
            if st["type"] == ST_DENS:
                if i not in ct:
                    continue
                add = sv.conj_and_double(state)
                add = (sv.q1gate(add, sv.q1_tr(ct[i]), st["p2"]) if st["p1"] < 0
                       else sv.q2gate(add, sv.q2_tr(ct[i]), st["p2"], st["p1"]))
                bwd = add if bwd is None else bwd + add
                continue
            g = gates[i]
            if kind in oc._DIAG_GATES:
                e, s2, s1 = self._eff_diag(g, st["p2"], st["p1"])
                j, ev = self._diag_sel(state, e, s2, s1)
                state = state * np.conj(ev)
                if bwd is not None:
                    if kind in oc._VAR:
                        prod = bwd * state
                        h = np.array([prod[j == jj].sum() for jj in range(4)])
                        out = np.zeros(4, dtype=self.dtype)
                        g2, g1 = st["p2"] >= self.n_loc, st["p1"] >= self.n_loc
                        if not g2 and not g1:
                            out = h
                        elif g2 and g1:
                            out[2 * self._rank_bit(st["p2"]) + self._rank_bit(st["p1"])] = h[0] + h[3]
                        elif g2:
                            c2 = self._rank_bit(st["p2"]); out[2 * c2], out[2 * c2 + 1] = h[0], h[3]
                        else:
                            c1 = self._rank_bit(st["p1"]); out[c1], out[2 + c1] = h[0], h[3]
                        grads[i] = out
                    bwd = bwd * ev
                elif kind in oc._VAR:
                    grads[i] = np.zeros(4, dtype=self.dtype)
                continue
            state = _apply_gate(state, kind, g, st["p2"], st["p1"], inverse=True)
            if bwd is not None:
                if kind in oc._VAR:
                    grads[i] = (sv.q1grad(state, bwd, st["p2"]) if kind in oc._Q1_GATES
                                else sv.q2grad(state, bwd, st["p2"], st["p1"]))
                bwd = _apply_gate(bwd, kind, g, st["p2"], st["p1"], transpose=True)
            elif kind in oc._VAR:
                grads[i] = np.zeros(g.size, dtype=self.dtype)
        order = sorted(grads)
        flat = self.allreduce(np.concatenate([grads[i] for i in order])) if order else np.zeros(0)
        glist, o = [], 0
        for i in order:
            m = grads[i].size
            glist.append(flat[o:o + m]); o += m
        return dens_list, glist, state
