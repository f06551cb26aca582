"""NumPy restatement of the circuit VM in /root/reference/src/circuit.rs.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`OracleCircuit` mirrors the PyO3 class `Circuit` (src/circuit.rs:86-430): the
same 14 instruction kinds, the same builder names, and `run` / `forward` /
`backward` with the same queue discipline (front pops going forward, back pops
going backward), the same panics (as AssertionError / ValueError) and the same
output ordering.  Arithmetic is delegated to oracle.statevector; nothing here
computes anything beyond transposes, conjugates and ordering -- exactly like
the Rust layer it follows.
"""
from __future__ import annotations

from collections import deque
from typing import List, Sequence

import numpy as np

from . import statevector as sv

# instruction kinds (src/circuit.rs:53-68)
CONST_Q2, VAR_Q2, CONST_Q2_NONU, VAR_Q2_NONU, CONST_Q2_DIAG, VAR_Q2_DIAG = range(6)
CONST_Q1, CONST_Q1_NONU, VAR_Q1, VAR_Q1_NONU = range(6, 10)
Q2_DENS, Q1_DENS, DIFF_Q2_DENS, DIFF_Q1_DENS = range(10, 14)

_Q1_GATES = (CONST_Q1, CONST_Q1_NONU, VAR_Q1, VAR_Q1_NONU)
_Q2_GATES = (CONST_Q2, VAR_Q2, CONST_Q2_NONU, VAR_Q2_NONU)
_DIAG_GATES = (CONST_Q2_DIAG, VAR_Q2_DIAG)
_VAR = (VAR_Q2, VAR_Q2_NONU, VAR_Q2_DIAG, VAR_Q1, VAR_Q1_NONU)
_NONU = (CONST_Q2_NONU, VAR_Q2_NONU, CONST_Q1_NONU, VAR_Q1_NONU)


class OracleCircuit:
    def __init__(self, qubits_number: int, dtype=np.complex128):
        # src/circuit.rs:95-103: state = |0..0>, initial_state = clone
        self.n = qubits_number
        self.dtype = np.dtype(dtype)
        self.instructions: List[tuple] = []
        self.initial_state = sv.standard_state(qubits_number, self.dtype)
        self.state = self.initial_state.copy()

    # ---- builders (src/circuit.rs:104-162) ----
    def set_state_from_vector(self, vector):
        vector = np.asarray(vector)
        assert sv.qubits_of(vector) == self.n, (
            "Size of the given state does not match the size of the tensor."
        )
        self.initial_state = vector.astype(self.dtype, copy=True)

    def add_q2_const_gate(self, pos2, pos1): self.instructions.append((CONST_Q2, pos2, pos1))
    def add_q2_const_gate_diag(self, pos2, pos1): self.instructions.append((CONST_Q2_DIAG, pos2, pos1))
    def add_q2_const_gate_nonu(self, pos2, pos1): self.instructions.append((CONST_Q2_NONU, pos2, pos1))
    def add_q2_var_gate(self, pos2, pos1): self.instructions.append((VAR_Q2, pos2, pos1))
    def add_q2_var_gate_diag(self, pos2, pos1): self.instructions.append((VAR_Q2_DIAG, pos2, pos1))
    def add_q2_var_gate_nonu(self, pos2, pos1): self.instructions.append((VAR_Q2_NONU, pos2, pos1))
    def add_q1_const_gate(self, pos): self.instructions.append((CONST_Q1, pos))
    def add_q1_const_gate_nonu(self, pos): self.instructions.append((CONST_Q1_NONU, pos))
    def add_q1_var_gate(self, pos): self.instructions.append((VAR_Q1, pos))
    def add_q1_var_gate_nonu(self, pos): self.instructions.append((VAR_Q1_NONU, pos))
    def get_q2_dens_op(self, pos2, pos1): self.instructions.append((Q2_DENS, pos2, pos1))
    def get_q1_dens_op(self, pos): self.instructions.append((Q1_DENS, pos))
    def get_q2_dens_op_with_grad(self, pos2, pos1): self.instructions.append((DIFF_Q2_DENS, pos2, pos1))
    def get_q1_dens_op_with_grad(self, pos): self.instructions.append((DIFF_Q1_DENS, pos))

    # ---- forward interpreters ----
    def _sweep(self, const_gates: Sequence, var_gates: Sequence, all_densities: bool):
        assert self.instructions, "The circuit is empty."
        out = []
        cq = deque(np.asarray(g, dtype=self.dtype).reshape(-1) for g in const_gates)
        vq = deque(np.asarray(g, dtype=self.dtype).reshape(-1) for g in var_gates)
        self.state = self.initial_state.copy()  # data_transfer, src/circuit.rs:174,225
        for inst in self.instructions:
            kind = inst[0]
            if kind in _Q1_GATES or kind in _Q2_GATES or kind in _DIAG_GATES:
                q = vq if kind in _VAR else cq
                if not q:
                    raise ValueError("The number of %s gates is less than required."
                                     % ("variable" if kind in _VAR else "constant"))
                gate = q.popleft()
                if kind in _Q1_GATES:
                    assert gate.size == 4, "Incorrect len of the gate's buffer."
                    self.state = sv.q1gate(self.state, gate, inst[1])
                elif kind in _Q2_GATES:
                    assert gate.size == 16, "Incorrect len of the gate's buffer."
                    self.state = sv.q2gate(self.state, gate, inst[1], inst[2])
                else:
                    assert gate.size == 4, "Incorrect len of the gate's buffer."
                    self.state = sv.q2gate_diag(self.state, gate, inst[1], inst[2])
            elif kind in (Q1_DENS, DIFF_Q1_DENS):
                if all_densities or kind == DIFF_Q1_DENS:
                    out.append(sv.q1density(self.state, inst[1]).reshape(2, 2))
            else:
                if all_densities or kind == DIFF_Q2_DENS:
                    out.append(sv.q2density(self.state, inst[1], inst[2]).reshape(4, 4))
        if cq:
            raise ValueError("Number of constant gates is more than required.")
        if vq:
            raise ValueError("Number of variable gates is more than required.")
        return out

    def run(self, const_gates, var_gates):
        """src/circuit.rs:164-212: every density instruction is evaluated."""
        return self._sweep(const_gates, var_gates, True)

    def forward(self, const_gates, var_gates):
        """src/circuit.rs:214-264: only the Diff* densities are evaluated."""
        return self._sweep(const_gates, var_gates, False)

    # ---- reverse interpreter (src/circuit.rs:266-429) ----
    def backward(self, grads_wrt_density, const_gates, var_gates):
        assert self.instructions, "The circuit is empty."
        gd = [np.asarray(g, dtype=self.dtype).reshape(-1) for g in grads_wrt_density]
        cg = [np.asarray(g, dtype=self.dtype).reshape(-1) for g in const_gates]
        vg = [np.asarray(g, dtype=self.dtype).reshape(-1) for g in var_gates]
        fwd = self.state
        bwd = None
        grads = deque()
        for inst in reversed(self.instructions):
            kind = inst[0]
            if kind in (Q1_DENS, Q2_DENS):
                continue
            if kind in (DIFF_Q1_DENS, DIFF_Q2_DENS):
                if not gd:
                    raise ValueError("The number of gradients wrt density matrices is less than required.")
                g = gd.pop()
                add = sv.conj_and_double(fwd)
                if kind == DIFF_Q1_DENS:
                    add = sv.q1gate(add, sv.q1_tr(g), inst[1])
                else:
                    add = sv.q2gate(add, sv.q2_tr(g), inst[1], inst[2])
                bwd = add if bwd is None else sv.add(add, bwd)
                continue
            is_var = kind in _VAR
            lst = vg if is_var else cg
            if not lst:
                raise ValueError("The number of gates is less than required.")
            gate = lst.pop()
            if kind in _Q1_GATES:
                pos = inst[1]
                if kind in _NONU:
                    fwd = sv.q1gate_inv(fwd, gate, pos)
                else:
                    fwd = sv.q1gate(fwd, sv.q1_conj_tr(gate), pos)
                if bwd is not None:
                    if is_var:
                        grads.appendleft(sv.q1grad(fwd, bwd, pos))
                    bwd = sv.q1gate(bwd, sv.q1_tr(gate), pos)
                elif is_var:
                    grads.appendleft(np.zeros(4, dtype=self.dtype))
            elif kind in _Q2_GATES:
                p2, p1 = inst[1], inst[2]
                if kind in _NONU:
                    fwd = sv.q2gate_inv(fwd, gate, p2, p1)
                else:
                    fwd = sv.q2gate(fwd, sv.q2_conj_tr(gate), p2, p1)
                if bwd is not None:
                    if is_var:
                        grads.appendleft(sv.q2grad(fwd, bwd, p2, p1))
                    bwd = sv.q2gate(bwd, sv.q2_tr(gate), p2, p1)
                elif is_var:
                    grads.appendleft(np.zeros(16, dtype=self.dtype))
            else:  # diagonal
                p2, p1 = inst[1], inst[2]
                fwd = sv.q2gate_diag(fwd, gate.conj(), p2, p1)
                if bwd is not None:
                    if is_var:
                        grads.appendleft(sv.q2grad_diag(fwd, bwd, p2, p1))
                    bwd = sv.q2gate_diag(bwd, gate, p2, p1)
                elif is_var:
                    grads.appendleft(np.zeros(4, dtype=self.dtype))
        self.state = fwd
        if cg:
            raise ValueError("Number of constant gates is more than required.")
        if vg:
            raise ValueError("Number of constant gates is more than required.")
        if gd:
            raise ValueError("Number of gradients wrt density matrices is more than required.")
        return [np.asarray(g, dtype=self.dtype) for g in grads]


def vjp(circuit, var_gates, const_gates, density_cotangents):
    """The `bwd_run` glue of src/qdc/circuit.py:190-197: conjugate the JAX
    cotangents, call backward, return per-var-gate flat gradients."""
    conj = [np.asarray(c).conj() for c in density_cotangents]
    return circuit.backward(conj, list(const_gates), list(var_gates))
