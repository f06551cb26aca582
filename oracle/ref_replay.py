"""ctypes replay of the reference's Rust host layer against the reference's own
CUDA library (oracle/_ref/libprimitives_ref*.so, built unmodified from
/root/reference/src/primitives.cu by oracle/Makefile).  Needs a GPU.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).

`RefTensor` issues exactly the calls `QuantizedTensor` issues
(src/quantized_tensor.rs:54-238) and `RefCircuit` exactly the sequence of
`Circuit::{run,forward,backward}` (src/circuit.rs:164-429) -- including the
transient `conj_and_double` buffer and the per-gradient cudaMalloc / blocking
D2H inside the library -- because that *is* the reference's behaviour; it is
what `bench.py --impl reference` times.

The unmodified library is valid for n <= 30 (int shifts);
`big=True` loads the shift-patched build for n = 31..32.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import deque

import numpy as np

from . import circuit as oc

_HERE = os.path.dirname(os.path.abspath(__file__))
_vp, _sz = C.c_void_p, C.c_size_t
_SIGS = {
    "set2standard": (None, [_vp, _sz]), "get_state": (_vp, [C.POINTER(_vp), _sz]), "drop_state": (_vp, [_vp]),
    "copy_to_host": (_vp, [_vp, _vp, _sz]), "set_from_host": (_vp, [_vp, _vp, _sz]),
    "q1gate": (_vp, [_vp, _vp, _sz, _sz]), "q1gate_inv": (_vp, [_vp, _vp, _sz, _sz]),
    "q2gate": (_vp, [_vp, _vp, _sz, _sz, _sz]), "q2gate_inv": (_vp, [_vp, _vp, _sz, _sz, _sz]),
    "q2gate_diag": (_vp, [_vp, _vp, _sz, _sz, _sz]),
    "get_q1density": (_vp, [_vp, _vp, _sz, _sz]), "get_q2density": (_vp, [_vp, _vp, _sz, _sz, _sz]),
    "q1grad": (_vp, [_vp, _vp, _vp, _sz, _sz]), "q2grad": (_vp, [_vp, _vp, _vp, _sz, _sz, _sz]),
    "q2grad_diag": (_vp, [_vp, _vp, _vp, _sz, _sz, _sz]),
    "conj_and_double": (None, [_vp, _vp, _sz]), "add": (None, [_vp, _vp, _sz]), "copy": (None, [_vp, _vp, _sz]),
}


def ref_lib_path(precision: str, big: bool = False) -> str:
    bits = "32" if precision == "f32" else "64"
    return os.path.join(_HERE, "_ref", f"libprimitives_ref{bits}{'_big' if big else ''}.so")


def ref_available(precision: str = "f32", big: bool = False) -> bool:
    return os.path.exists(ref_lib_path(precision, big))


class RefLib:
    def __init__(self, precision: str, big: bool = False, path: str | None = None):
        self.precision = precision
        self.cdtype = np.dtype(np.complex64 if precision == "f32" else np.complex128)
        self.path = path or ref_lib_path(precision, big)
        self.cdll = C.CDLL(self.path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        for name, (res, args) in _SIGS.items():
            fn = getattr(self.cdll, name)
            fn.restype, fn.argtypes = res, args

    def call(self, name, *args):
        fn = getattr(self.cdll, name)
        out = fn(*args)
        if fn.restype is _vp and out:  # cuda_panic, src/quantized_tensor.rs:37-42
            raise RuntimeError(C.string_at(out).decode(errors="replace"))

    def arr(self, a, size=None):
        a = np.ascontiguousarray(np.asarray(a, dtype=self.cdtype).reshape(-1))
        assert size is None or a.size == size, "Incorrect len of the gate's buffer."
        return a


class RefTensor:
    """src/quantized_tensor.rs:54-238, call for call."""

    def __init__(self, lib: RefLib, n: int, ptr=None):
        self.lib, self.n = lib, n
        if ptr is None:
            p = _vp()
            lib.call("get_state", C.byref(p), n)
            ptr = p
        self.ptr = ptr

    @classmethod
    def new_standard(cls, lib, n):
        t = cls(lib, n)
        lib.call("set2standard", t.ptr, n)
        return t

    @classmethod
    def new_from_host(cls, lib, state):
        state = lib.arr(state)
        n = state.size.bit_length() - 1
        t = cls(lib, n)
        lib.call("set_from_host", t.ptr, state.ctypes.data, n)
        return t

    def drop(self):
        if self.ptr is not None:
            self.lib.call("drop_state", self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.drop()
        except Exception:
            pass

    def clone(self):
        t = RefTensor(self.lib, self.n)
        self.lib.call("copy", self.ptr, t.ptr, self.n)
        return t

    def set_from_host(self, state):
        state = self.lib.arr(state, 1 << self.n)
        self.lib.call("set_from_host", self.ptr, state.ctypes.data, self.n)

    def conj_and_double(self):
        t = RefTensor(self.lib, self.n)
        self.lib.call("conj_and_double", self.ptr, t.ptr, self.n)
        return t

    def add(self, other):
        self.lib.call("add", other.ptr, self.ptr, self.n)
        other.drop()  # `other` is moved into add() and dropped, src/quantized_tensor.rs:87-90

    def get_cpu_state_copy(self):
        out = np.empty(1 << self.n, dtype=self.lib.cdtype)
        self.lib.call("copy_to_host", self.ptr, out.ctypes.data, self.n)
        return out

    def apply_q1_gate(self, g, pos):
        g = self.lib.arr(g, 4)
        self.lib.call("q1gate", self.ptr, g.ctypes.data, pos, self.n)

    def apply_q1_gate_inv(self, g, pos):
        g = self.lib.arr(g, 4)
        self.lib.call("q1gate_inv", self.ptr, g.ctypes.data, pos, self.n)

    def apply_q1_gate_tr(self, g, pos):
        self.apply_q1_gate(self.lib.arr(g, 4).reshape(2, 2).T, pos)

    def apply_q1_gate_conj_tr(self, g, pos):
        self.apply_q1_gate(self.lib.arr(g, 4).reshape(2, 2).conj().T, pos)

    def apply_q2_gate(self, g, pos2, pos1):
        g = self.lib.arr(g, 16)
        self.lib.call("q2gate", self.ptr, g.ctypes.data, pos2, pos1, self.n)

    def apply_q2_gate_inv(self, g, pos2, pos1):
        g = self.lib.arr(g, 16)
        self.lib.call("q2gate_inv", self.ptr, g.ctypes.data, pos2, pos1, self.n)

    def apply_q2_gate_tr(self, g, pos2, pos1):
        self.apply_q2_gate(self.lib.arr(g, 16).reshape(4, 4).T, pos2, pos1)

    def apply_q2_gate_conj_tr(self, g, pos2, pos1):
        self.apply_q2_gate(self.lib.arr(g, 16).reshape(4, 4).conj().T, pos2, pos1)

    def apply_q2_gate_diag(self, g, pos2, pos1):
        g = self.lib.arr(g, 4)
        self.lib.call("q2gate_diag", self.ptr, g.ctypes.data, pos2, pos1, self.n)

    def apply_q2_gate_diag_conj(self, g, pos2, pos1):
        self.apply_q2_gate_diag(self.lib.arr(g, 4).conj(), pos2, pos1)

    def get_q1_density(self, pos):
        out = np.zeros(4, dtype=self.lib.cdtype)
        self.lib.call("get_q1density", self.ptr, out.ctypes.data, pos, self.n)
        return out

    def get_q2_density(self, pos2, pos1):
        out = np.zeros(16, dtype=self.lib.cdtype)
        self.lib.call("get_q2density", self.ptr, out.ctypes.data, pos2, pos1, self.n)
        return out


def get_q1_grad(fwd, bwd, pos):
    out = np.zeros(4, dtype=fwd.lib.cdtype)
    fwd.lib.call("q1grad", fwd.ptr, bwd.ptr, out.ctypes.data, pos, bwd.n)
    return out


def get_q2_grad(fwd, bwd, pos2, pos1):
    out = np.zeros(16, dtype=fwd.lib.cdtype)
    fwd.lib.call("q2grad", fwd.ptr, bwd.ptr, out.ctypes.data, pos2, pos1, bwd.n)
    return out


def get_q2_grad_diag(fwd, bwd, pos2, pos1):
    out = np.zeros(4, dtype=fwd.lib.cdtype)
    fwd.lib.call("q2grad_diag", fwd.ptr, bwd.ptr, out.ctypes.data, pos2, pos1, bwd.n)
    return out


class RefCircuit(oc.OracleCircuit):
    """`Circuit` of src/circuit.rs driven against the reference CUDA library.

    Reuses the builder methods / instruction list of OracleCircuit; run /
    forward / backward below issue the library calls of src/circuit.rs:164-429.
    """

    def __init__(self, qubits_number, precision="f32", big=None):
        if big is None:
            big = qubits_number > 30
        self.lib = RefLib(precision, big=big)
        self.n = qubits_number
        self.dtype = self.lib.cdtype
        self.instructions = []
        self.state_t = RefTensor.new_standard(self.lib, qubits_number)  # src/circuit.rs:96
        self.initial_t = self.state_t.clone()                           # :100
        # Optional timing hook of bench.py --impl reference: called with "gates" / "densities" whenever the
        # replay moves from gate instructions to density instructions or back (the caller records an event).
        self.on_phase = None
        self._phase = None

    def _enter(self, phase):
        if self.on_phase is not None and phase != self._phase:
            self.on_phase(phase)
        self._phase = phase

    def set_state_from_vector(self, vector):
        self.initial_t.set_from_host(vector)

    def _sweep(self, const_gates, var_gates, all_densities):
        assert self.instructions, "The circuit is empty."
        out = []
        cq, vq = deque(const_gates), deque(var_gates)
        self._phase = None
        self._enter("gates")
        self.lib.call("copy", self.initial_t.ptr, self.state_t.ptr, self.n)  # data_transfer
        st = self.state_t
        for inst in self.instructions:
            k = inst[0]
            self._enter("gates" if k in oc._Q1_GATES or k in oc._Q2_GATES or k in oc._DIAG_GATES else "densities")
            if k in oc._Q1_GATES:
                st.apply_q1_gate((vq if k in oc._VAR else cq).popleft(), inst[1])
            elif k in oc._Q2_GATES:
                st.apply_q2_gate((vq if k in oc._VAR else cq).popleft(), inst[1], inst[2])
            elif k in oc._DIAG_GATES:
                st.apply_q2_gate_diag((vq if k in oc._VAR else cq).popleft(), inst[1], inst[2])
            elif k in (oc.Q1_DENS, oc.DIFF_Q1_DENS):
                if all_densities or k == oc.DIFF_Q1_DENS:
                    out.append(st.get_q1_density(inst[1]).reshape(2, 2))
            else:
                if all_densities or k == oc.DIFF_Q2_DENS:
                    out.append(st.get_q2_density(inst[1], inst[2]).reshape(4, 4))
        assert not cq and not vq
        return out

    def backward(self, grads_wrt_density, const_gates, var_gates):
        gd, cg, vg = list(grads_wrt_density), list(const_gates), list(var_gates)
        fwd, bwd = self.state_t, None
        grads = deque()
        z = lambda m: np.zeros(m, dtype=self.dtype)  # noqa: E731
        self._phase = None
        for inst in reversed(self.instructions):
            k = inst[0]
            if k in (oc.Q1_DENS, oc.Q2_DENS):
                continue
            self._enter("densities" if k in (oc.DIFF_Q1_DENS, oc.DIFF_Q2_DENS) else "gates")
            if k in (oc.DIFF_Q1_DENS, oc.DIFF_Q2_DENS):
                g = gd.pop()
                add = fwd.conj_and_double()
                if k == oc.DIFF_Q1_DENS:
                    add.apply_q1_gate_tr(g, inst[1])
                else:
                    add.apply_q2_gate_tr(g, inst[1], inst[2])
                if bwd is None:
                    bwd = add
                else:
                    bwd.add(add)
                continue
            var = k in oc._VAR
            gate = (vg if var else cg).pop()
            if k in oc._Q1_GATES:
                pos = inst[1]
                (fwd.apply_q1_gate_inv if k in oc._NONU else fwd.apply_q1_gate_conj_tr)(gate, pos)
                if bwd is not None:
                    if var:
                        grads.appendleft(get_q1_grad(fwd, bwd, pos))
                    bwd.apply_q1_gate_tr(gate, pos)
                elif var:
                    grads.appendleft(z(4))
            elif k in oc._Q2_GATES:
                p2, p1 = inst[1], inst[2]
                (fwd.apply_q2_gate_inv if k in oc._NONU else fwd.apply_q2_gate_conj_tr)(gate, p2, p1)
                if bwd is not None:
                    if var:
                        grads.appendleft(get_q2_grad(fwd, bwd, p2, p1))
                    bwd.apply_q2_gate_tr(gate, p2, p1)
                elif var:
                    grads.appendleft(z(16))
            else:
                p2, p1 = inst[1], inst[2]
                fwd.apply_q2_gate_diag_conj(gate, p2, p1)
                if bwd is not None:
                    if var:
                        grads.appendleft(get_q2_grad_diag(fwd, bwd, p2, p1))
                    bwd.apply_q2_gate_diag(gate, p2, p1)
                elif var:
                    grads.appendleft(z(4))
        if bwd is not None:
            bwd.drop()
        assert not cg and not vg and not gd
        return list(grads)

    def get_cpu_state_copy(self):
        return self.state_t.get_cpu_state_copy()
