"""`QuantizedTensor`: RAII owner of a device state, Python mirror of the Rust
struct in /root/reference/src/quantized_tensor.rs:54-238.  Every method goes
through one of the 18 legacy C symbols (include/qdc_primitives.h), with the
same host-side gate transforms and the same assertions as the Rust original,
so parity tests read like src/quantized_tensor.rs:400-609.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._ffi import Lib, QdcError, default_precision, get_lib, precision_of


def _get_qubits_number(size: int) -> int:
    # src/quantized_tensor.rs:44-52
    if size <= 0 or size & (size - 1):
        raise QdcError("State size is not a power of 2.")
    return size.bit_length() - 1


class QuantizedTensor:
    def __init__(self, lib: Lib, ptr: int, qubits_number: int):
        self._lib = lib
        self._ptr = C.c_void_p(ptr)
        self.qubits_number = qubits_number

    # ---- constructors (src/quantized_tensor.rs:61-75) ----
    @classmethod
    def _alloc(cls, lib: Lib, n: int) -> "QuantizedTensor":
        p = C.c_void_p()
        lib.call("get_state", C.byref(p), n)
        return cls(lib, p.value, n)

    @classmethod
    def new_standard(cls, qubits_number: int, precision: str | None = None) -> "QuantizedTensor":
        lib = get_lib(precision or default_precision())
        t = cls._alloc(lib, qubits_number)
        lib.call("set2standard", t._ptr, qubits_number)
        return t

    @classmethod
    def new_from_host(cls, state: np.ndarray) -> "QuantizedTensor":
        state = np.ascontiguousarray(state)
        lib = get_lib(precision_of(state.dtype))
        n = _get_qubits_number(state.size)
        t = cls._alloc(lib, n)
        lib.call("set_from_host", t._ptr, state.ctypes.data, n)
        return t

    # ---- Drop (src/quantized_tensor.rs:223-227) ----
    def drop(self):
        if self._ptr is not None and self._ptr.value:
            self._lib.call("drop_state", self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.drop()
        except Exception:
            pass

    # ---- Clone (src/quantized_tensor.rs:229-236) ----
    def clone(self) -> "QuantizedTensor":
        t = self._alloc(self._lib, self.qubits_number)
        self._lib.call("copy", self._ptr, t._ptr, self.qubits_number)
        return t

    @property
    def device_ptr(self) -> int:
        return self._ptr.value

    @property
    def dtype(self):
        return self._lib.cdtype

    # ---- methods (src/quantized_tensor.rs:76-167) ----
    def set_from_host(self, state: np.ndarray):
        state = self._lib.host(state)
        n = _get_qubits_number(state.size)
        if n != self.qubits_number:
            raise QdcError("Size of the given state does not match the size of the tensor.")
        self._lib.call("set_from_host", self._ptr, state.ctypes.data, n)

    def conj_and_double(self) -> "QuantizedTensor":
        t = self._alloc(self._lib, self.qubits_number)
        self._lib.call("conj_and_double", self._ptr, t._ptr, self.qubits_number)
        return t

    def add(self, other: "QuantizedTensor"):
        if self.qubits_number != other.qubits_number:
            raise QdcError("Tensors have diferent sizes.")
        self._lib.call("add", other._ptr, self._ptr, self.qubits_number)

    def get_cpu_state_copy(self) -> np.ndarray:
        out = np.empty(1 << self.qubits_number, dtype=self._lib.cdtype)
        self._lib.call("copy_to_host", self._ptr, out.ctypes.data, self.qubits_number)
        return out

    def _check1(self, pos):
        if not pos < self.qubits_number:
            raise QdcError("pos is out of the bound.")

    def _check2(self, pos2, pos1):
        if pos1 == pos2:
            raise QdcError("pos1 and pos2 must be different.")
        if not pos1 < self.qubits_number:
            raise QdcError("pos1 is out of the bound.")
        if not pos2 < self.qubits_number:
            raise QdcError("pos2 is out of the bound.")

    def apply_q1_gate(self, gate, pos: int):
        g = self._lib.host(gate, 4)
        self._check1(pos)
        self._lib.call("q1gate", self._ptr, g.ctypes.data, pos, self.qubits_number)

    def apply_q1_gate_inv(self, gate, pos: int):
        g = self._lib.host(gate, 4)
        self._check1(pos)
        self._lib.call("q1gate_inv", self._ptr, g.ctypes.data, pos, self.qubits_number)

    def apply_q1_gate_tr(self, gate, pos: int):
        g = np.array(self._lib.host(gate, 4), copy=True)
        g[[1, 2]] = g[[2, 1]]  # swap(1, 2), src/quantized_tensor.rs:112
        self.apply_q1_gate(g, pos)

    def apply_q1_gate_conj_tr(self, gate, pos: int):
        g = self._lib.host(gate, 4).conj()
        g[[1, 2]] = g[[2, 1]]
        self.apply_q1_gate(g, pos)

    def apply_q2_gate(self, gate, pos2: int, pos1: int):
        g = self._lib.host(gate, 16)
        self._check2(pos2, pos1)
        self._lib.call("q2gate", self._ptr, g.ctypes.data, pos2, pos1, self.qubits_number)

    def apply_q2_gate_inv(self, gate, pos2: int, pos1: int):
        g = self._lib.host(gate, 16)
        self._check2(pos2, pos1)
        self._lib.call("q2gate_inv", self._ptr, g.ctypes.data, pos2, pos1, self.qubits_number)

    @staticmethod
    def _tr16(g):
        # the six swaps of src/quantized_tensor.rs:136-137
        for a, b in ((1, 4), (2, 8), (6, 9), (3, 12), (7, 13), (11, 14)):
            g[a], g[b] = g[b], g[a]
        return g

    def apply_q2_gate_tr(self, gate, pos2: int, pos1: int):
        g = self._tr16(np.array(self._lib.host(gate, 16), copy=True))
        self.apply_q2_gate(g, pos2, pos1)

    def apply_q2_gate_conj_tr(self, gate, pos2: int, pos1: int):
        g = self._tr16(self._lib.host(gate, 16).conj())
        self.apply_q2_gate(g, pos2, pos1)

    def apply_q2_gate_diag(self, gate, pos2: int, pos1: int):
        g = self._lib.host(gate, 4)
        self._check2(pos2, pos1)
        self._lib.call("q2gate_diag", self._ptr, g.ctypes.data, pos2, pos1, self.qubits_number)

    def apply_q2_gate_diag_conj(self, gate, pos2: int, pos1: int):
        self.apply_q2_gate_diag(self._lib.host(gate, 4).conj(), pos2, pos1)

    def get_q1_density(self, pos: int) -> np.ndarray:
        out = np.zeros(4, dtype=self._lib.cdtype)
        self._lib.call("get_q1density", self._ptr, out.ctypes.data, pos, self.qubits_number)
        return out

    def get_q2_density(self, pos2: int, pos1: int) -> np.ndarray:
        out = np.zeros(16, dtype=self._lib.cdtype)
        self._lib.call("get_q2density", self._ptr, out.ctypes.data, pos2, pos1, self.qubits_number)
        return out


# ---- free functions (src/quantized_tensor.rs:169-221) ----
def data_transfer(src: QuantizedTensor, dst: QuantizedTensor):
    if src.qubits_number != dst.qubits_number:
        raise QdcError("fwd and bwd have different lengths.")
    src._lib.call("copy", src._ptr, dst._ptr, src.qubits_number)


def _check_pair(fwd, bwd):
    if fwd.qubits_number != bwd.qubits_number:
        raise QdcError("fwd and bwd have different lengths.")


def get_q1_grad(fwd: QuantizedTensor, bwd: QuantizedTensor, pos: int) -> np.ndarray:
    _check_pair(fwd, bwd)
    if not pos < fwd.qubits_number:
        raise QdcError("pos out of range.")
    out = np.zeros(4, dtype=fwd.dtype)
    fwd._lib.call("q1grad", fwd._ptr, bwd._ptr, out.ctypes.data, pos, bwd.qubits_number)
    return out


def get_q2_grad(fwd: QuantizedTensor, bwd: QuantizedTensor, pos2: int, pos1: int) -> np.ndarray:
    _check_pair(fwd, bwd)
    fwd._check2(pos2, pos1)
    out = np.zeros(16, dtype=fwd.dtype)
    fwd._lib.call("q2grad", fwd._ptr, bwd._ptr, out.ctypes.data, pos2, pos1, bwd.qubits_number)
    return out


def get_q2_grad_diag(fwd: QuantizedTensor, bwd: QuantizedTensor, pos2: int, pos1: int) -> np.ndarray:
    _check_pair(fwd, bwd)
    fwd._check2(pos2, pos1)
    out = np.zeros(4, dtype=fwd.dtype)
    fwd._lib.call("q2grad_diag", fwd._ptr, bwd._ptr, out.ctypes.data, pos2, pos1, bwd.qubits_number)
    return out
