"""Drop-in for the reference's native module `quantum_differentiable_circuit`
(/root/reference/src/circuit.rs:432-436): class `Circuit` with the same
builder methods, `run`, `forward`, `backward` -- same names, argument orders
(`const_gates` before `var_gates`), output ordering and failure conditions.

Where the reference walks its instruction list issuing one FFI call per
instruction (src/circuit.rs:175, 226, 278), this class hands the whole program
to the C ABI (include/qdc_circuit.h) once per call.

Precision: the reference is compiled for exactly one precision (cargo feature
`f64`).  Here both builds ship; `Circuit(n)` uses `QDC_PRECISION` (default
"f32", the reference's default feature set) and `Circuit(n, precision="f64")`
or the aliases `quantum_differentiable_circuit.f64.Circuit` pick the other.
Arrays must be contiguous and of the build's complex dtype, as with PyO3's
`PyReadonlyArray1<Complex>`.
"""
from __future__ import annotations

import ctypes as C
import types
from typing import List, Sequence

import numpy as np

from .._ffi import Lib, ProfileEntry, QdcError, Stats, default_precision, get_lib

# enum Instruction, src/circuit.rs:53-68
(CONST_Q2, VAR_Q2, CONST_Q2_NONU, VAR_Q2_NONU, CONST_Q2_DIAG, VAR_Q2_DIAG, CONST_Q1, CONST_Q1_NONU, VAR_Q1,
 VAR_Q1_NONU, Q2_DENS, Q1_DENS, DIFF_Q2_DENS, DIFF_Q1_DENS) = range(14)

(N_INSTRUCTIONS, N_CONST_GATES, N_VAR_GATES, N_DENSITIES, N_DIFF_DENSITIES, RUN_OUT_LEN, FORWARD_OUT_LEN,
 BACKWARD_OUT_LEN) = range(8)


class Circuit:
    def __init__(self, qubits_number: int, precision: str | None = None):
        self._lib: Lib = get_lib(precision or default_precision())
        self.qubits_number = int(qubits_number)
        h = C.c_void_p()
        self._lib.call("qdc_circuit_new", C.byref(h), self.qubits_number)
        self._h = h
        self._kinds: List[int] = []

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self._lib.call("qdc_circuit_free", self._h)
                self._h = None
        except Exception:
            pass

    @property
    def dtype(self):
        return self._lib.cdtype

    # ---- src/circuit.rs:104-106 ----
    def set_state_from_vector(self, vector):
        v = self._lib.host(vector)
        if v.ndim != 1:
            raise TypeError("vector must be one-dimensional")
        self._lib.call("qdc_circuit_set_state_from_host", self._h, v.ctypes.data, v.size)

    # ---- builders, src/circuit.rs:108-162 ----
    def _add(self, kind, pos2, pos1=0):
        if pos2 < 0 or pos1 < 0:
            raise OverflowError("can't convert negative int to unsigned")
        self._lib.call("qdc_circuit_add", self._h, kind, pos2, pos1)
        self._kinds.append(kind)

    def add_q2_const_gate(self, pos2: int, pos1: int): self._add(CONST_Q2, pos2, pos1)
    def add_q2_const_gate_diag(self, pos2: int, pos1: int): self._add(CONST_Q2_DIAG, pos2, pos1)
    def add_q2_const_gate_nonu(self, pos2: int, pos1: int): self._add(CONST_Q2_NONU, pos2, pos1)
    def add_q2_var_gate(self, pos2: int, pos1: int): self._add(VAR_Q2, pos2, pos1)
    def add_q2_var_gate_diag(self, pos2: int, pos1: int): self._add(VAR_Q2_DIAG, pos2, pos1)
    def add_q2_var_gate_nonu(self, pos2: int, pos1: int): self._add(VAR_Q2_NONU, pos2, pos1)
    def add_q1_const_gate(self, pos: int): self._add(CONST_Q1, pos)
    def add_q1_const_gate_nonu(self, pos: int): self._add(CONST_Q1_NONU, pos)
    def add_q1_var_gate(self, pos: int): self._add(VAR_Q1, pos)
    def add_q1_var_gate_nonu(self, pos: int): self._add(VAR_Q1_NONU, pos)
    def get_q2_dens_op(self, pos2: int, pos1: int): self._add(Q2_DENS, pos2, pos1)
    def get_q1_dens_op(self, pos: int): self._add(Q1_DENS, pos)
    def get_q2_dens_op_with_grad(self, pos2: int, pos1: int): self._add(DIFF_Q2_DENS, pos2, pos1)
    def get_q1_dens_op_with_grad(self, pos: int): self._add(DIFF_Q1_DENS, pos)

    # ---- marshalling ----
    def _flatten(self, arrays: Sequence, ndim: int):
        """list of small arrays -> (flat complex array, uint32 lens).  Strict dtype."""
        arrs = []
        for a in arrays:
            a = np.asarray(a)
            if a.dtype != self._lib.cdtype:
                raise TypeError(f"array has dtype {a.dtype}, this build expects {self._lib.cdtype}")
            if a.ndim != ndim:
                raise TypeError(f"expected a {ndim}-D array, got {a.ndim}-D")
            arrs.append(np.ascontiguousarray(a).reshape(-1))
        lens = np.array([a.size for a in arrs], dtype=np.uint32)
        flat = np.concatenate(arrs) if arrs else np.zeros(0, dtype=self._lib.cdtype)
        return np.ascontiguousarray(flat, dtype=self._lib.cdtype), lens

    def _count(self, what: int) -> int:
        return int(self._lib.cdll.qdc_circuit_count(self._h, what))

    def _sweep(self, name, const_gates, var_gates, out_len_sel, dens_kinds):
        cf, cl = self._flatten(const_gates, 1)
        vf, vl = self._flatten(var_gates, 1)
        cap = self._count(out_len_sel)
        out = np.empty(max(cap, 1), dtype=self._lib.cdtype)
        n_out = C.c_size_t(0)
        self._lib.call(name, self._h, cf.ctypes.data, cl.ctypes.data, cl.size, vf.ctypes.data, vl.ctypes.data,
                       vl.size, out.ctypes.data, cap, C.byref(n_out))
        res, o = [], 0
        for k in self._kinds:
            if k in dens_kinds:
                m = 2 if k in (Q1_DENS, DIFF_Q1_DENS) else 4
                res.append(out[o:o + m * m].reshape(m, m).copy())
                o += m * m
        assert o == n_out.value
        return res

    # ---- src/circuit.rs:164-212 ----
    def run(self, const_gates, var_gates) -> List[np.ndarray]:
        return self._sweep("qdc_circuit_run", const_gates, var_gates, RUN_OUT_LEN,
                           (Q1_DENS, Q2_DENS, DIFF_Q1_DENS, DIFF_Q2_DENS))

    # ---- src/circuit.rs:214-264 ----
    def forward(self, const_gates, var_gates) -> List[np.ndarray]:
        return self._sweep("qdc_circuit_forward", const_gates, var_gates, FORWARD_OUT_LEN,
                           (DIFF_Q1_DENS, DIFF_Q2_DENS))

    # ---- src/circuit.rs:266-429 ----
    def backward(self, grads_wrt_density, const_gates, var_gates) -> List[np.ndarray]:
        df, dl = self._flatten(grads_wrt_density, 2)
        cf, cl = self._flatten(const_gates, 1)
        vf, vl = self._flatten(var_gates, 1)
        cap = self._count(BACKWARD_OUT_LEN)
        out = np.empty(max(cap, 1), dtype=self._lib.cdtype)
        n_out = C.c_size_t(0)
        self._lib.call("qdc_circuit_backward", self._h, df.ctypes.data, dl.ctypes.data, dl.size,
                       cf.ctypes.data, cl.ctypes.data, cl.size, vf.ctypes.data, vl.ctypes.data, vl.size,
                       out.ctypes.data, cap, C.byref(n_out))
        res, o = [], 0
        for k in self._kinds:
            if k in (VAR_Q2, VAR_Q2_NONU):
                res.append(out[o:o + 16].copy()); o += 16
            elif k in (VAR_Q2_DIAG, VAR_Q1, VAR_Q1_NONU):
                res.append(out[o:o + 4].copy()); o += 4
        assert o == n_out.value
        return res

    # ---- extras (not in the reference's Python surface) ----
    def get_cpu_state_copy(self) -> np.ndarray:
        """QuantizedTensor::get_cpu_state_copy of the working state (src/quantized_tensor.rs:91-99)."""
        out = np.empty(1 << self.qubits_number, dtype=self._lib.cdtype)
        self._lib.call("qdc_circuit_copy_state_to_host", self._h, out.ctypes.data)
        return out

    def save_state(self, path: str):
        """Stream this rank's working state to `path` (format: include/qdc_circuit.h, `state_io.read_state`)."""
        self._lib.call("qdc_circuit_save_state", self._h, str(path).encode())

    def load_state(self, path: str):
        """Install the state file `path` (identity layout) as the initial state, like set_state_from_vector."""
        self._lib.call("qdc_circuit_load_state", self._h, str(path).encode())

    def state_layout(self) -> List[int]:
        """Physical position of every logical qubit in the working state (identity unless sharded mid-sweep)."""
        out = (C.c_int * self.qubits_number)()
        self._lib.call("qdc_circuit_state_layout", self._h, out)
        return list(out)

    def set_option(self, key: str, value: int):
        self._lib.call("qdc_circuit_set_option", self._h, key.encode(), int(value))

    def set_stream(self, cuda_stream: int):
        self._lib.call("qdc_circuit_set_stream", self._h, C.c_void_p(cuda_stream))

    def last_stats(self) -> dict:
        s = Stats()
        self._lib.call("qdc_circuit_last_stats", self._h, C.byref(s))
        return {"kernel_launches": s.kernel_launches, "hbm_passes": s.hbm_passes,
                "algorithmic_bytes": s.algorithmic_bytes}


    def last_profile(self) -> dict:
        """Per-category device time of the last call (needs set_option("profile", 1))."""
        out = {}
        for cat in range(self._lib.cdll.qdc_profile_categories()):
            e = ProfileEntry()
            self._lib.call("qdc_circuit_last_profile", self._h, cat, C.byref(e))
            if e.launches:
                out[self._lib.cdll.qdc_profile_category_name(cat).decode()] = {
                    "launches": e.launches, "ms": e.ms, "algorithmic_bytes": e.algorithmic_bytes}
        return out


def _variant(precision: str):
    mod = types.ModuleType(f"{__name__}.{precision}")

    class _Circuit(Circuit):
        def __init__(self, qubits_number: int):
            super().__init__(qubits_number, precision=precision)

    _Circuit.__name__ = "Circuit"
    mod.Circuit = _Circuit
    return mod


f32 = _variant("f32")
f64 = _variant("f64")

__all__ = ["Circuit", "QdcError", "f32", "f64"]
