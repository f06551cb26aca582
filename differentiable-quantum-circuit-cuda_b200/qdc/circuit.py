"""`AutoGradCircuit`: Python wrapper with reverse-mode differentiation, mirror
of /root/reference/src/qdc/circuit.py:8-202 (same builder methods, same
`build() -> (simple_run, autodiff_run)` contract, argument order
`(var_gates, const_gates)` at this level vs `(const, var)` one level down).

The reference glues `Circuit.forward/backward` into JAX with `jax.custom_vjp`
(src/qdc/circuit.py:177-201).  JAX is optional here:
  * if `jax` is importable, `autodiff_run` is the same `custom_vjp` function;
  * otherwise `autodiff_run` accepts torch tensors (differentiable through
    `torch.autograd`) or NumPy arrays (forward only), and exposes
    `autodiff_run.vjp(var_gates, const_gates, density_cotangents)` which is the
    reference's `bwd_run` (cotangents are conjugated, src/qdc/circuit.py:193).
Cotangent convention: JAX's.  The returned gate gradients g satisfy
dL = Re sum(g * dU)  (src/test_autodiff.py:159-164).
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np

from ..quantum_differentiable_circuit import Circuit

try:  # pragma: no cover - jax is not part of this image
    import jax  # type: ignore
    from jax import custom_vjp  # type: ignore
    _HAVE_JAX = True
except Exception:  # noqa: BLE001
    _HAVE_JAX = False


def _np(x, dtype):
    """np.asarray() of a JAX / torch / NumPy array (src/qdc/circuit.py:173-174)."""
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    a = np.asarray(x)
    return a if a.dtype == dtype else a.astype(dtype)


class _AutodiffRun:
    """Callable returned as `autodiff_run` when JAX is absent."""

    def __init__(self, circuit: Circuit):
        self._c = circuit

    def forward(self, var_gates, const_gates) -> List[np.ndarray]:
        dt = self._c.dtype
        return self._c.forward([_np(g, dt) for g in const_gates], [_np(g, dt) for g in var_gates])

    def vjp(self, var_gates, const_gates, density_cotangents) -> List[np.ndarray]:
        """`bwd_run` of src/qdc/circuit.py:190-197 (must follow `forward`)."""
        dt = self._c.dtype
        return self._c.backward(
            [_np(g, dt).conj() for g in density_cotangents],
            [_np(g, dt) for g in const_gates],
            [_np(g, dt) for g in var_gates],
        )

    def __call__(self, var_gates, const_gates):
        try:
            import torch
        except Exception:  # noqa: BLE001
            torch = None
        if torch is not None and any(isinstance(g, torch.Tensor) for g in list(var_gates) + list(const_gates)):
            return _torch_apply(self, list(var_gates), list(const_gates))
        return self.forward(var_gates, const_gates)


def _torch_apply(run: _AutodiffRun, var_gates, const_gates):
    import torch

    n_var = len(var_gates)

    class _Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, *gates):
            vg, cg = gates[:n_var], gates[n_var:]
            ctx.save_for_backward(*gates)
            dens = run.forward(vg, cg)
            return tuple(torch.from_numpy(d) for d in dens)

        @staticmethod
        def backward(ctx, *grad_outputs):
            gates = ctx.saved_tensors
            vg, cg = gates[:n_var], gates[n_var:]
            # torch's complex gradient is the conjugate of the JAX cotangent
            cts = [g.detach().cpu().numpy().conj() for g in grad_outputs]
            grads = run.vjp(vg, cg, cts)
            out = [torch.from_numpy(g.conj()).to(v.dtype).reshape(v.shape) for g, v in zip(grads, vg)]
            return tuple(out) + (None,) * len(cg)

    return list(_Fn.apply(*var_gates, *const_gates))


class AutoGradCircuit:

    def __init__(self, qubits_number: int, precision: str | None = None):
        """Quantum circuit with automatic differentiation (src/qdc/circuit.py:10-12)."""
        self.circuit = Circuit(qubits_number, precision=precision)

    def set_state_from_vector(self, vec):
        """Set initial state from an array (src/qdc/circuit.py:14-22)."""
        self.circuit.set_state_from_vector(_np(vec, self.circuit.dtype))

    # 1:1 forwarding builders, src/qdc/circuit.py:24-158.  Qubits are enumerated
    # starting from the innermost one; pos2 is the 'control', pos1 the 'target'.
    def add_q2_const_gate(self, pos2: int, pos1: int): self.circuit.add_q2_const_gate(pos2, pos1)
    def add_q2_const_gate_nonu(self, pos2: int, pos1: int): self.circuit.add_q2_const_gate_nonu(pos2, pos1)
    def add_q2_const_gate_diag(self, pos2: int, pos1: int): self.circuit.add_q2_const_gate_diag(pos2, pos1)
    def add_q2_var_gate(self, pos2: int, pos1: int): self.circuit.add_q2_var_gate(pos2, pos1)
    def add_q2_var_gate_nonu(self, pos2: int, pos1: int): self.circuit.add_q2_var_gate_nonu(pos2, pos1)
    def add_q2_var_gate_diag(self, pos2: int, pos1: int): self.circuit.add_q2_var_gate_diag(pos2, pos1)
    def add_q1_const_gate(self, pos: int): self.circuit.add_q1_const_gate(pos)
    def add_q1_const_gate_nonu(self, pos: int): self.circuit.add_q1_const_gate_nonu(pos)
    def add_q1_var_gate(self, pos: int): self.circuit.add_q1_var_gate(pos)
    def add_q1_var_gate_nonu(self, pos: int): self.circuit.add_q1_var_gate_nonu(pos)
    def get_q2_dens_op(self, pos2: int, pos1: int): self.circuit.get_q2_dens_op(pos2, pos1)
    def get_q1_dens_op(self, pos: int): self.circuit.get_q1_dens_op(pos)
    def get_q2_dens_op_with_grad(self, pos2: int, pos1: int): self.circuit.get_q2_dens_op_with_grad(pos2, pos1)
    def get_q1_dens_op_with_grad(self, pos: int): self.circuit.get_q1_dens_op_with_grad(pos)

    def build(self) -> Tuple[Callable, Callable]:
        """Returns (simple_run, autodiff_run), src/qdc/circuit.py:160-202.

        simple_run(var_gates, const_gates) evaluates every requested density
        matrix and has no backward pass; autodiff_run evaluates only those with
        gradient and supports the backward pass."""
        dt = self.circuit.dtype

        def simple_run(var_gates, const_gates):
            return self.circuit.run([_np(g, dt) for g in const_gates], [_np(g, dt) for g in var_gates])

        if not _HAVE_JAX:
            return simple_run, _AutodiffRun(self.circuit)

        @custom_vjp
        def autodiff_run(var_gates, const_gates):  # pragma: no cover
            return self.circuit.forward([_np(g, dt) for g in const_gates], [_np(g, dt) for g in var_gates])

        def fwd_run(var_gates, const_gates):  # pragma: no cover
            dens = self.circuit.forward([_np(g, dt) for g in const_gates], [_np(g, dt) for g in var_gates])
            return dens, (const_gates, var_gates)

        def bwd_run(res, density_grads):  # pragma: no cover
            const_gates, var_gates = res
            grads = self.circuit.backward(
                [_np(g, dt).conj() for g in density_grads],
                [_np(g, dt) for g in const_gates],
                [_np(g, dt) for g in var_gates],
            )
            return grads, None

        autodiff_run.defvjp(fwd_run, bwd_run)
        return simple_run, autodiff_run
