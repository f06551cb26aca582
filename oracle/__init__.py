"""CPU oracle for the differentiable statevector hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or as
the reported baseline -- never as the thing measured or shipped.

Parity pinning (see DESIGN.md "Oracle"):
  * ``oracle.statevector`` restates the arithmetic of
    ``/root/reference/src/primitives.cu`` in NumPy (complex128 by default).
  * ``oracle.literal`` re-executes the reference kernels' *index arithmetic*
    (INSERT_ZERO bit insertion, flat gate indices) element by element in pure
    Python for small n; tests check statevector == literal.
  * ``oracle.circuit`` restates the control flow of ``src/circuit.rs`` and the
    host-side gate transforms of ``src/quantized_tensor.rs``.
  * Known answers from the reference's own tests (GHZ amplitudes / densities,
    the einsum specs of ``src/quantized_tensor.rs:287-398``, the 8th-order
    finite-difference identity of ``src/test_autodiff.py``) are checked in
    ``tests/test_oracle.py``.
  * ``oracle/_ref`` (git-ignored) holds the reference's own CUDA file compiled
    unmodified for sm_100a; ``oracle.ref_replay`` drives it through ctypes with
    the exact call sequence of ``src/circuit.rs`` (needs a GPU).  Golden
    vectors produced by it on a B200 are committed under ``tests/golden/``.
"""
