"""B200-native differentiable statevector simulator -- host-side package.

The directory name is fixed by the project layout and is not a valid Python
identifier; import it with

    import importlib
    pkg = importlib.import_module("differentiable-quantum-circuit-cuda_b200")

Importing it registers the two drop-in module names of the reference:

    quantum_differentiable_circuit   (the PyO3 module, src/circuit.rs:432-436)
    qdc                              (the Python wrapper, src/qdc/__init__.py)

so that `from qdc import AutoGradCircuit` and
`from quantum_differentiable_circuit import Circuit` work unchanged.
"""
import sys as _sys

from . import _ffi
from ._ffi import QdcError, get_lib, lib_path
from . import common_gates
from . import state_io
from .quantized_tensor import QuantizedTensor, get_q1_grad, get_q2_grad, get_q2_grad_diag, data_transfer
from . import quantum_differentiable_circuit as _qdc_native
from . import qdc as _qdc_py
from .fusion import FusedCircuit

_sys.modules.setdefault("quantum_differentiable_circuit", _qdc_native)
_sys.modules.setdefault("qdc", _qdc_py)

Circuit = _qdc_native.Circuit
AutoGradCircuit = _qdc_py.AutoGradCircuit

__all__ = [
    "Circuit", "FusedCircuit", "AutoGradCircuit", "QuantizedTensor", "get_q1_grad", "get_q2_grad", "get_q2_grad_diag",
    "data_transfer", "common_gates", "state_io", "QdcError", "get_lib", "lib_path",
]
