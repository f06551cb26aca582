"""Gate set of the reference's src/common_gates.rs:19-34 (flat, row-major)."""
import numpy as np


def get_hadamard(dtype=np.complex64) -> np.ndarray:
    """src/common_gates.rs:19-24."""
    s = 1.0 / np.sqrt(2.0)
    return np.array([s, s, s, -s], dtype=dtype)


def get_cnot(dtype=np.complex64) -> np.ndarray:
    """src/common_gates.rs:27-34: control = pos2 (index-2 legs), target = pos1."""
    return np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0], dtype=dtype)
