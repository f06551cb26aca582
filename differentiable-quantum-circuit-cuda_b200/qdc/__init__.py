from .circuit import AutoGradCircuit

__all__ = ["AutoGradCircuit"]
