"""CPU checks of the drop-in boundary: the built libraries load without a GPU
and export every symbol that include/*.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


def declared_symbols():
    names = []
    for fn in sorted(os.listdir(INCLUDE)):
        text = open(os.path.join(INCLUDE, fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"^\s*(?:const\s+char\s*\*|void|size_t|int)\s+(\w+)\s*\(", text, flags=re.M):
            names.append(m.group(1))
    return sorted(set(names))


def test_headers_declare_the_18_reference_symbols():
    """The FFI block of /root/reference/src/primitives_bind.rs:15-119."""
    legacy = {"set2standard", "get_state", "drop_state", "copy_to_host", "q1gate", "q1gate_inv", "q2gate",
              "q2gate_inv", "q2gate_diag", "set_from_host", "get_q1density", "get_q2density", "q1grad", "q2grad",
              "q2grad_diag", "conj_and_double", "add", "copy"}
    assert legacy <= set(declared_symbols())


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_library_loads_and_exports_every_declared_symbol(pkg, precision):
    lib = pkg.get_lib(precision)          # raises ImportError if the .so is missing (no fallback)
    assert lib.cdll.qdc_precision().decode() == precision
    assert lib.cdll.qdc_abi_version() >= 1
    for name in declared_symbols():
        assert hasattr(lib.cdll, name), f"{name} declared in include/ but not exported by {lib.path}"
    out = subprocess.run(["nm", "-D", "--defined-only", lib.path], capture_output=True, text=True).stdout
    # every defined dynamic symbol whatever its type (T, W weak template instantiations, V, D, B ...)
    exported = {line.split()[-1] for line in out.splitlines() if len(line.split()) >= 3}
    assert set(declared_symbols()) <= exported
    # nothing but the ABI leaks out of the library (csrc/qdc_exports.map)
    leaked = sorted(exported - set(declared_symbols()))
    assert leaked == [], f"symbols outside include/*.h are exported: {leaked[:8]}"


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_library_does_not_embed_the_static_cuda_runtime(pkg, precision):
    """Built with -cudart shared: the library names only the runtime entry points it calls."""
    out = subprocess.run(["nm", "-D", "--undefined-only", pkg.lib_path(precision)], capture_output=True,
                         text=True).stdout
    assert "cudaLaunchKernel" in out
    blob = open(pkg.lib_path(precision), "rb").read()
    assert b"MemcpyBatch" not in blob


def test_python_binding_covers_every_declared_symbol(pkg):
    assert set(declared_symbols()) == set(pkg._ffi.ALL_SYMBOLS)


def test_missing_library_fails_loudly(pkg, tmp_path):
    with pytest.raises(ImportError, match="no CPU fallback"):
        pkg._ffi.Lib("f32", path=str(tmp_path / "nope.so"))


def test_sm100a_code_is_embedded(pkg):
    out = subprocess.run(["cuobjdump", "-lelf", pkg.lib_path("f32")], capture_output=True, text=True).stdout
    assert "sm_100a" in out
