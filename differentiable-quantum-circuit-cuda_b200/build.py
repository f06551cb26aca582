"""Build the sm_100a shared libraries in-tree (no JIT cache: the built .so
travels to the GPU box with the repo snapshot).

    python differentiable-quantum-circuit-cuda_b200/build.py [--force]

Produces lib/libqdc_b200_f32.so and lib/libqdc_b200_f64.so (one precision per
library, like the reference's `--features f64` build, /root/reference/build.rs:17-20).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xlinker", "-Bsymbolic",
    # only the C ABI is exported (weak libstdc++ instantiations stay local)
    "-Xlinker", "--version-script=" + os.path.join(CSRC, "qdc_exports.map"),
    # dynamic CUDA runtime: the static one embeds the names of every runtime entry point
    "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64",
    "--extended-lambda",
]


def _sources():
    return sorted(
        os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp", ".h", ".map"))
    )


def _digest(extra: str) -> str:
    h = hashlib.sha256(extra.encode())
    for p in _sources():
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def lib_path(precision: str) -> str:
    return os.path.join(LIB, f"libqdc_b200_{precision}.so")


def _build_one(precision, defs, force, verbose):
    out = lib_path(precision)
    stamp = out + ".sha256"
    digest = _digest(" ".join(FLAGS + defs))
    if not force and os.path.exists(out) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return
    cmd = [NVCC] + FLAGS + defs + ["-o", out, os.path.join(CSRC, "qdc_lib.cu")]
    if verbose:
        print("[build]", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, cwd=CSRC)
    with open(stamp, "w") as f:
        f.write(digest)


def build(force: bool = False, verbose: bool = True) -> None:
    """Both precision builds, side by side (each is one ~3 minute nvcc translation unit)."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(LIB, exist_ok=True)
    with ThreadPoolExecutor(2) as ex:
        futs = [ex.submit(_build_one, precision, defs, force, verbose)
                for precision, defs in (("f32", []), ("f64", ["-DQDC_F64"]))]
        for f in futs:
            f.result()


if __name__ == "__main__":
    build(force="--force" in sys.argv)
