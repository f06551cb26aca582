"""Build experimental f32 variants of the library side by side (A/B runs on one GPU box):
    python profiles/scripts/build_variants.py name1:-DFOO=1,-DBAR=2 name2:-DBAZ=0 ...
-> lib/libqdc_b200_f32_<name>.so, selected at run time with QDC_LIB_VARIANT=<name>."""
import importlib.util, os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spec = importlib.util.spec_from_file_location("b", os.path.join(ROOT, "differentiable-quantum-circuit-cuda_b200", "build.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)


def one(arg):
    name, _, defs = arg.partition(":")
    out = os.path.join(b.LIB, f"libqdc_b200_f32_{name}.so")
    cmd = [b.NVCC] + b.FLAGS + [d for d in defs.split(",") if d] + ["-o", out, os.path.join(b.CSRC, "qdc_lib.cu")]
    subprocess.run(cmd, check=True, cwd=b.CSRC)
    return out


with ThreadPoolExecutor(4) as ex:
    for o in ex.map(one, sys.argv[1:]):
        print("built", o)
