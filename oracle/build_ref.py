"""Compile the REFERENCE's own native reduction for this path, in place, into oracle/_ref/.

    /root/reference/third_party/ColBERT/colbert/modeling/segmented_maxsim.cpp
        -> oracle/_ref/segmented_maxsim_cpp.so   (pybind11 torch extension, same name the reference
           JIT-builds at colbert/modeling/colbert.py:44-62)

Compiled with g++ directly on the source where it lies (no reference build system, no copy of the
source into this repository).  Needs /root/reference, so it only runs in the build container; the
resulting .so is git-ignored but ships to the GPU box with the repository snapshot.  TEST
INFRASTRUCTURE: used to validate the oracle restatement and as the `reference` CPU baseline.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/third_party/ColBERT/colbert/modeling/segmented_maxsim.cpp"
OUT_DIR = os.path.join(HERE, "_ref")
OUT = os.path.join(OUT_DIR, "segmented_maxsim_cpp.so")


def build(force: bool = False) -> str | None:
    if not os.path.exists(SRC):
        return OUT if os.path.exists(OUT) else None
    if os.path.exists(OUT) and not force and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    import torch
    from torch.utils import cpp_extension
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in cpp_extension.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-pthread",
           "-DTORCH_EXTENSION_NAME=segmented_maxsim_cpp", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={abi}", *inc, SRC, "-o", OUT,
           f"-L{torch_lib}", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10",
           f"-Wl,-rpath,{torch_lib}"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("g++ failed on the reference source:\n" + proc.stderr[-4000:])
    return OUT


def load():
    """Import the compiled reference extension (returns the module, or None if absent)."""
    if not os.path.exists(OUT):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("segmented_maxsim_cpp", OUT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
