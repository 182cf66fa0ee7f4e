"""Compile the REFERENCE's own native CPU kernels for this path, in place, into oracle/_ref/.

    /root/reference/third_party/ColBERT/colbert/modeling/segmented_maxsim.cpp   -> segmented_maxsim_cpp.so
    /root/reference/third_party/ColBERT/colbert/search/filter_pids.cpp          -> filter_pids_cpp.so
    /root/reference/third_party/ColBERT/colbert/search/decompress_residuals.cpp -> decompress_residuals_cpp.so
    /root/reference/third_party/ColBERT/colbert/search/segmented_lookup.cpp     -> segmented_lookup_cpp.so

(pybind11 torch extensions, the same names the reference JIT-builds at colbert/modeling/colbert.py:44-62,
colbert/search/index_storage.py:29-59 and colbert/search/strided_tensor.py:20-38.)

Compiled with g++ directly on the sources where they lie (no reference build system, no copy of the
sources into this repository).  Needs /root/reference, so it only runs in the build container; the
resulting .so files are git-ignored but ship to the GPU box with the repository snapshot.  TEST
INFRASTRUCTURE: used to validate the oracle restatements and as the `reference` CPU baselines
(exhaustive MaxSim and the PLAID CPU search of oracle/plaid_search.py).
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/third_party/ColBERT/colbert"
SOURCES = {
    "segmented_maxsim_cpp": os.path.join(REF, "modeling", "segmented_maxsim.cpp"),
    "filter_pids_cpp": os.path.join(REF, "search", "filter_pids.cpp"),
    "decompress_residuals_cpp": os.path.join(REF, "search", "decompress_residuals.cpp"),
    "segmented_lookup_cpp": os.path.join(REF, "search", "segmented_lookup.cpp"),
}
OUT_DIR = os.path.join(HERE, "_ref")
SRC = SOURCES["segmented_maxsim_cpp"]
OUT = os.path.join(OUT_DIR, "segmented_maxsim_cpp.so")


def out_path(name: str) -> str:
    return os.path.join(OUT_DIR, name + ".so")


def build_one(name: str, force: bool = False) -> str | None:
    src, out = SOURCES[name], out_path(name)
    if not os.path.exists(src):
        return out if os.path.exists(out) else None
    if os.path.exists(out) and not force and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    import torch
    from torch.utils import cpp_extension
    os.makedirs(OUT_DIR, exist_ok=True)
    inc = [f"-I{p}" for p in cpp_extension.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cmd = ["g++", "-O3", "-std=c++17", "-shared", "-fPIC", "-pthread",
           f"-DTORCH_EXTENSION_NAME={name}", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={abi}", *inc, src, "-o", out,
           f"-L{torch_lib}", "-ltorch", "-ltorch_cpu", "-ltorch_python", "-lc10",
           f"-Wl,-rpath,{torch_lib}"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("g++ failed on the reference source %s:\n%s" % (src, proc.stderr[-4000:]))
    return out


def build(force: bool = False) -> str | None:
    """Build every reference extension; returns the segmented_maxsim path (None when unavailable)."""
    paths = {name: build_one(name, force) for name in SOURCES}
    return paths["segmented_maxsim_cpp"]


def load(name: str = "segmented_maxsim_cpp"):
    """Import one compiled reference extension (returns the module, or None if absent)."""
    out = out_path(name)
    if not os.path.exists(out):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location(name, out)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    for n in SOURCES:
        print(n, out_path(n) if os.path.exists(out_path(n)) else None)
