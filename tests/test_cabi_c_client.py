"""CPU: a plain-C (gcc -std=c99) client compiles against include/flmr_maxsim.h, links the shared library and
sees the documented status codes — the drop-in boundary is a C ABI."""
import os
import subprocess

from helpers import ROOT
from ravqa_b200 import build


def test_c_client_compiles_links_and_runs(tmp_path):
    lib = build.build()
    exe = str(tmp_path / "c_client")
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_client.c"), "-o", exe, lib, "-Wl,-rpath," + os.path.dirname(lib)]
    subprocess.run(cmd, check=True, capture_output=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "c_client ok" in out.stdout
