"""Shared helpers for the test-suite (golden fixtures, C oracle loader)."""
from __future__ import annotations

import ctypes as C
import glob
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    return (bits.astype(np.uint32) << 16).view(np.float32)


def load_golden(name: str) -> dict:
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["Q"] = bf16_bits_to_f32(g.pop("Q_bf16"))
    g["D"] = bf16_bits_to_f32(g.pop("D_bf16"))
    g["meta"] = json.loads(str(g["meta"]))
    return g


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "g[0-9]*.npz")))


def c_oracle():
    """Build (if needed) and load oracle/_ref/libmaxsim_oracle.so."""
    path = os.path.join(ROOT, "oracle", "_ref", "libmaxsim_oracle.so")
    src = os.path.join(ROOT, "oracle", "maxsim_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    L = C.CDLL(path)
    L.flmr_oracle_maxsim_scores.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.flmr_oracle_topk.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    return L


def c_oracle_scores(Q, D, doclens, relu=False, nthreads=4):
    L = c_oracle()
    Q = np.ascontiguousarray(Q, dtype=np.float32)
    D = np.ascontiguousarray(D, dtype=np.float32)
    dl = np.ascontiguousarray(doclens, dtype=np.int32)
    out = np.empty((Q.shape[0], len(dl)), dtype=np.float32)
    rc = L.flmr_oracle_maxsim_scores(Q.ctypes.data, Q.shape[0], Q.shape[1], D.ctypes.data, dl.ctypes.data,
                                     len(dl), Q.shape[2], int(relu), nthreads, out.ctypes.data)
    assert rc == 0
    return out
