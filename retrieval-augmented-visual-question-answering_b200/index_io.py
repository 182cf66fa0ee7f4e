"""Flat index on disk: the format that replaces the reference's PLAID index directory
(``centroids.pt`` / ``*.codes.pt`` / ``*.residuals.pt`` / ``ivf.pid.pt`` ..., SURVEY.md Appendix C;
written by third_party/ColBERT/colbert/indexing/collection_indexer.py + index_saver.py).

A flat index is simply what the scan kernel streams:
    metadata.json   {"format": "flmr-flat-v1", "n_passages", "n_tokens", "dim", "dtype": "bfloat16"}
    doclens.npy     int32 [n_passages]
    tokens.bf16     raw little-endian bf16 [n_tokens, dim], passage after passage
"""
from __future__ import annotations

import json
import os
from typing import Tuple

import numpy as np
import torch

FORMAT = "flmr-flat-v1"


def save_flat_index(path: str, tokens: torch.Tensor, doclens) -> str:
    os.makedirs(path, exist_ok=True)
    doclens = np.ascontiguousarray(np.asarray(doclens), dtype=np.int32)
    tokens = tokens.detach().to("cpu", torch.bfloat16).contiguous()
    if tokens.dim() != 2 or int(doclens.sum()) != tokens.size(0):
        raise ValueError("tokens must be [sum(doclens), dim]")
    tokens.view(torch.int16).numpy().tofile(os.path.join(path, "tokens.bf16"))
    np.save(os.path.join(path, "doclens.npy"), doclens)
    with open(os.path.join(path, "metadata.json"), "w") as f:
        json.dump({"format": FORMAT, "n_passages": int(doclens.shape[0]),
                   "n_tokens": int(tokens.size(0)), "dim": int(tokens.size(1)),
                   "dtype": "bfloat16"}, f)
    return path


def load_flat_index(path: str) -> Tuple[torch.Tensor, np.ndarray, dict]:
    with open(os.path.join(path, "metadata.json")) as f:
        meta = json.load(f)
    if meta.get("format") != FORMAT:
        raise ValueError("%s is not a %s index" % (path, FORMAT))
    doclens = np.load(os.path.join(path, "doclens.npy"))
    raw = np.fromfile(os.path.join(path, "tokens.bf16"), dtype=np.int16)
    tokens = torch.from_numpy(raw).view(torch.bfloat16).view(meta["n_tokens"], meta["dim"])
    return tokens, doclens, meta
