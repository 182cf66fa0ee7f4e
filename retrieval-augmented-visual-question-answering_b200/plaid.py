"""Reader for the reference's PLAID index directories (written by
third_party/ColBERT/colbert/indexing/collection_indexer.py + index_saver.py + codecs/residual.py)
and GPU decode into the flat bf16 store the scan kernel streams, so a user of the
reference can switch without re-encoding the collection.

Only what the exhaustive scan needs is read: ``metadata.json`` (nbits, dim, num_chunks),
``centroids.pt``, ``buckets.pt``, ``<c>.codes.pt``, ``<c>.residuals.pt``, ``doclens.<c>.json``.
The IVF (``ivf.pid.pt``) and ``avg_residual.pt`` only serve PLAID's candidate generation / pruning.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Tuple

import numpy as np
import torch

from . import _cabi


def read_plaid_metadata(path: str) -> dict:
    with open(os.path.join(path, "metadata.json")) as f:
        meta = json.load(f)
    cfg = meta["config"]
    return {"nbits": int(cfg["nbits"]), "dim": int(cfg["dim"]), "num_chunks": int(meta["num_chunks"]),
            "num_embeddings": int(meta.get("num_embeddings", -1))}


def decode_chunk(codes: torch.Tensor, residuals: torch.Tensor, centroids: torch.Tensor,
                 bucket_weights: torch.Tensor, nbits: int, out: torch.Tensor, normalize: bool = True) -> None:
    """GPU decode of one chunk into ``out`` (bf16 ``[n, 128]`` CUDA, may be a slice of the corpus matrix)."""
    dev = out.device
    codes = codes.to(dev, torch.int32).contiguous()
    residuals = residuals.to(dev, torch.uint8).contiguous()
    n = codes.numel()
    if residuals.shape != (n, _cabi.DIM * nbits // 8) or out.shape != (n, _cabi.DIM) or out.dtype != torch.bfloat16:
        raise ValueError("shape mismatch: codes %s residuals %s out %s" % (tuple(codes.shape),
                         tuple(residuals.shape), tuple(out.shape)))
    if not out.is_contiguous():
        raise ValueError("out must be contiguous")
    with torch.cuda.device(dev):
        _cabi.check(_cabi.lib().flmr_plaid_decode(
            C.c_void_p(codes.data_ptr()), C.c_void_p(residuals.data_ptr()), n,
            C.c_void_p(centroids.data_ptr()), centroids.size(0), C.c_void_p(bucket_weights.data_ptr()),
            nbits, _cabi.DIM, int(normalize), C.c_void_p(out.data_ptr()), int(dev.index),
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))


def plaid_to_flat(path: str, device=None) -> Tuple[torch.Tensor, np.ndarray]:
    """Decode a whole PLAID index: returns (tokens bf16 ``[n_tokens, 128]`` on the GPU, doclens int32)."""
    if not torch.cuda.is_available():
        raise RuntimeError("PLAID decode runs on the GPU; there is no CPU fallback")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    meta = read_plaid_metadata(path)
    if meta["dim"] != _cabi.DIM:
        raise ValueError("dim=%d (only %d is supported)" % (meta["dim"], _cabi.DIM))
    centroids = torch.load(os.path.join(path, "centroids.pt"), map_location="cpu").float().to(device).contiguous()
    _cutoffs, weights = torch.load(os.path.join(path, "buckets.pt"), map_location="cpu")
    weights = weights.float().to(device).contiguous()
    doclens, sizes = [], []
    for c in range(meta["num_chunks"]):
        with open(os.path.join(path, "doclens.%d.json" % c)) as f:
            dl = json.load(f)
        doclens.extend(dl)
        sizes.append(int(sum(dl)))
    tokens = torch.empty((sum(sizes), _cabi.DIM), dtype=torch.bfloat16, device=device)
    row = 0
    for c, n in enumerate(sizes):
        codes = torch.load(os.path.join(path, "%d.codes.pt" % c), map_location="cpu")
        residuals = torch.load(os.path.join(path, "%d.residuals.pt" % c), map_location="cpu")
        if codes.numel() != n:
            raise ValueError("chunk %d holds %d codes but its doclens sum to %d" % (c, codes.numel(), n))
        decode_chunk(codes, residuals, centroids, weights, meta["nbits"], tokens[row:row + n])
        row += n
    return tokens, np.asarray(doclens, dtype=np.int32)
