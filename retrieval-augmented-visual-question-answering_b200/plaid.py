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
from typing import Optional, Tuple

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


def read_plaid_doclens(path: str, num_chunks: int):
    """Per-chunk passage lengths (``doclens.<c>.json``, index_saver.py:83-85)."""
    out = []
    for c in range(num_chunks):
        with open(os.path.join(path, "doclens.%d.json" % c)) as f:
            out.append(np.asarray(json.load(f), dtype=np.int64))
    return out


def plaid_to_flat(path: str, device=None, passage_range: Optional[Tuple[int, int]] = None
                  ) -> Tuple[torch.Tensor, np.ndarray]:
    """Decode a PLAID index — all of it, or the passages ``[p0, p1)`` only (one GPU's shard: only the chunks
    that overlap are read) — into (tokens bf16 ``[n_tokens, 128]`` on the GPU, doclens int32)."""
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
    chunk_doclens = read_plaid_doclens(path, meta["num_chunks"])
    n_total = int(sum(len(d) for d in chunk_doclens))
    p0, p1 = (0, n_total) if passage_range is None else passage_range
    if not 0 <= p0 <= p1 <= n_total:
        raise ValueError("passage_range %s outside [0, %d]" % (passage_range, n_total))
    pieces, c0 = [], 0                       # (chunk, first token row, one-past-last row, doclens slice)
    for c, dl in enumerate(chunk_doclens):
        a, b = max(p0, c0) - c0, min(p1, c0 + len(dl)) - c0
        if a < b:
            off = np.concatenate([[0], np.cumsum(dl)])
            pieces.append((c, int(off[a]), int(off[b]), dl[a:b], int(off[-1])))
        c0 += len(dl)
    tokens = torch.empty((sum(p[2] - p[1] for p in pieces), _cabi.DIM), dtype=torch.bfloat16, device=device)
    row = 0
    for c, r0, r1, _, n_chunk in pieces:
        codes = torch.load(os.path.join(path, "%d.codes.pt" % c), map_location="cpu")
        residuals = torch.load(os.path.join(path, "%d.residuals.pt" % c), map_location="cpu")
        if codes.numel() != n_chunk:
            raise ValueError("chunk %d holds %d codes but its doclens sum to %d" % (c, codes.numel(), n_chunk))
        decode_chunk(codes[r0:r1], residuals[r0:r1], centroids, weights, meta["nbits"], tokens[row:row + r1 - r0])
        row += r1 - r0
    doclens = np.concatenate([p[3] for p in pieces]) if pieces else np.empty(0, dtype=np.int64)
    return tokens, doclens.astype(np.int32)
