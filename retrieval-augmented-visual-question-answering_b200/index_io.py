"""Flat index on disk: the format that replaces the reference's PLAID index directory
(``centroids.pt`` / ``*.codes.pt`` / ``*.residuals.pt`` / ``ivf.pid.pt`` ..., SURVEY.md Appendix C;
written by third_party/ColBERT/colbert/indexing/collection_indexer.py + index_saver.py).

A flat index is simply what the scan kernel streams, in one piece or in chunks of passages
(the reference also writes per-chunk files, index_saver.py:75-90):

    metadata.json          {"format": "flmr-flat-v1", "n_passages", "n_tokens", "dim",
                            "dtype": "bfloat16", "num_chunks": C}            (C = 0: single file)
    single file layout     doclens.npy (int32 [n_passages]) + tokens.bf16 (raw bf16 [n_tokens, dim])
    chunked layout         doclens.<c>.npy + tokens.<c>.bf16 + <c>.metadata.json
                           {"passage_offset", "num_passages", "num_embeddings"}   for c < C
"""
from __future__ import annotations

import json
import os
from typing import Optional, Tuple

import numpy as np
import torch

FORMAT = "flmr-flat-v1"


def _write_tokens(path: str, tokens: torch.Tensor) -> None:
    tokens = tokens.detach().to("cpu", torch.bfloat16).contiguous()
    tmp = path + ".tmp"
    tokens.view(torch.int16).numpy().tofile(tmp)
    os.replace(tmp, path)          # a chunk either exists completely or not at all (resume safety)


def save_flat_index(path: str, tokens: torch.Tensor, doclens) -> str:
    os.makedirs(path, exist_ok=True)
    doclens = np.ascontiguousarray(np.asarray(doclens), dtype=np.int32)
    if tokens.dim() != 2 or int(doclens.sum()) != tokens.size(0):
        raise ValueError("tokens must be [sum(doclens), dim]")
    _write_tokens(os.path.join(path, "tokens.bf16"), tokens)
    np.save(os.path.join(path, "doclens.npy"), doclens)
    with open(os.path.join(path, "metadata.json"), "w") as f:
        json.dump({"format": FORMAT, "n_passages": int(doclens.shape[0]), "n_tokens": int(tokens.size(0)),
                   "dim": int(tokens.size(1)), "dtype": "bfloat16", "num_chunks": 0}, f)
    return path


def save_flat_chunk(path: str, chunk_idx: int, passage_offset: int, tokens: torch.Tensor, doclens) -> None:
    """One chunk of a chunked flat index (cf. IndexSaver._write_chunk_to_disk, index_saver.py:75-90).
    The chunk's metadata file is written LAST: its presence marks the chunk complete."""
    os.makedirs(path, exist_ok=True)
    doclens = np.ascontiguousarray(np.asarray(doclens), dtype=np.int32)
    if tokens.dim() != 2 or int(doclens.sum()) != tokens.size(0):
        raise ValueError("tokens must be [sum(doclens), dim]")
    _write_tokens(os.path.join(path, "tokens.%d.bf16" % chunk_idx), tokens)
    np.save(os.path.join(path, "doclens.%d.npy" % chunk_idx), doclens)
    meta = {"passage_offset": int(passage_offset), "num_passages": int(doclens.shape[0]),
            "num_embeddings": int(tokens.size(0)), "dim": int(tokens.size(1))}
    tmp = os.path.join(path, "%d.metadata.json.tmp" % chunk_idx)
    with open(tmp, "w") as f:
        json.dump(meta, f)
    os.replace(tmp, os.path.join(path, "%d.metadata.json" % chunk_idx))


def chunk_exists(path: str, chunk_idx: int) -> bool:
    """cf. IndexSaver.check_chunk_exists (index_saver.py:30-50)."""
    return all(os.path.exists(os.path.join(path, n % chunk_idx))
               for n in ("%d.metadata.json", "doclens.%d.npy", "tokens.%d.bf16"))


def finalize_chunked_index(path: str, num_chunks: int) -> dict:
    """Check the chunks tile the passage range and write the index-level metadata.json
    (cf. CollectionIndexer.finalize, collection_indexer.py:341-444, minus codec/IVF)."""
    offset = n_tokens = 0
    dim = None
    for c in range(num_chunks):
        with open(os.path.join(path, "%d.metadata.json" % c)) as f:
            m = json.load(f)
        if m["passage_offset"] != offset:
            raise ValueError("chunk %d starts at passage %d, expected %d" % (c, m["passage_offset"], offset))
        offset += m["num_passages"]
        n_tokens += m["num_embeddings"]
        dim = m["dim"] if dim is None else dim
    meta = {"format": FORMAT, "n_passages": offset, "n_tokens": n_tokens, "dim": dim,
            "dtype": "bfloat16", "num_chunks": num_chunks}
    with open(os.path.join(path, "metadata.json"), "w") as f:
        json.dump(meta, f)
    return meta


def index_token_files(path: str):
    """(doclens of the whole index int32 [n_passages], [(token file, first passage, passages, rows), ...]) —
    what a streaming loader needs to pick the byte ranges of a passage shard without reading any token."""
    with open(os.path.join(path, "metadata.json")) as f:
        meta = json.load(f)
    if meta.get("format") != FORMAT:
        raise ValueError("%s is not a %s index" % (path, FORMAT))
    if meta.get("num_chunks", 0) == 0:
        doclens = np.load(os.path.join(path, "doclens.npy")).astype(np.int32)
        return doclens, [(os.path.join(path, "tokens.bf16"), 0, len(doclens), int(doclens.sum()))]
    dls, files, offset = [], [], 0
    for c in range(meta["num_chunks"]):
        dl = np.load(os.path.join(path, "doclens.%d.npy" % c)).astype(np.int32)
        files.append((os.path.join(path, "tokens.%d.bf16" % c), offset, len(dl), int(dl.sum())))
        dls.append(dl)
        offset += len(dl)
    return np.concatenate(dls), files


def _read_tokens(fname: str, n_rows: int, dim: int, row0: int = 0, row1: Optional[int] = None) -> torch.Tensor:
    row1 = n_rows if row1 is None else row1
    mm = np.memmap(fname, dtype=np.int16, mode="r", shape=(n_rows, dim))
    return torch.from_numpy(np.array(mm[row0:row1], copy=True)).view(torch.bfloat16)


def load_flat_index(path: str, passage_range: Optional[Tuple[int, int]] = None
                    ) -> Tuple[torch.Tensor, np.ndarray, dict]:
    """Returns (tokens bf16 [n_tok, dim] on CPU, doclens int32, metadata) of the whole index or of the
    passages ``[p0, p1)`` only (a GPU's shard) — only the chunks that overlap are read."""
    with open(os.path.join(path, "metadata.json")) as f:
        meta = json.load(f)
    if meta.get("format") != FORMAT:
        raise ValueError("%s is not a %s index" % (path, FORMAT))
    dim = meta["dim"]
    p0, p1 = passage_range if passage_range is not None else (0, meta["n_passages"])
    if not 0 <= p0 <= p1 <= meta["n_passages"]:
        raise ValueError("passage_range %s outside [0, %d]" % (passage_range, meta["n_passages"]))
    if meta.get("num_chunks", 0) == 0:
        doclens = np.load(os.path.join(path, "doclens.npy"))
        off = np.concatenate([[0], np.cumsum(doclens, dtype=np.int64)])
        tokens = _read_tokens(os.path.join(path, "tokens.bf16"), meta["n_tokens"], dim, int(off[p0]), int(off[p1]))
        return tokens, doclens[p0:p1], meta
    toks, dls = [], []
    for c in range(meta["num_chunks"]):
        with open(os.path.join(path, "%d.metadata.json" % c)) as f:
            m = json.load(f)
        c0, c1 = m["passage_offset"], m["passage_offset"] + m["num_passages"]
        if c1 <= p0 or c0 >= p1:
            continue
        doclens = np.load(os.path.join(path, "doclens.%d.npy" % c))
        off = np.concatenate([[0], np.cumsum(doclens, dtype=np.int64)])
        a, b = max(p0, c0) - c0, min(p1, c1) - c0
        toks.append(_read_tokens(os.path.join(path, "tokens.%d.bf16" % c), m["num_embeddings"], dim,
                                 int(off[a]), int(off[b])))
        dls.append(doclens[a:b])
    tokens = torch.cat(toks) if toks else torch.empty((0, dim), dtype=torch.bfloat16)
    doclens = np.concatenate(dls) if dls else np.empty(0, dtype=np.int32)
    return tokens, doclens.astype(np.int32), meta
