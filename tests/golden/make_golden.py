"""Generate golden vectors by RUNNING THE REFERENCE's own scoring code (CPU) on seeded inputs.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Imports, unmodified, from /root/reference/third_party/ColBERT:
    colbert.modeling.colbert.colbert_score          (padded path, true max with -9999 fill)
    colbert.modeling.colbert.colbert_score_reduce
    colbert.modeling.colbert.colbert_score_packed   (CPU packed path -> segmented_maxsim_cpp)
    ColBERT.segmented_maxsim                        (JIT-built from segmented_maxsim.cpp)
through the harness-side shims of SURVEY.md Appendix A (stdlib json as ujson, hashable DefaultVal,
transformers.AdamW alias).  Inputs are bf16-rounded so the CUDA path can consume them exactly.
Fixtures are written next to this file as compressed .npz (bf16 payloads stored as uint16 bits).
"""
from __future__ import annotations

import dataclasses
import json
import os
import sys

import numpy as np
import torch
import transformers

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/third_party/ColBERT"


def import_reference():
    sys.modules.setdefault("ujson", json)
    _orig = dataclasses.dataclass

    def _patched(cls=None, /, **kw):
        def wrap(c):
            c = _orig(c, **kw)
            if c.__name__ == "DefaultVal":
                c.__hash__ = object.__hash__
            return c
        return wrap if cls is None else wrap(cls)

    dataclasses.dataclass = _patched
    if not hasattr(transformers, "AdamW"):
        transformers.AdamW = torch.optim.AdamW
    sys.path.insert(0, REF)
    os.environ.setdefault("TORCH_EXTENSIONS_DIR", "/tmp/flmr_ref_torch_ext")
    from colbert.infra.config import ColBERTConfig
    from colbert.modeling.colbert import (ColBERT, colbert_score, colbert_score_packed,
                                          colbert_score_reduce)
    ColBERT.try_load_torch_extensions(False)  # JIT-builds segmented_maxsim.cpp
    return ColBERTConfig, ColBERT, colbert_score, colbert_score_packed, colbert_score_reduce


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def make_case(name, n_passages, nd_lo, nd_hi, n_queries, nq, seed, fns, zero_query_rows=0):
    ColBERTConfig, ColBERT, colbert_score, colbert_score_packed, colbert_score_reduce = fns
    cfg = ColBERTConfig(total_visible_gpus=0)
    g = torch.Generator().manual_seed(seed)
    doclens = torch.randint(nd_lo, nd_hi + 1, (n_passages,), generator=g)
    n_tok = int(doclens.sum())
    D = torch.nn.functional.normalize(torch.randn(n_tok, 128, generator=g), dim=-1).bfloat16().float()
    Q = torch.nn.functional.normalize(torch.randn(n_queries, nq, 128, generator=g), dim=-1)
    if zero_query_rows:
        Q[:, -zero_query_rows:, :] = 0.0  # masked query tokens are exact zero rows (FLMR.py:80)
    Q = Q.bfloat16().float()
    off = torch.cat([torch.zeros(1, dtype=torch.long), doclens.cumsum(0)])
    nd_max = int(doclens.max())
    D_padded = torch.zeros(n_passages, nd_max, 128)
    D_mask = torch.zeros(n_passages, nd_max, dtype=torch.bool)
    for p in range(n_passages):
        D_padded[p, :doclens[p]] = D[off[p]:off[p + 1]]
        D_mask[p, :doclens[p]] = True

    scores_padded_path = []   # colbert_score: true max
    scores_packed_path = []   # colbert_score_packed (CPU): max clamped at 0
    for b in range(n_queries):
        scores_padded_path.append(colbert_score(Q[b:b + 1], D_padded.clone(), D_mask.unsqueeze(-1), cfg))
        scores_packed_path.append(colbert_score_packed(Q[b:b + 1], D, doclens.long(), cfg))
    scores_padded_path = torch.stack(scores_padded_path).float()
    scores_packed_path = torch.stack(scores_packed_path).float()

    # raw pieces, for pinning the oracle's helper functions
    raw = (D @ Q[0].T).contiguous()                                   # [n_tok, nq]
    seg = ColBERT.segmented_maxsim(raw, doclens.long()).float()
    red_in = (D_padded @ Q[0:1].permute(0, 2, 1)).contiguous()         # [n, nd_max, nq]
    red = colbert_score_reduce(red_in.clone(), D_mask.unsqueeze(-1), cfg).float()

    k = min(10, n_passages)
    order = scores_padded_path.sort(dim=1, descending=True)           # IndexScorer.rank: scores.sort
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(
        path,
        Q_bf16=bf16_bits(Q), D_bf16=bf16_bits(D), doclens=doclens.numpy().astype(np.int32),
        scores_true_max=scores_padded_path.numpy(), scores_relu=scores_packed_path.numpy(),
        segmented_maxsim_q0=seg.numpy(), reduce_q0=red.numpy(),
        topk_scores=order.values[:, :k].numpy(), topk_pids=order.indices[:, :k].numpy().astype(np.int64),
        meta=np.array(json.dumps(dict(name=name, seed=seed, n_passages=n_passages, nq=nq,
                                      n_queries=n_queries, torch=torch.__version__,
                                      reference="LinWeizheDragon/Retrieval-Augmented-Visual-Question-Answering@9b2b656"))))
    diff = (scores_padded_path - scores_packed_path).abs().max().item()
    print("%s: tokens=%d  |true_max - relu|max=%.4f  -> %s (%.1f KB)" %
          (name, n_tok, diff, path, os.path.getsize(path) / 1024))


def main():
    fns = import_reference()
    # ragged short docs: the ReLU and true-max paths differ here (SURVEY hazard 1)
    make_case("g1_ragged_short", n_passages=40, nd_lo=1, nd_hi=64, n_queries=2, nq=32, seed=0, fns=fns)
    # config C1 shape (Nq=32, Nd=64), cut down to fixture size
    make_case("g2_c1_shape", n_passages=64, nd_lo=64, nd_hi=64, n_queries=4, nq=32, seed=1, fns=fns)
    # north-star shape (Nq=320, Nd=180): several M-tiles
    make_case("g3_nq320_nd180", n_passages=12, nd_lo=180, nd_hi=180, n_queries=2, nq=320, seed=2, fns=fns)
    # ragged + zero query rows + nq not a multiple of 32
    make_case("g4_ragged_zero_rows", n_passages=30, nd_lo=3, nd_hi=97, n_queries=3, nq=45, seed=3,
              fns=fns, zero_query_rows=5)


if __name__ == "__main__":
    main()
