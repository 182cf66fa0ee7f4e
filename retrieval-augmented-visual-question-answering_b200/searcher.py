"""``Searcher`` façade with the call surface of the reference's ``colbert.Searcher``
(third_party/ColBERT/colbert/searcher.py:22-132) so ``src/executors/FLMR_executor.py:785-792`` and
``src/models/rag/rag_model_blip.py:397`` can call it unchanged:

    Searcher(index=..., checkpoint=None, collection=None, config=None)
    ._search_all_Q(queries, Q, k, filter_fn=None, progress=True, remove_zero_tensors=False) -> Ranking
    .dense_search(Q[1,Nq,d], k, filter_fn=None, remove_zero_tensors=False) -> (pids, ranks, scores)
    .search(text, k) / .search_all(queries, k)        (need an ``encode_fn``: encoders stay PyTorch)
    Ranking.todict() -> {qid: [(pid, rank, score), ...]}   (colbert/data/ranking.py:48)

What differs, by design: scoring is exhaustive and exact (no PLAID candidate
generation / centroid pruning), all queries of a call go through one batched fused scan instead of
a Python loop of per-query ``rank`` calls, and results always hold exactly ``min(k, n_passages)``
hits.
"""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, Iterable, List, Optional, Tuple, Union

import torch

from . import _cabi
from .corpus import FlatCorpus
from .index_io import load_flat_index
from .maxsim import maxsim_scores, maxsim_topk, topk_select


class Ranking:
    """Minimal equivalent of colbert.data.Ranking (colbert/data/ranking.py:25-94): ``data`` maps
    qid -> [(pid, rank (1-based), score), ...]."""

    def __init__(self, data: Dict, provenance: Optional[dict] = None):
        self.data = data
        self.provenance_ = provenance or {}

    def provenance(self):
        return self.provenance_

    def todict(self):
        return dict(self.data)

    def tolist(self):
        return [(qid, pid, rank, score) for qid, hits in self.data.items() for pid, rank, score in hits]

    def save(self, path: str) -> str:
        with open(path, "w") as f:
            for qid, pid, rank, score in self.tolist():
                f.write("\t".join(map(str, (qid, pid, rank, score))) + "\n")
        with open(path + ".meta", "w") as f:
            json.dump(self.provenance_, f)
        return path


def _query_keys(queries, n: int) -> List:
    if queries is None:
        return list(range(n))
    if hasattr(queries, "keys"):
        keys = list(queries.keys())
    else:
        keys = list(queries)
    if len(keys) != n:
        raise ValueError("got %d query ids for %d query matrices" % (len(keys), n))
    return keys


class Searcher:
    def __init__(self, index: Union[str, FlatCorpus], checkpoint=None, collection=None, config=None,
                 disable_gpu: bool = True, device: Optional[Union[int, torch.device]] = None,
                 encode_fn: Optional[Callable] = None, index_root: Optional[str] = None,
                 query_batch: int = 64):
        # `disable_gpu` is part of the reference signature (default True) but dead there: colbert/searcher.py:23
        # never reads it, the device is chosen by config.total_visible_gpus (:40-43), which FLMR_executor.py:778-781
        # zeroes under DDP to force CPU search.  Here the search always runs on the GPU (no CPU path exists);
        # FlatCorpus raises when CUDA is absent.
        del disable_gpu
        self.config = config
        self.checkpoint = checkpoint
        self.collection = collection
        self.encode_fn = encode_fn
        self.query_batch = int(query_batch)
        if isinstance(index, FlatCorpus):
            self.corpus = index
            self.index = None
        else:
            root = index_root or (getattr(config, "index_root_", None) if config is not None else None)
            path = index if os.path.isabs(index) or root is None else os.path.join(root, index)
            tokens, doclens, meta = load_flat_index(path)
            self.index = path
            self.corpus = FlatCorpus(tokens, doclens, device=device)
        self._relu = False

    # -- reference-compatible surface ------------------------------------------------------------
    def configure(self, **kw):
        """colbert.Searcher.configure (searcher.py:49-50).  PLAID knobs (ncells,
        centroid_score_threshold, ndocs) are accepted and ignored: the scan is exhaustive."""
        self._relu = bool(kw.pop("relu", self._relu))
        return self

    def encode(self, text, full_length_search=False):
        if self.encode_fn is None:
            raise RuntimeError("text search needs encode_fn=...; the query encoder stays in PyTorch "
                               "(FLMR.query / Checkpoint.queryFromText) and is out of scope here")
        queries = text if isinstance(text, list) else [text]
        return self.encode_fn(queries)

    def search(self, text: str, k: int = 10, filter_fn=None):
        return self.dense_search(self.encode(text), k, filter_fn=filter_fn)

    def search_all(self, queries, k: int = 10, filter_fn=None):
        texts = list(queries.values()) if hasattr(queries, "values") else list(queries)
        return self._search_all_Q(queries, self.encode(texts), k, filter_fn=filter_fn)

    def _search_tensors(self, Q: torch.Tensor, k: int, filter_fn=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """(scores [B,k'], pids [B,k']) on the GPU, k' = min(k, n_passages)."""
        n = self.corpus.n_passages
        kk = min(int(k), n)
        if filter_fn is None and kk <= _cabi.MAX_K:
            outs, outp = [], []
            for b0 in range(0, Q.size(0), self.query_batch):
                s, p = maxsim_topk(self.corpus, Q[b0:b0 + self.query_batch], kk, relu=self._relu)
                outs.append(s)
                outp.append(p)
            return torch.cat(outs), torch.cat(outp)
        # filtered or very large k: all scores from the fused scan, selection as plain plumbing
        outs, outp = [], []
        keep = None
        if filter_fn is not None:
            all_pids = torch.arange(n, device=self.corpus.device) + self.corpus.pid_base
            keep = torch.as_tensor(filter_fn(all_pids), device=self.corpus.device).long() - self.corpus.pid_base
        for b0 in range(0, Q.size(0), self.query_batch):
            s = maxsim_scores(self.corpus, Q[b0:b0 + self.query_batch], relu=self._relu)
            if keep is not None:
                s = s[:, keep]
            kq = min(kk, s.size(1))
            if kq <= _cabi.SELECT_MAX_K:
                vals, idx = topk_select(s, kq)                       # radix-select kernel
            else:
                vals, idx = torch.sort(s, dim=1, descending=True, stable=True)
                vals, idx = vals[:, :kq], idx[:, :kq]
            pids = (keep[idx] if keep is not None else idx) + self.corpus.pid_base
            outs.append(vals)
            outp.append(pids)
        return torch.cat(outs), torch.cat(outp)

    def retrieve_and_rescore(self, Q: torch.Tensor, n_docs: int, generator: Optional[torch.Generator] = None):
        """The retrieval block of ``RagModelForBlip.main_retrieve`` (src/models/rag/rag_model_blip.py:388-443)
        in one call: search ``max(5, n_docs)`` passages per query (:392-397), keep ``n_docs`` of them (a random
        subset when ``n_docs < 5``, as :411-412 does with ``random.sample``), fetch their token embeddings and
        re-score them with the differentiable ``score`` so the gradient reaches the query encoder (:430-435).

        ``Q [B, Nq, d]`` stays on the GPU; the retrieved embeddings are gathered out of the resident corpus
        (no host dictionary, no H2D copy).  Returns a dict with ``doc_scores [B, n_docs]`` (differentiable
        w.r.t. ``Q``), ``retrieved_doc_ids`` (int64 numpy ``[B, n_docs]``, as :441), ``search_scores``, and the
        gathered ``item_embeddings [B, n_docs, Nd, d]`` / ``item_mask [B, n_docs, Nd, 1]``."""
        from .modeling import all_pairs_maxsim, colbert_score
        if Q.dim() != 3:
            raise ValueError("Q must be [B, Nq, d]")
        n_retrieve = max(5, int(n_docs))
        Qd = Q.to(self.corpus.device)
        s, p = self._search_tensors(Qd.detach(), n_retrieve)
        if s.size(1) < n_docs:
            raise ValueError("the corpus holds %d passages, fewer than n_docs=%d" % (s.size(1), n_docs))
        if s.size(1) != n_docs:
            pick = torch.rand(s.shape, generator=generator, device="cpu").argsort(dim=1)[:, :n_docs].to(s.device)
            s, p = s.gather(1, pick), p.gather(1, pick)
        D, mask = self.corpus.gather_padded(p)                               # [B, n_docs, Nd, d]
        B = Qd.size(0)
        if B <= 16:
            # one launch: every question against every retrieved passage, keep the block diagonal (the extra
            # pairs cost less than B separate launches; their upstream gradient is zero and is skipped)
            S = all_pairs_maxsim(Qd, D.flatten(0, 1), mask.flatten(0, 1))    # [B, B * n_docs]
            cols = torch.arange(B, device=S.device)[:, None] * n_docs + torch.arange(n_docs, device=S.device)
            doc_scores = S.gather(1, cols)
        else:
            doc_scores = torch.stack([
                colbert_score(Qd[b:b + 1].repeat_interleave(n_docs, dim=0), D[b], mask[b]) for b in range(B)])
        return {"doc_scores": doc_scores, "retrieved_doc_ids": p.cpu().numpy(), "search_scores": s,
                "item_embeddings": D, "item_mask": mask}

    def dense_search(self, Q: torch.Tensor, k: int = 10, filter_fn=None, remove_zero_tensors: bool = False):
        """searcher.py:91-132 -> ``(pids[:k], [1..k], scores[:k])`` for ONE query ``Q [1, Nq, d]``.

        ``remove_zero_tensors`` (searcher.py:120-126) is accepted for compatibility: all-zero query
        rows contribute exactly 0 to every passage here, so dropping them cannot change the result.
        """
        if Q.dim() == 2:
            Q = Q.unsqueeze(0)
        assert Q.size(0) == 1, "dense_search takes a single query (use _search_all_Q for batches)"
        s, p = self._search_tensors(Q, k, filter_fn)
        pids, scores = p[0].tolist(), s[0].tolist()
        return pids, list(range(1, len(pids) + 1)), scores

    def _search_all_Q(self, queries, Q: torch.Tensor, k: int, filter_fn=None, progress: bool = True,
                      remove_zero_tensors: bool = False) -> Ranking:
        """searcher.py:73-89, batched: one fused scan per ``query_batch`` queries instead of a Python
        loop of per-query ``dense_search`` calls."""
        keys = _query_keys(queries, Q.size(0))
        s, p = self._search_tensors(Q, k, filter_fn)
        s, p = s.cpu().tolist(), p.cpu().tolist()
        data = {qid: [(pid, rank + 1, score) for rank, (pid, score) in enumerate(zip(pp, ss))]
                for qid, pp, ss in zip(keys, p, s)}
        prov = {"source": "ravqa_b200.Searcher::search_all", "k": k,
                "n_passages": self.corpus.n_passages, "exhaustive": True}
        return Ranking(data=data, provenance=prov)
