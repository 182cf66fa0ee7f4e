"""``Searcher`` façade with the call surface of the reference's ``colbert.Searcher``
(third_party/ColBERT/colbert/searcher.py:22-132) so ``src/executors/FLMR_executor.py:774-792`` and
``src/models/rag/rag_model_blip.py:297-301, 397`` run unchanged:

    with Run().context(RunConfig(nranks=1, rank=..., root=..., experiment=...)):
        searcher = Searcher(index="temp_index.nbits=8", config=ColBERTConfig(total_visible_gpus=...))
    ._search_all_Q(queries, Q, k, filter_fn=None, progress=True, remove_zero_tensors=False) -> Ranking
    .dense_search(Q[1,Nq,d], k, filter_fn=None, remove_zero_tensors=False) -> (pids, ranks, scores)
    .search(text, k) / .search_all(queries, k)        (``encode_fn``, or the reference's ``Checkpoint`` built lazily
                                                       from ``checkpoint``: encoders stay PyTorch)
    Ranking.todict() -> {qid: [(pid, rank, score), ...]}   (colbert/data/ranking.py:48)

The index is addressed exactly as the reference does (``<Run().root>/<experiment>/indexes/<index>``,
infra.resolve_index_path) and whatever lives there is opened: a flat index written by this package's
``Indexer`` (index_io.py), or a PLAID directory written by the reference's own ``Indexer`` — decoded on the
GPU into the flat bf16 store (plaid.py).

What differs, by design: scoring is exhaustive and exact (no PLAID candidate
generation / centroid pruning), all queries of a call go through one batched fused scan instead of
a Python loop of per-query ``rank`` calls, and results always hold exactly ``min(k, n_passages)``
hits.  ``shard_across_ranks=True`` (an extension; needs an initialised ``torch.distributed`` group) keeps
only this rank's passage shard resident and merges the per-shard top-k with one all-gather, where the
reference repeats the whole search on every rank (FLMR_executor.py:778-781).
"""
from __future__ import annotations

import json
import os
from typing import Callable, Dict, List, Optional, Tuple, Union

import torch

from . import _cabi
from .corpus import FlatCorpus
from .index_io import FORMAT as FLAT_FORMAT
from .infra import ColBERTConfig, active_run_config, resolve_index_path
from .maxsim import maxsim_scores, maxsim_topk, topk_select


class Ranking:
    """colbert.data.Ranking (colbert/data/ranking.py:25-94): ``data`` maps qid -> [(pid, rank (1-based),
    score), ...]; ``flat_ranking`` / ``tolist`` is the same as (qid, pid, rank, score) rows; ``save`` writes
    the reference's TSV plus a ``.meta`` JSON with the provenance."""

    def __init__(self, path: Optional[str] = None, data=None, metrics=None, provenance=None):
        del metrics
        self._provenance = provenance if provenance is not None else (path or {})
        if data is None:
            if path is None:
                raise ValueError("Ranking needs data= or path=")
            data = self._load_file(path)
        if isinstance(data, dict):
            self.flat_ranking = [(qid, *rest) for qid, sub in data.items() for rest in sub]
            self.data = data
        else:                                       # flat rows: group by qid, order preserved
            self.flat_ranking = [tuple(r) for r in data]
            grouped: Dict = {}
            for qid, *rest in self.flat_ranking:
                grouped.setdefault(qid, []).append(tuple(rest))
            self.data = grouped

    @staticmethod
    def _load_file(path: str):
        def num(v):
            return float(v) if "." in v else int(v)
        with open(path) as f:
            return [list(map(num, line.strip().split("\t"))) for line in f if line.strip()]

    def provenance(self):
        return self._provenance

    def toDict(self):
        return {"provenance": self.provenance()}

    def todict(self):
        return dict(self.data)

    def tolist(self):
        return list(self.flat_ranking)

    def items(self):
        return self.data.items()

    def save(self, new_path: str) -> str:
        """ranking.py:64-82.  The reference resolves ``new_path`` under ``Run().path_`` (a per-script
        timestamped directory); here a relative path is taken as given."""
        assert "tsv" in new_path.strip("/").split("/")[-1].split("."), "TODO: Support .json[l] too."
        os.makedirs(os.path.dirname(os.path.abspath(new_path)), exist_ok=True)
        with open(new_path, "w") as f:
            for items in self.flat_ranking:
                f.write("\t".join(str(int(x) if type(x) is bool else x) for x in items) + "\n")
        prov = self.provenance()
        with open(new_path + ".meta", "w") as f:
            json.dump({"metadata": {}, "provenance": prov.toDict() if hasattr(prov, "toDict") else prov},
                      f, indent=4, default=str)
        return new_path

    @classmethod
    def cast(cls, obj):
        if type(obj) is str:
            return cls(path=obj)
        if isinstance(obj, (dict, list)):
            return cls(data=obj)
        if type(obj) is cls:
            return obj
        assert False, "obj has type %s which is not compatible with cast()" % type(obj)


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


def detect_index_kind(path: str) -> str:
    """'flat' (index_io.py) or 'plaid' (the reference's format) for the directory the
    reference's ``ColBERTConfig.load_from_index`` (base_config.py:71-87) would read its config from."""
    meta_path = os.path.join(path, "metadata.json")
    if os.path.exists(meta_path):
        with open(meta_path) as f:
            meta = json.load(f)
        if meta.get("format") == FLAT_FORMAT:
            return "flat"
        if "nbits" in meta.get("config", {}) and os.path.exists(os.path.join(path, "centroids.pt")):
            return "plaid"
        raise ValueError("%s holds a metadata.json that is neither a %s index nor a PLAID index "
                         "(config.nbits + centroids.pt)" % (path, FLAT_FORMAT))
    if os.path.exists(os.path.join(path, "plan.json")):
        raise ValueError("%s holds only plan.json: the PLAID index build did not finish "
                         "(collection_indexer.py:428-444 writes metadata.json last)" % path)
    raise FileNotFoundError("no index at %s (no metadata.json)" % path)


class Searcher:
    def __init__(self, index: Union[str, FlatCorpus], checkpoint=None, collection=None, config=None,
                 disable_gpu: bool = True, device: Optional[Union[int, torch.device]] = None,
                 encode_fn: Optional[Callable] = None, index_root: Optional[str] = None,
                 query_batch: int = 64, shard_across_ranks: bool = False, group=None):
        # `disable_gpu` is part of the reference signature (default True) but dead there: colbert/searcher.py:23
        # never reads it, the device is chosen by config.total_visible_gpus (:40-43), which FLMR_executor.py:778-781
        # zeroes under DDP to force CPU search.  Here the search always runs on the GPU (no CPU path exists);
        # FlatCorpus raises when CUDA is absent.
        del disable_gpu
        self.checkpoint = checkpoint
        self.collection = collection
        self.encode_fn = encode_fn
        self.query_batch = int(query_batch)
        self._sharded = None
        self._nccl = None
        self._checkpoint_model = None
        self.index_config = None
        run_cfg = active_run_config(config)
        if isinstance(index, FlatCorpus):
            self.corpus = index
            self.index = None
            self.index_kind = "resident"
        else:
            # searcher.py:26-30: index path = <index_root_ of from_existing(config, Run().config)>/<index>
            self.index = resolve_index_path(index, config, index_root)
            self.index_kind = detect_index_kind(self.index)
            self.index_config = ColBERTConfig.load_from_index(self.index) if self.index_kind == "plaid" else None
            rank, world = 0, 1
            if shard_across_ranks:
                import torch.distributed as dist
                if not (dist.is_available() and dist.is_initialized()):
                    raise RuntimeError("shard_across_ranks needs an initialised torch.distributed process group")
                rank, world = dist.get_rank(group), dist.get_world_size(group)
            if self.index_kind == "flat":
                self.corpus = FlatCorpus.from_index(self.index, device=device, rank=rank, world_size=world)
            else:
                self.corpus = FlatCorpus.from_plaid(self.index, device=device, rank=rank, world_size=world)
        # searcher.py:35: checkpoint config < index config < (config + Run().config); kept for callers that
        # read searcher.config — none of its PLAID knobs influence the exhaustive scan.
        cfg_type = type(config) if (config is not None and hasattr(type(config), "from_existing")) else ColBERTConfig
        try:
            self.config = cfg_type.from_existing(self.index_config if cfg_type is ColBERTConfig else None,
                                                 config, run_cfg)
        except Exception:          # foreign config type with a different constructor: keep what was given
            self.config = config
        if getattr(self.config, "interaction", "colbert") != "colbert":
            # the fused scan reduces with the plain sum of colbert_score_reduce (colbert.py:263); ranking a corpus
            # by the 'flipr' partial top-k sums (:248-261, unused by FLMR) is only available through colbert_score
            raise NotImplementedError("Searcher ranks with interaction='colbert'; got %r"
                                      % (self.config.interaction,))
        if shard_across_ranks:
            import torch.distributed as dist
            from .maxsim import topk_merge
            from .sharded import NcclExchange, ShardedSearcher
            self._sharded = ShardedSearcher(None, lambda s, p, k: topk_merge(s, p, k), group)
            self._group = group
            self._want_nccl = dist.get_backend(group) == "nccl"
        if self.corpus is not None:
            self.device = self.corpus.device
        else:
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self._sharded is not None and self._want_nccl:
            from .sharded import NcclExchange
            self._nccl = NcclExchange(self.device, self._group)       # collective: every rank constructs it
        self._relu = False

    # -- reference-compatible surface ------------------------------------------------------------
    def configure(self, **kw):
        """colbert.Searcher.configure (searcher.py:49-50).  PLAID knobs (ncells,
        centroid_score_threshold, ndocs) are recorded on ``config`` and otherwise ignored: the scan is
        exhaustive."""
        self._relu = bool(kw.pop("relu", self._relu))
        if kw and self.config is not None and hasattr(self.config, "configure"):
            self.config.configure(**kw)
        return self

    def _text_encoder(self):
        """The reference's own query encoder, built on first use as searcher.py:31-43 does at construction:
        ``Checkpoint(checkpoint or index_config.checkpoint, colbert_config=self.config)`` from the reference's
        ``colbert`` package (it stays PyTorch; nothing of it is reimplemented here)."""
        if self._checkpoint_model is None:
            name = self.checkpoint or getattr(self.index_config, "checkpoint", None)
            if name is None:
                raise RuntimeError("text search needs encode_fn=... or checkpoint=... (the query encoder stays in "
                                   "PyTorch: FLMR.query / Checkpoint.queryFromText, out of scope here)")
            try:
                from colbert.modeling.checkpoint import Checkpoint
            except ImportError as e:
                raise RuntimeError("text search with checkpoint=%r needs the reference's colbert package on "
                                   "sys.path (third_party/ColBERT), or pass encode_fn=..." % (name,)) from e
            model = Checkpoint(name, colbert_config=self.config)
            self._checkpoint_model = model.cuda() if torch.cuda.is_available() else model
        return self._checkpoint_model

    def encode(self, text, full_length_search=False):
        """searcher.py:52-59.  ``encode_fn`` (queries -> ``[n, Nq, d]``) when given, else the reference's
        ``Checkpoint.queryFromText``; the embeddings stay on the GPU (the reference moves them to the CPU)."""
        queries = text if isinstance(text, list) else [text]
        if self.encode_fn is not None:
            return self.encode_fn(queries)
        model = self._text_encoder()
        bsize = 128 if len(queries) > 128 else None
        maxlen = getattr(self.config, "query_maxlen", None)
        if maxlen is not None and hasattr(model, "query_tokenizer"):
            model.query_tokenizer.query_maxlen = maxlen
        return model.queryFromText(queries, bsize=bsize, to_cpu=False)

    def search(self, text: str, k: int = 10, filter_fn=None):
        return self.dense_search(self.encode(text), k, filter_fn=filter_fn)

    def search_all(self, queries, k: int = 10, filter_fn=None):
        texts = list(queries.values()) if hasattr(queries, "values") else list(queries)
        Q = self.encode(texts)
        if isinstance(Q, (list, tuple)):               # queryFromText with a batch size returns per-batch tensors
            Q = torch.cat(list(Q))
        return self._search_all_Q(queries, Q, k, filter_fn=filter_fn)

    @staticmethod
    def _drop_zero_rows(Q: torch.Tensor) -> torch.Tensor:
        """``remove_zero_tensors`` (searcher.py:120-126) for a batch: all-zero query rows are dropped and the
        remaining rows of every query packed to the front (order kept), padded with zero rows to the longest
        query of the batch.  A zero row adds ``max_j 0 = 0`` to every passage, so scores are unchanged; the
        scan just has fewer rows to go through (WIT pre-training queries: 32 live rows out of 32 + text)."""
        live = Q.abs().sum(dim=-1) > 0                                   # [B, Nq]
        n_live = int(live.sum(dim=1).max().item()) if Q.numel() else 0
        if n_live == Q.size(1):
            return Q
        order = torch.argsort((~live).to(torch.int8), dim=1, stable=True)[:, :max(n_live, 1)]
        return torch.gather(Q, 1, order.unsqueeze(-1).expand(-1, -1, Q.size(2)))

    def _local_topk(self, Q: torch.Tensor, kk: int, keep: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        """Top-``kk`` of this shard: (scores [B, kk'], global pids [B, kk']), kk' = min(kk, candidates)."""
        corpus = self.corpus
        if corpus is None:                               # this rank's shard is empty (more ranks than passages)
            return (torch.empty((Q.size(0), 0), dtype=torch.float32, device=self.device),
                    torch.empty((Q.size(0), 0), dtype=torch.int64, device=self.device))
        if keep is None and kk <= _cabi.MAX_K:
            kq = min(kk, corpus.n_passages)
            outs, outp = [], []
            for b0 in range(0, Q.size(0), self.query_batch):
                s, p = maxsim_topk(corpus, Q[b0:b0 + self.query_batch], kq, relu=self._relu)
                outs.append(s)
                outp.append(p)
            return torch.cat(outs), torch.cat(outp)
        # filtered or very large k: all scores from the fused scan, selection as plain plumbing
        outs, outp = [], []
        n_cand = corpus.n_passages if keep is None else int(keep.numel())
        kq = min(kk, n_cand)
        if kq == 0:
            return (torch.empty((Q.size(0), 0), dtype=torch.float32, device=corpus.device),
                    torch.empty((Q.size(0), 0), dtype=torch.int64, device=corpus.device))
        for b0 in range(0, Q.size(0), self.query_batch):
            s = maxsim_scores(corpus, Q[b0:b0 + self.query_batch], relu=self._relu)
            if keep is not None:
                s = s[:, keep]
            if kq <= _cabi.SELECT_MAX_K:
                vals, idx = topk_select(s, kq)                       # radix-select kernel
            else:
                vals, idx = torch.sort(s, dim=1, descending=True, stable=True)
                vals, idx = vals[:, :kq], idx[:, :kq]
            outs.append(vals)
            outp.append((keep[idx] if keep is not None else idx) + corpus.pid_base)
        return torch.cat(outs), torch.cat(outp)

    def _search_tensors(self, Q: torch.Tensor, k: int, filter_fn=None,
                        remove_zero_tensors: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """(scores [B,k'], pids [B,k']) on the GPU, k' = min(k, passages that pass ``filter_fn``)."""
        if Q.dim() == 2:
            Q = Q.unsqueeze(0)
        if remove_zero_tensors:
            Q = self._drop_zero_rows(Q)
        corpus = self.corpus
        keep = None
        if filter_fn is not None and corpus is not None:
            all_pids = torch.arange(corpus.n_passages, device=corpus.device) + corpus.pid_base
            keep = torch.as_tensor(filter_fn(all_pids), device=corpus.device).long() - corpus.pid_base
        if self._sharded is not None and int(k) > _cabi.MAX_K:
            raise ValueError("sharded search merges at most k=%d results per query" % _cabi.MAX_K)
        if self._nccl is not None and keep is None and corpus is not None:
            # the C-level sharded search: scan + one grouped all-gather + merge per call, no torch glue
            outs = [self._nccl.search(corpus, Q[b0:b0 + self.query_batch], int(k), relu=self._relu)
                    for b0 in range(0, Q.size(0), self.query_batch)]
            s, p = torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])
            n_valid = int((p[0] >= 0).sum()) if p.numel() else 0
            return s[:, :n_valid], p[:, :n_valid]
        s, p = self._local_topk(Q, int(k), keep)
        if self._sharded is not None:
            k_merge = int(k)
            if s.size(1) < k_merge:                      # short shard: pad with empty entries
                s = torch.cat([s, s.new_full((s.size(0), k_merge - s.size(1)), float("-inf"))], dim=1)
                p = torch.cat([p, p.new_full((p.size(0), k_merge - p.size(1)), -1)], dim=1)
            if self._nccl is not None:
                s, p = self._nccl.exchange(s, p, k_merge)
            else:
                s, p = self._sharded.exchange(s.to(self.device), p.to(self.device), k_merge)
            n_valid = int((p[0] >= 0).sum()) if p.numel() else 0    # same for every query: min(k, candidates)
            s, p = s[:, :n_valid], p[:, :n_valid]
        return s, p

    def retrieve_and_rescore(self, Q: torch.Tensor, n_docs: int, generator: Optional[torch.Generator] = None):
        """The retrieval block of ``RagModelForBlip.main_retrieve`` (src/models/rag/rag_model_blip.py:388-443)
        in one call: search ``max(5, n_docs)`` passages per query (:392-397), keep ``n_docs`` of them (a random
        subset when ``n_docs < 5``, as :411-412 does with ``random.sample``), fetch their token embeddings and
        re-score them with the differentiable ``score`` so the gradient reaches the query encoder (:430-435).

        ``Q [B, Nq, d]`` stays on the GPU; the retrieved embeddings are gathered out of the resident corpus
        (no host dictionary, no H2D copy).  Returns a dict with ``doc_scores [B, n_docs]`` (differentiable
        w.r.t. ``Q``), ``retrieved_doc_ids`` (int64 numpy ``[B, n_docs]``, as :441), ``search_scores``, and the
        gathered ``item_embeddings [B, n_docs, Nd, d]`` / ``item_mask [B, n_docs, Nd, 1]``."""
        from .modeling import grouped_maxsim
        if Q.dim() != 3:
            raise ValueError("Q must be [B, Nq, d]")
        if self._sharded is not None:
            raise RuntimeError("retrieve_and_rescore gathers embeddings from the local shard only; "
                               "use an unsharded Searcher for the RAG block")
        n_retrieve = max(5, int(n_docs))
        Qd = Q.to(self.corpus.device)
        s, p = self._search_tensors(Qd.detach(), n_retrieve)
        if s.size(1) < n_docs:
            raise ValueError("the corpus holds %d passages, fewer than n_docs=%d" % (s.size(1), n_docs))
        if s.size(1) != n_docs:
            pick = torch.rand(s.shape, generator=generator, device="cpu").argsort(dim=1)[:, :n_docs].to(s.device)
            s, p = s.gather(1, pick), p.gather(1, pick)
        D, mask = self.corpus.gather_padded(p)                               # [B, n_docs, Nd, d]
        # one block-diagonal launch: question b against ITS n_docs passages only
        doc_scores = grouped_maxsim(Qd, D.flatten(0, 1), mask.flatten(0, 1), n_docs)
        return {"doc_scores": doc_scores, "retrieved_doc_ids": p.cpu().numpy(), "search_scores": s,
                "item_embeddings": D, "item_mask": mask}

    def _fill_search_knobs(self, k: int) -> None:
        """searcher.py:92-118: the first search fills the PLAID knobs still unset on ``config`` from ``k`` (and later
        searches keep them).  They steer nothing here — the scan is exhaustive — but callers that read
        ``searcher.config.ncells`` etc. after a search find what the reference would have left there."""
        cfg = self.config
        if cfg is None or not hasattr(cfg, "configure") or not hasattr(cfg, "ncells"):
            return
        ncells, thr, ndocs = (2, 0.45, 1024) if k <= 100 else (4, 0.4, max(k * 4, 4096))
        if cfg.ncells is None:
            cfg.configure(ncells=ncells)
        if cfg.centroid_score_threshold is None:
            cfg.configure(centroid_score_threshold=thr)
        if cfg.ndocs is None:
            cfg.configure(ndocs=ndocs)

    def dense_search(self, Q: torch.Tensor, k: int = 10, filter_fn=None, remove_zero_tensors: bool = False):
        """searcher.py:91-132 -> ``(pids[:k], [1..k], scores[:k])`` for ONE query ``Q [1, Nq, d]``."""
        if Q.dim() == 2:
            Q = Q.unsqueeze(0)
        assert Q.size(0) == 1, "dense_search takes a single query (use _search_all_Q for batches)"
        self._fill_search_knobs(k)
        s, p = self._search_tensors(Q, k, filter_fn, remove_zero_tensors)
        pids, scores = p[0].tolist(), s[0].tolist()
        return pids, list(range(1, len(pids) + 1)), scores

    def _search_all_Q(self, queries, Q: torch.Tensor, k: int, filter_fn=None, progress: bool = True,
                      remove_zero_tensors: bool = False) -> Ranking:
        """searcher.py:73-89, batched: one fused scan per ``query_batch`` queries instead of a Python
        loop of per-query ``dense_search`` calls."""
        del progress
        self._fill_search_knobs(k)
        keys = _query_keys(queries, Q.size(0))
        s, p = self._search_tensors(Q, k, filter_fn, remove_zero_tensors)
        s, p = s.cpu().tolist(), p.cpu().tolist()
        data = {qid: [(pid, rank + 1, score) for rank, (pid, score) in enumerate(zip(pp, ss))]
                for qid, pp, ss in zip(keys, p, s)}
        prov = {"source": "Searcher::search_all", "k": k,
                "queries": queries.provenance() if hasattr(queries, "provenance") else None,
                "config": self.config.export() if hasattr(self.config, "export") else None,
                "backend": "ravqa_b200 exhaustive fused scan", "index": self.index,
                "index_kind": self.index_kind, "exhaustive": True,
                "n_passages": self.corpus.n_passages if self.corpus is not None else 0}
        return Ranking(data=data, provenance=prov)
