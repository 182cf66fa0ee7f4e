"""Scoring surface of the reference's models on the CUDA path:

    colbert_score(Q, D_padded, D_mask)          third_party/ColBERT/colbert/modeling/colbert.py:268-286
    ColBERT.score / FLMR*.score                 colbert.py:217-224  (callers: colbert.py:71-73 training,
                                                src/models/rag/rag_model_blip.py:432-435 RAG re-score,
                                                src/executors/FLMR_executor.py:828-833 exhaustive eval)
    compute_ib_loss_new (in-batch negatives)    colbert.py:82-113

Forward, training-sized batches (the usual case here): ONE launch of ``flmr_maxsim_argmax`` computes, for
every (query token, document) pair, the winning document token and its inner product — the row maxima
summed over query tokens are the all-pairs scores ``[B, n]`` (the aligned ``Q_dup`` form the reference
builds with ``repeat_interleave`` is a gather of that matrix, the in-batch-negatives matrix IS that
matrix), and the winners (4 bytes per pair) are what the backward needs.  Large inputs (exhaustive
evaluation through ``score``): the ``n`` padded documents are packed into a temporary FlatCorpus and
scored by one launch of the tcgen05 scan kernel; the backward then recomputes the winners.  Either way the
``[n, Nd, Nq]`` score tensor the reference materialises (218 MB per rank at C4) never exists.

Backward: ``flmr_maxsim_backward`` routes the gradient through the winners (``dQ_i += g * D_argmax``
gathered, ``dD_argmax += g * Q_i`` scattered with fp32 atomics).  Both are CUDA kernels behind the C ABI
(csrc/flmr_train_kernels.cuh).  Inputs are rounded to bf16 for forward and backward, like the reference's
fp16 GPU path (colbert.py:205-206).
"""
from __future__ import annotations

from typing import Optional

import torch

from .corpus import FlatCorpus
from .maxsim import maxsim_argmax, maxsim_backward, maxsim_scores


# Below this many multiply-accumulates the whole forward is ONE launch of the arg-max kernel (scores =
# row maxima summed), which also saves the winners for the backward; above it the tcgen05 scan kernel over
# a temporary packed corpus wins despite its per-call setup (allocation, partition build, TMA descriptor:
# ~2.3 ms measured against ~5e13 MAC/s of the warp-MMA kernel, profiles/r01_train_step_probe.md) and the backward recomputes the winners.
_FUSED_SMALL_MAX_MACS = 1e11


def _forward_scores(Q: torch.Tensor, D_padded: torch.Tensor, D_mask: torch.Tensor):
    """All-pairs scores ``[B, n]``; returns (scores, bool mask [n, Nd], saved arg-max or None)."""
    if not (Q.is_cuda and D_padded.is_cuda):
        raise RuntimeError("the scoring path is CUDA-only (no CPU fallback)")
    n, nd = D_padded.size(0), D_padded.size(1)
    mask = D_mask.reshape(n, nd).bool()
    if not bool(mask.any(dim=1).all()):
        raise ValueError("a document has no unmasked token: its MaxSim score is undefined "
                         "(-9999 * Nq on the reference's padded path)")
    if float(Q.size(0)) * n * Q.size(1) * nd * Q.size(2) <= _FUSED_SMALL_MAX_MACS:
        arg, rowmax = maxsim_argmax(Q, D_padded, mask, return_rowmax=True)
        return rowmax.sum(dim=-1), mask, arg
    packed = D_padded.detach()[mask]
    corpus = FlatCorpus(packed.to(torch.bfloat16), mask.sum(dim=1).cpu(), device=Q.device, adopt=True)
    try:
        scores = maxsim_scores(corpus, Q.detach())
        torch.cuda.current_stream(Q.device).synchronize()   # corpus buffers die with this scope
    finally:
        corpus.close()
    return scores, mask, None


class _AllPairsMaxSim(torch.autograd.Function):
    """scores[b, p] = sum_i max_{j: mask[p, j]} <Q[b, i], D[p, j]>  for all (b, p)."""

    @staticmethod
    def forward(ctx, Q: torch.Tensor, D_padded: torch.Tensor, D_mask: torch.Tensor) -> torch.Tensor:
        scores, mask, arg = _forward_scores(Q, D_padded, D_mask)
        ctx.has_arg = arg is not None
        ctx.save_for_backward(Q, D_padded, mask, *([arg] if ctx.has_arg else []))
        return scores

    @staticmethod
    def backward(ctx, grad: torch.Tensor):
        Q, D_padded, mask = ctx.saved_tensors[:3]
        need_dq, need_dd = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_dq or need_dd):
            return None, None, None
        arg = ctx.saved_tensors[3] if ctx.has_arg else maxsim_argmax(Q, D_padded, mask)
        dQ, dD = maxsim_backward(Q, D_padded, arg, grad, need_dq=need_dq, need_dd=need_dd)
        return (dQ.to(Q.dtype) if dQ is not None else None,
                dD.to(D_padded.dtype) if dD is not None else None, None)


def _consecutive_runs(Q: torch.Tensor):
    """Runs of identical consecutive queries: (one representative per run ``[U, Nq, d]``, run index of every
    row ``[n]``).  Same result as ``torch.unique_consecutive(Q, dim=0, return_inverse=True)``, whose CUDA
    implementation compares whole rows (Nq*d elements) serially in a single thread — ~80 ms per call for an
    832-token query, measured — where two vectorised passes take microseconds."""
    n = Q.size(0)
    if n == 1:
        return Q, torch.zeros(1, dtype=torch.long, device=Q.device)
    new_run = torch.ones(n, dtype=torch.bool, device=Q.device)
    new_run[1:] = (Q[1:] != Q[:-1]).flatten(1).any(dim=1)
    inverse = torch.cumsum(new_run.long(), dim=0) - 1
    return Q[new_run], inverse


class _AlignedMaxSim(torch.autograd.Function):
    """scores[p] = MaxSim(Q[p], D[p]) for ``Q [n, Nq, d]`` aligned with ``D [n, Nd, d]``.

    Callers build ``Q`` with ``repeat_interleave`` (colbert.py:71, rag_model_blip.py:433,
    FLMR_executor.py:828): runs of identical queries are scored once against all documents (one scan
    launch over the unique queries) and the aligned entries are gathered."""

    @staticmethod
    def forward(ctx, Q: torch.Tensor, D_padded: torch.Tensor, D_mask: torch.Tensor) -> torch.Tensor:
        Qu, inverse = _consecutive_runs(Q.detach())
        S, mask, arg = _forward_scores(Qu, D_padded, D_mask)
        cols = torch.arange(D_padded.size(0), device=S.device)
        ctx.has_arg = arg is not None
        ctx.save_for_backward(Qu, D_padded, mask, inverse, *([arg] if ctx.has_arg else []))
        ctx.q_dtype = Q.dtype
        return S[inverse, cols]

    @staticmethod
    def backward(ctx, grad: torch.Tensor):
        Qu, D_padded, mask, inverse = ctx.saved_tensors[:4]
        need_dq, need_dd = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_dq or need_dd):
            return None, None, None
        n = D_padded.size(0)
        cols = torch.arange(n, device=grad.device)
        arg = ctx.saved_tensors[4] if ctx.has_arg else maxsim_argmax(Qu, D_padded, mask)   # [U, n, Nq]
        dQ = dD = None
        if need_dd:
            G = torch.zeros((Qu.size(0), n), dtype=torch.float32, device=grad.device)
            G[inverse, cols] = grad.float()
            dD = maxsim_backward(Qu, D_padded, arg, G, need_dq=False, need_dd=True)[1].to(D_padded.dtype)
        if need_dq:                                                            # row p: grad[p] * D[p, arg[p]]
            idx = arg[inverse, cols].long().clamp_min(0)                       # [n, Nq]
            Dsel = D_padded.detach().to(torch.bfloat16).float().gather(
                1, idx.unsqueeze(-1).expand(-1, -1, D_padded.size(2)))
            dQ = (grad.float()[:, None, None] * Dsel).to(ctx.q_dtype)
        return dQ, dD, None


def all_pairs_maxsim(Q: torch.Tensor, D_padded: torch.Tensor, D_mask: torch.Tensor) -> torch.Tensor:
    """``[B, n]`` MaxSim of every query against every padded document (differentiable)."""
    return _AllPairsMaxSim.apply(Q, D_padded, D_mask)


def colbert_score(Q: torch.Tensor, D_padded: torch.Tensor, D_mask: torch.Tensor, config=None,
                  use_gpu: bool = True) -> torch.Tensor:
    """Drop-in for colbert.modeling.colbert.colbert_score (colbert.py:268-286).

    ``Q.size(0)`` is 1 (compare with all documents) or ``n`` (each query against its aligned
    document — callers build it with ``repeat_interleave``).  Returns ``[n]`` scores, true-max
    semantics of the padded path."""
    assert Q.dim() == 3 and D_padded.dim() == 3, (Q.size(), D_padded.size())
    assert Q.size(0) in [1, D_padded.size(0)]
    if not use_gpu and not Q.is_cuda:
        raise RuntimeError("this colbert_score is the CUDA path; there is no CPU fallback")
    dev = Q.device if Q.is_cuda else torch.device("cuda", torch.cuda.current_device())
    Q, D_padded, D_mask = Q.to(dev), D_padded.to(dev), D_mask.to(dev)
    if Q.size(0) == 1:
        return all_pairs_maxsim(Q, D_padded, D_mask)[0]
    return _AlignedMaxSim.apply(Q, D_padded, D_mask)


class _GatherCat(torch.autograd.Function):
    """``cat(all_gather(x))`` along dim 0 whose backward returns every rank's gradient to the owner of the
    rows (sum over ranks of the owner's slice): the differentiable form of the gather the reference sketches
    and leaves disabled (colbert.py:68-69, 115-163 — there remote rows are detached, so a document never
    receives the gradient of another rank's queries)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, group):
        import torch.distributed as dist
        world = dist.get_world_size(group)
        ctx.group, ctx.rank, ctx.rows = group, dist.get_rank(group), x.size(0)
        outs = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(outs, x.contiguous(), group=group)
        return torch.cat(outs, dim=0)

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        import torch.distributed as dist
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g[ctx.rank * ctx.rows:(ctx.rank + 1) * ctx.rows], None


def gather_documents(D_padded: torch.Tensor, D_mask: torch.Tensor, group=None):
    """Every rank's documents (and masks), concatenated in rank order, differentiable w.r.t. the local rows.
    Ranks may hold different padded lengths; all are padded (masked out) to the longest."""
    import torch.distributed as dist
    n, nd = D_padded.size(0), D_padded.size(1)
    nd_max = torch.tensor([nd], dtype=torch.int64, device=D_padded.device)
    dist.all_reduce(nd_max, op=dist.ReduceOp.MAX, group=group)
    pad = int(nd_max.item()) - nd
    mask = D_mask.reshape(n, nd).to(torch.uint8)
    if pad:
        D_padded = torch.nn.functional.pad(D_padded, (0, 0, 0, pad))
        mask = torch.nn.functional.pad(mask, (0, pad))
    return _GatherCat.apply(D_padded, group), _GatherCat.apply(mask, group).bool().unsqueeze(-1)


def in_batch_negatives_loss(Q: torch.Tensor, D_padded: torch.Tensor, D_mask: torch.Tensor, nway: int,
                            return_scores: bool = False, cross_rank_negatives: bool = False, group=None,
                            all_pairs_fn=None):
    """compute_ib_loss_new (colbert.py:82-113): ``Q [B, Nq, d]``, ``D [B*nway, Nd, d]`` with the positive
    of query i at row ``i*nway`` (colbert.py:103-108); cross-entropy over all ``B*nway`` documents.

    ``cross_rank_negatives=True`` (needs an initialised process group; every rank the same B and nway):
    this rank's queries are scored against the documents of ALL ranks — ``[B, world*B*nway]``, the positive
    of local query i at column ``rank*B*nway + i*nway`` — one all-gather of the document embeddings forward,
    one all-reduce of their gradient backward.  The mean over ranks of the returned losses is the
    cross-entropy of the global batch, and its gradient reaches every document from every rank's queries.
    ``all_pairs_fn`` replaces the CUDA scorer (host-logic tests only)."""
    score_fn = all_pairs_maxsim if all_pairs_fn is None else all_pairs_fn
    B = Q.size(0)
    assert D_padded.size(0) == B * nway, (D_padded.size(), B, nway)
    first = 0
    if cross_rank_negatives:
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("cross_rank_negatives needs an initialised torch.distributed process group")
        first = dist.get_rank(group) * B * nway
        D_padded, D_mask = gather_documents(D_padded, D_mask, group)
    scores = score_fn(Q, D_padded, D_mask)                                # [B, (world*)B*nway]
    labels = first + torch.arange(B, device=scores.device) * nway
    loss = torch.nn.functional.cross_entropy(scores, labels)
    return (loss, scores) if return_scores else loss


class FLMRModelForRetrieval(torch.nn.Module):
    """Façade with the name the reference announces for its HF API (README.md:25) and the call surface
    of the in-repo ``FLMR*`` classes (src/models/retriever/FLMR.py): ``query`` / ``doc`` delegate to the
    wrapped PyTorch encoders (out of scope here), ``score`` / ``forward`` route into the CUDA path."""

    def __init__(self, query_encoder: Optional[torch.nn.Module] = None,
                 doc_encoder: Optional[torch.nn.Module] = None, nway: int = 2, use_ib_negatives: bool = True,
                 cross_rank_negatives: bool = False):
        super().__init__()
        self.cross_rank_negatives = cross_rank_negatives
        self.query_encoder = query_encoder
        self.doc_encoder = doc_encoder
        self.nway = nway
        self.use_ib_negatives = use_ib_negatives

    def query(self, *args, **kw):
        if self.query_encoder is None:
            raise RuntimeError("no query encoder wrapped: FLMR.query stays in PyTorch (SURVEY 8a, a12)")
        return self.query_encoder(*args, **kw)

    def doc(self, *args, **kw):
        if self.doc_encoder is None:
            raise RuntimeError("no document encoder wrapped: ColBERT.doc stays in PyTorch")
        return self.doc_encoder(*args, **kw)

    def score(self, Q: torch.Tensor, D_padded: torch.Tensor, D_mask: torch.Tensor) -> torch.Tensor:
        """ColBERT.score (colbert.py:217-224), similarity == 'cosine'."""
        return colbert_score(Q, D_padded, D_mask)

    def forward(self, Q: torch.Tensor, D: torch.Tensor, D_mask: torch.Tensor):
        """ColBERT.forward on pre-computed embeddings (colbert.py:64-80): ``Q [B, Nq, d]``,
        ``D [B*nway, Nd, d]`` -> (scores [B*nway] of the aligned pairs, ib_loss)."""
        loss, S = in_batch_negatives_loss(Q, D, D_mask, self.nway, return_scores=True,
                                          cross_rank_negatives=self.cross_rank_negatives)
        B = Q.size(0)
        first = 0
        if self.cross_rank_negatives:
            first = torch.distributed.get_rank() * B * self.nway
        rows = torch.arange(B, device=S.device).repeat_interleave(self.nway)
        aligned = S[rows, first + torch.arange(B * self.nway, device=S.device)]
        return (aligned, loss) if self.use_ib_negatives else aligned
