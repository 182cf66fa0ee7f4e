"""Scoring surface of the reference's models on the CUDA path:

    colbert_score(Q, D_padded, D_mask)          third_party/ColBERT/colbert/modeling/colbert.py:268-286
    ColBERT.score / FLMR*.score                 colbert.py:217-224  (callers: colbert.py:71-73 training,
                                                src/models/rag/rag_model_blip.py:432-435 RAG re-score,
                                                src/executors/FLMR_executor.py:828-833 exhaustive eval)
    compute_ib_loss_new (in-batch negatives)    colbert.py:82-113

The forward is ONE launch of the fused scan kernel: the ``n`` padded documents are packed into a
temporary FlatCorpus and every query is scored against every document (all pairs ``[B, n]``) — the
aligned ``Q_dup`` form the reference builds with ``repeat_interleave`` is a gather of that matrix,
and the in-batch-negatives matrix IS that matrix.  The ``[n, Nd, Nq]`` score tensor the reference
materialises (218 MB per rank at C4) never exists.

Backward (first version): the arg-max token of every (query token, document) pair is recomputed in
chunks with plain torch ops and the gradients are scattered (``dQ_i += g * D_argmax``,
``dD_argmax += g * Q_i``); a fused arg-max-saving kernel is round-2 work.
Inputs are rounded to bf16 for the forward, like the reference's fp16 GPU path (colbert.py:205-206).
"""
from __future__ import annotations

from typing import Optional

import torch

from .corpus import FlatCorpus
from .maxsim import maxsim_scores


def _pack(D_padded: torch.Tensor, D_mask: torch.Tensor):
    n, nd, d = D_padded.shape
    mask = D_mask.reshape(n, nd).bool()
    doclens = mask.sum(dim=1)
    return D_padded[mask], doclens, mask


class _AllPairsMaxSim(torch.autograd.Function):
    """scores[b, p] = sum_i max_{j: mask[p, j]} <Q[b, i], D[p, j]>  for all (b, p)."""

    @staticmethod
    def forward(ctx, Q: torch.Tensor, D_padded: torch.Tensor, D_mask: torch.Tensor) -> torch.Tensor:
        if not Q.is_cuda:
            raise RuntimeError("the scoring path is CUDA-only (no CPU fallback)")
        packed, doclens, mask = _pack(D_padded.detach(), D_mask)
        if int(doclens.min()) < 1:
            raise ValueError("a document has no unmasked token: its MaxSim score is undefined "
                             "(-9999 * Nq on the reference's padded path)")
        corpus = FlatCorpus(packed.to(torch.bfloat16), doclens.cpu(), device=Q.device, adopt=True)
        try:
            scores = maxsim_scores(corpus, Q.detach())
            torch.cuda.current_stream(Q.device).synchronize()   # corpus buffers die with this scope
        finally:
            corpus.close()
        ctx.save_for_backward(Q, D_padded, mask)
        return scores

    @staticmethod
    def backward(ctx, grad: torch.Tensor):
        Q, D_padded, mask = ctx.saved_tensors
        B, nq, d = Q.shape
        n, nd, _ = D_padded.shape
        Qf = Q.detach().to(torch.bfloat16).float()
        Df = D_padded.detach().to(torch.bfloat16).float()
        grad = grad.float()
        dQ = torch.zeros_like(Qf) if ctx.needs_input_grad[0] else None
        dD = torch.zeros_like(Df) if ctx.needs_input_grad[1] else None
        chunk = max(1, int((256 << 20) // max(1, B * nq * nd * 4)))      # ~256 MB of scores per chunk
        for p0 in range(0, n, chunk):
            p1 = min(n, p0 + chunk)
            S = torch.einsum("bqd,pkd->bpqk", Qf, Df[p0:p1])              # [B, c, nq, nd]
            S = S.masked_fill(~mask[p0:p1, None, :].unsqueeze(0), float("-inf"))
            arg = S.argmax(dim=-1)                                        # [B, c, nq]
            del S
            g = grad[:, p0:p1]                                            # [B, c]
            Dsel = Df[p0:p1]                                              # [c, nd, d]
            if dQ is not None:
                gathered = Dsel[torch.arange(p1 - p0, device=arg.device)[None, :, None], arg]   # [B, c, nq, d]
                dQ += (g[:, :, None, None] * gathered).sum(dim=1)
            if dD is not None:
                contrib = g[:, :, None, None] * Qf[:, None, :, :]        # [B, c, nq, d]
                flat_idx = (torch.arange(p1 - p0, device=arg.device)[None, :, None] * nd + arg).reshape(-1)
                dD[p0:p1].view(-1, d).index_add_(0, flat_idx, contrib.expand(B, p1 - p0, nq, d).reshape(-1, d))
        return (dQ.to(Q.dtype) if dQ is not None else None,
                dD.to(D_padded.dtype) if dD is not None else None, None)


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
    S = all_pairs_maxsim(Q, D_padded, D_mask)
    if Q.size(0) == 1:
        return S[0]
    return S.diagonal()


def in_batch_negatives_loss(Q: torch.Tensor, D_padded: torch.Tensor, D_mask: torch.Tensor, nway: int,
                            return_scores: bool = False):
    """compute_ib_loss_new (colbert.py:82-113): ``Q [B, Nq, d]``, ``D [B*nway, Nd, d]`` with the positive
    of query i at row ``i*nway`` (colbert.py:103-108); cross-entropy over all ``B*nway`` documents."""
    B = Q.size(0)
    assert D_padded.size(0) == B * nway, (D_padded.size(), B, nway)
    scores = all_pairs_maxsim(Q, D_padded, D_mask)                        # [B, B*nway]
    labels = torch.arange(B, device=scores.device) * nway
    loss = torch.nn.functional.cross_entropy(scores, labels)
    return (loss, scores) if return_scores else loss


class FLMRModelForRetrieval(torch.nn.Module):
    """Façade with the name the reference announces for its HF API (README.md:25) and the call surface
    of the in-repo ``FLMR*`` classes (src/models/retriever/FLMR.py): ``query`` / ``doc`` delegate to the
    wrapped PyTorch encoders (out of scope here), ``score`` / ``forward`` route into the CUDA path."""

    def __init__(self, query_encoder: Optional[torch.nn.Module] = None,
                 doc_encoder: Optional[torch.nn.Module] = None, nway: int = 2, use_ib_negatives: bool = True):
        super().__init__()
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
        loss, S = in_batch_negatives_loss(Q, D, D_mask, self.nway, return_scores=True)
        B = Q.size(0)
        rows = torch.arange(B, device=S.device).repeat_interleave(self.nway)
        aligned = S[rows, torch.arange(B * self.nway, device=S.device)]
        return (aligned, loss) if self.use_ib_negatives else aligned
