"""TEST INFRASTRUCTURE: substitutes the lowest layer of the product (the ctypes calls into the CUDA library in
maxsim.py / corpus.py) by the numpy oracle, so that the host layer above it — Searcher, index addressing, PLAID
detection, the autograd functions, integration.patch_colbert — can be driven on a machine without a GPU by the
reference's own classes (tests/test_reference_callsites.py).  Nothing here is imported by the product package;
on the GPU box the same call sites run against the real kernels (tests/test_callsites_gpu.py).
"""
from __future__ import annotations

import collections
import json
import os

import numpy as np
import torch

from oracle import maxsim_oracle as O


def _bf16(x: torch.Tensor) -> np.ndarray:
    """fp32 view of the operands as the kernels see them (rounded to bf16)."""
    return x.detach().to(torch.bfloat16).float().cpu().numpy()


class OracleCorpus:
    """Stands in for corpus.FlatCorpus: same attributes the Searcher touches, numpy storage."""

    def __init__(self, tokens, doclens, device=None, pid_base: int = 0, adopt: bool = True):
        self.tokens = _bf16(torch.as_tensor(tokens))
        self.doclens = np.asarray(doclens, dtype=np.int32)
        self.pid_base = int(pid_base)
        self.device = torch.device("cpu")

    @property
    def n_passages(self):
        return len(self.doclens)

    @classmethod
    def from_plaid(cls, path, device=None, rank=0, world_size=1):
        assert world_size == 1
        with open(os.path.join(path, "metadata.json")) as f:
            meta = json.load(f)
        nbits = int(meta["config"]["nbits"])
        centroids = torch.load(os.path.join(path, "centroids.pt"), map_location="cpu").float().numpy()
        weights = torch.load(os.path.join(path, "buckets.pt"), map_location="cpu")[1].float().numpy()
        toks, dls = [], []
        for c in range(int(meta["num_chunks"])):
            codes = torch.load(os.path.join(path, "%d.codes.pt" % c), map_location="cpu").numpy()
            res = torch.load(os.path.join(path, "%d.residuals.pt" % c), map_location="cpu").numpy()
            toks.append(O.plaid_decode(codes, res, centroids, weights, nbits))
            with open(os.path.join(path, "doclens.%d.json" % c)) as f:
                dls.extend(json.load(f))
        return cls(torch.from_numpy(np.concatenate(toks)), dls)

    @classmethod
    def from_index(cls, path, device=None, rank=0, world_size=1):
        from ravqa_b200.index_io import load_flat_index
        tokens, doclens, _ = load_flat_index(path)
        return cls(tokens, doclens)

    def close(self):
        pass


def install(monkeypatch):
    """Patch the product's lowest layer with oracle-backed functions; returns a Counter of how often each
    substituted entry point ran (tests use it to prove which path a reference method took)."""
    import ravqa_b200.modeling as M
    import ravqa_b200.searcher as S
    calls = collections.Counter()

    def maxsim_scores(corpus, Q, relu=False, out=None):
        calls["scores"] += 1
        return torch.from_numpy(O.maxsim_scores(_bf16(Q), corpus.tokens, corpus.doclens, relu=relu))

    def maxsim_topk(corpus, Q, k, relu=False):
        calls["topk"] += 1
        s, p = O.topk(O.maxsim_scores(_bf16(Q), corpus.tokens, corpus.doclens, relu=relu), k, corpus.pid_base)
        return torch.from_numpy(s), torch.from_numpy(p)

    def topk_select(scores, k, pid_base=0):
        calls["select"] += 1
        s, p = O.topk(scores.numpy(), k, pid_base)
        return torch.from_numpy(s), torch.from_numpy(p)

    def _pairs(Q, D, mask, stride):
        """winners + maxima for query b against documents [b*stride, b*stride + n_per)."""
        Qn, Dn, m = _bf16(Q), _bf16(D), mask.reshape(D.size(0), D.size(1)).bool().cpu().numpy()
        B = Qn.shape[0]
        n_per = Dn.shape[0] if stride == 0 else stride
        arg = np.full((B, n_per, Qn.shape[1]), -1, dtype=np.int32)
        rowmax = np.full((B, n_per, Qn.shape[1]), -np.inf, dtype=np.float32)
        for b in range(B):
            for t in range(n_per):
                p = b * stride + t
                if not m[p].any():
                    continue
                sc = Qn[b] @ Dn[p].T
                sc[:, ~m[p]] = -np.inf
                arg[b, t] = sc.argmax(axis=1)
                rowmax[b, t] = sc.max(axis=1)
        return torch.from_numpy(arg), torch.from_numpy(rowmax)

    def _grads(Q, D, arg, grad, stride, need_dq, need_dd):
        Qn, Dn, a, g = _bf16(Q), _bf16(D), arg.numpy(), grad.detach().float().numpy()
        dQ, dD = np.zeros_like(Qn), np.zeros_like(Dn)
        for b in range(a.shape[0]):
            for t in range(a.shape[1]):
                p = b * stride + t
                j = a[b, t]
                live = j >= 0                    # per token, like the kernels (fully masked document, or a token
                if g[b, t] == 0.0 or not live.any():   # the 'flipr' reduction did not select)
                    continue
                dQ[b][live] += g[b, t] * Dn[p, j[live]]
                np.add.at(dD[p], j[live], g[b, t] * Qn[b][live])
        return (torch.from_numpy(dQ) if need_dq else None, torch.from_numpy(dD) if need_dd else None)

    def maxsim_argmax(Q, D, mask, return_rowmax=False):
        calls["argmax"] += 1
        arg, rowmax = _pairs(Q, D, mask, 0)
        return (arg, rowmax) if return_rowmax else arg

    def maxsim_argmax_grouped(Q, D, mask, docs_per_query, return_rowmax=False):
        calls["argmax_grouped"] += 1
        assert D.size(0) == Q.size(0) * docs_per_query
        arg, rowmax = _pairs(Q, D, mask, docs_per_query)
        return (arg, rowmax) if return_rowmax else arg

    def maxsim_backward(Q, D, arg, grad, need_dq=True, need_dd=True):
        calls["backward"] += 1
        return _grads(Q, D, arg, grad, 0, need_dq, need_dd)

    def maxsim_backward_grouped(Q, D, arg, grad, need_dq=True, need_dd=True):
        calls["backward_grouped"] += 1
        return _grads(Q, D, arg, grad, arg.size(1), need_dq, need_dd)

    for name, fn in (("maxsim_argmax", maxsim_argmax), ("maxsim_argmax_grouped", maxsim_argmax_grouped),
                     ("maxsim_backward", maxsim_backward), ("maxsim_backward_grouped", maxsim_backward_grouped)):
        monkeypatch.setattr(M, name, fn)
    for name, fn in (("maxsim_scores", maxsim_scores), ("maxsim_topk", maxsim_topk), ("topk_select", topk_select),
                     ("FlatCorpus", OracleCorpus)):
        monkeypatch.setattr(S, name, fn)
    return calls
