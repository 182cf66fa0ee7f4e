"""GPU diagnostic: wall-clock breakdown (with synchronisation) of Searcher.retrieve_and_rescore's pieces."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ravqa_b200 as R  # noqa: E402
from ravqa_b200.modeling import all_pairs_maxsim, colbert_score  # noqa: E402


def t(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


def main():
    n_p, B, nq, n_docs = 112_000, 8, 832, 5
    g = torch.Generator(device="cuda").manual_seed(0)
    doclens = torch.randint(40, 221, (n_p,), generator=torch.Generator().manual_seed(0)).numpy().astype(np.int32)
    n_tok = int(doclens.sum())
    D = torch.nn.functional.normalize(torch.randn((n_tok, 128), device="cuda", generator=g), dim=-1).bfloat16()
    corpus = R.FlatCorpus(D, doclens)
    searcher = R.Searcher(index=corpus)
    Q = torch.nn.functional.normalize(torch.randn((B, nq, 128), device="cuda", generator=g), dim=-1).requires_grad_(True)
    ms, (s, p) = t(lambda: searcher._search_tensors(Q.detach(), n_docs))
    print("search            %.3f ms" % ms)
    ms, (E, M) = t(lambda: corpus.gather_padded(p))
    print("gather_padded     %.3f ms  -> %s" % (ms, tuple(E.shape)))
    ms, S = t(lambda: all_pairs_maxsim(Q, E.flatten(0, 1), M.flatten(0, 1)))
    print("all_pairs fwd     %.3f ms" % ms)
    def fb():
        Q.grad = None
        all_pairs_maxsim(Q, E.flatten(0, 1), M.flatten(0, 1)).sum().backward()
    ms, _ = t(fb)
    print("all_pairs fwd+bwd %.3f ms" % ms)
    ms, _ = t(lambda: colbert_score(Q[0:1].repeat_interleave(n_docs, dim=0), E[0], M[0]))
    print("aligned score x1  %.3f ms" % ms)
    ms, _ = t(lambda: torch.unique_consecutive(Q.detach()[0:1].repeat_interleave(n_docs, dim=0), dim=0, return_inverse=True))
    print("unique_consecutive(dim=0) on [5, 832, 128]: %.3f ms" % ms)
    def whole():
        Q.grad = None
        out = searcher.retrieve_and_rescore(Q, n_docs)
        out["doc_scores"].sum().backward()
    ms, _ = t(whole)
    print("retrieve_and_rescore fwd+bwd %.3f ms" % ms)


if __name__ == "__main__":
    main()
