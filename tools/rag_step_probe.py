"""GPU probe: the retrieval block of the RAG loop (C5) at the OK-VQA shape — 8 questions with
FLMR queries of Nq = 832 tokens (512 text + 320 vision), 112k passages (ragged Nd 40-220), n_docs = 5:
search -> gather the retrieved passages out of HBM -> differentiable re-score -> backward to the queries.

    python tools/rag_step_probe.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ravqa_b200 as R  # noqa: E402


def ev():
    return torch.cuda.Event(enable_timing=True)


def main():
    n_p, B, nq, n_docs = 112_000, 8, 832, 5
    g = torch.Generator(device="cuda").manual_seed(0)
    doclens = torch.randint(40, 221, (n_p,), generator=torch.Generator().manual_seed(0)).numpy().astype(np.int32)
    n_tok = int(doclens.sum())
    D = torch.empty((n_tok, 128), dtype=torch.bfloat16, device="cuda")
    for a in range(0, n_tok, 1 << 22):
        b = min(n_tok, a + (1 << 22))
        D[a:b] = torch.nn.functional.normalize(torch.randn((b - a, 128), device="cuda", generator=g), dim=-1).bfloat16()
    searcher = R.Searcher(index=R.FlatCorpus(D, doclens))
    Q = torch.nn.functional.normalize(torch.randn((B, nq, 128), device="cuda", generator=g), dim=-1).requires_grad_(True)

    def step():
        Q.grad = None
        out = searcher.retrieve_and_rescore(Q, n_docs)
        torch.log_softmax(out["doc_scores"], dim=-1)[:, 0].sum().backward()
        return out

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    reps = 10
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    total = e0.elapsed_time(e1) / reps
    e2, e3 = ev(), ev()
    e2.record()
    for _ in range(reps):
        searcher._search_tensors(Q.detach(), n_docs)
    e3.record()
    torch.cuda.synchronize()
    search = e2.elapsed_time(e3) / reps
    print("| step (B = %d questions, Nq = %d, %d passages, n_docs = %d) | ms |\n|---|---:|" % (B, nq, n_p, n_docs))
    print("| whole block: search + gather + re-score forward + backward to Q | %.3f |" % total)
    print("| of which exhaustive search (`_search_tensors`) | %.3f |" % search)
    print("| gather + differentiable re-score + backward | %.3f |" % (total - search))


if __name__ == "__main__":
    main()
