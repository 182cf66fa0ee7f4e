"""GPU: the retrieval block of the RAG loop (rag_model_blip.py:388-443) — batched search,
device-side gather of the retrieved passages, differentiable re-score — against the oracle and a torch fp32
restatement of the reference's re-score."""
import numpy as np
import pytest
import torch

from oracle import maxsim_oracle as O

pytestmark = pytest.mark.gpu


def test_gather_padded_returns_the_stored_passages():
    import ravqa_b200 as R
    Q, D, dl = O.synth(300, 37, 1, 32, seed=21, ragged=True)
    corpus = R.FlatCorpus(torch.from_numpy(D).to(torch.bfloat16), dl, pid_base=1000)
    off = np.concatenate([[0], np.cumsum(dl)])
    pids = torch.tensor([[1000, 1299, 1123], [1007, 999, -1]])           # two ids outside the shard
    E, M = corpus.gather_padded(pids)
    nd = E.size(2)
    assert E.shape == (2, 3, nd, 128) and M.shape == (2, 3, nd, 1) and nd == int(dl[[0, 299, 123, 7]].max())
    for (r, c), pid in np.ndenumerate(pids.numpy()):
        p = pid - 1000
        if 0 <= p < 300:
            n = int(dl[p])
            ref = torch.from_numpy(D[off[p]:off[p] + n]).to(torch.bfloat16)
            assert torch.equal(E[r, c, :n].cpu(), ref) and bool(M[r, c, :n].all()) and not bool(M[r, c, n:].any())
            assert float(E[r, c, n:].abs().sum()) == 0.0
        else:
            assert not bool(M[r, c].any()) and float(E[r, c].abs().sum()) == 0.0
    E2, M2 = corpus.gather_padded(pids[:, :1], nd_max=5)                    # explicit cut
    assert E2.shape == (2, 1, 5, 128)
    assert torch.equal(E2[0, 0].cpu(), torch.from_numpy(D[off[0]:off[0] + 5]).to(torch.bfloat16))


@pytest.mark.parametrize("n_docs", [5, 3])
def test_retrieve_and_rescore_matches_reference_flow(n_docs):
    import ravqa_b200 as R
    B, nq = 4, 48
    Q, D, dl = O.synth(400, 40, B, nq, seed=22, ragged=True)
    searcher = R.Searcher(index=R.FlatCorpus(torch.from_numpy(D).to(torch.bfloat16), dl))
    Qt = torch.from_numpy(Q).cuda().requires_grad_(True)
    out = searcher.retrieve_and_rescore(Qt, n_docs, generator=torch.Generator().manual_seed(0))
    exact = O.maxsim_scores(Q, D, dl)
    top5 = O.topk(exact, 5)[1]
    ids = out["retrieved_doc_ids"]
    assert ids.shape == (B, n_docs)
    for b in range(B):                                                      # retrieved = (subset of) the exact top-5
        assert set(ids[b]).issubset(set(top5[b])) and len(set(ids[b])) == n_docs
        np.testing.assert_allclose(out["doc_scores"][b].detach().cpu().numpy(), exact[b, ids[b]], rtol=2e-5)
    if n_docs == 5:
        assert np.array_equal(ids, top5)
    # gradient of sum(softmax-weighted scores) w.r.t. the query == torch restatement on the same passages
    w = torch.linspace(0.5, 1.5, B * n_docs, device="cuda").view(B, n_docs)
    (out["doc_scores"] * w).sum().backward()
    Qr = torch.from_numpy(Q).cuda().requires_grad_(True)
    E, M = out["item_embeddings"].float(), out["item_mask"].squeeze(-1)
    S = torch.einsum("bqd,bpkd->bpqk", Qr, E).masked_fill(~M[:, :, None, :], -9999.0).max(-1).values.sum(-1)
    (S * w).sum().backward()
    np.testing.assert_allclose(Qt.grad.cpu().numpy(), Qr.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
