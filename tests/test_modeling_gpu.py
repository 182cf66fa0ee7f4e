"""GPU: differentiable scoring surface — colbert_score / model.score / in-batch negatives —
against a plain torch fp32 restatement of the reference (colbert.py:82-113, 235-286) on the same
bf16-rounded inputs: forward scores and gradients."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_all_pairs(Q, D, mask):
    """colbert_score semantics for every (query, document) pair, autograd-capable torch fp32."""
    S = torch.einsum("bqd,pkd->bpqk", Q, D)
    S = S.masked_fill(~mask[None, :, None, :], -9999.0)              # colbert_score_reduce (colbert.py:240)
    return S.max(dim=-1).values.sum(dim=-1)                           # [B, n]


def _inputs(B, nq, n, nd, seed):
    g = torch.Generator().manual_seed(seed)
    Q = torch.nn.functional.normalize(torch.randn(B, nq, 128, generator=g), dim=-1).bfloat16().float().cuda()
    D = torch.nn.functional.normalize(torch.randn(n, nd, 128, generator=g), dim=-1).bfloat16().float().cuda()
    lens = torch.randint(max(1, nd // 3), nd + 1, (n,), generator=g)
    mask = (torch.arange(nd)[None, :] < lens[:, None]).cuda()
    return Q, D * mask[..., None], mask


def test_colbert_score_forms_and_grads():
    import ravqa_b200 as R
    Q, D, mask = _inputs(6, 40, 6, 50, seed=0)
    # Q.size(0) == 1: one query against all documents
    s1 = R.colbert_score(Q[:1], D, mask.unsqueeze(-1))
    np.testing.assert_allclose(s1.cpu().numpy(), _ref_all_pairs(Q[:1], D, mask)[0].cpu().numpy(), rtol=2e-5)
    # Q.size(0) == n: aligned pairs (the repeat_interleave form of the callers)
    Qg, Dg = Q.clone().requires_grad_(True), D.clone().requires_grad_(True)
    s = R.colbert_score(Qg, Dg, mask.unsqueeze(-1))
    Qr, Dr = Q.clone().requires_grad_(True), D.clone().requires_grad_(True)
    sr = _ref_all_pairs(Qr, Dr, mask).diagonal()
    np.testing.assert_allclose(s.detach().cpu().numpy(), sr.detach().cpu().numpy(), rtol=2e-5)
    w = torch.linspace(0.5, 1.5, s.numel(), device="cuda")
    (s * w).sum().backward()
    (sr * w).sum().backward()
    np.testing.assert_allclose(Qg.grad.cpu().numpy(), Qr.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(Dg.grad.cpu().numpy(), Dr.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)


def test_in_batch_negatives_loss_and_model_facade():
    import ravqa_b200 as R
    B, nway = 5, 2
    Q, D, mask = _inputs(B, 64, B * nway, 70, seed=1)
    Qg, Dg = Q.clone().requires_grad_(True), D.clone().requires_grad_(True)
    model = R.FLMRModelForRetrieval(nway=nway)
    aligned, loss = model(Qg, Dg, mask)
    Qr, Dr = Q.clone().requires_grad_(True), D.clone().requires_grad_(True)
    S = _ref_all_pairs(Qr, Dr, mask)
    labels = torch.arange(B, device="cuda") * nway                  # positive of query i at column i*nway
    loss_r = torch.nn.functional.cross_entropy(S, labels)
    np.testing.assert_allclose(loss.item(), loss_r.item(), rtol=1e-5)
    rows = torch.arange(B, device="cuda").repeat_interleave(nway)
    np.testing.assert_allclose(aligned.detach().cpu().numpy(),
                               S[rows, torch.arange(B * nway, device="cuda")].detach().cpu().numpy(), rtol=2e-5)
    loss.backward()
    loss_r.backward()
    np.testing.assert_allclose(Qg.grad.cpu().numpy(), Qr.grad.cpu().numpy(), rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(Dg.grad.cpu().numpy(), Dr.grad.cpu().numpy(), rtol=1e-3, atol=1e-6)
    with pytest.raises(RuntimeError):
        model.query(None)
    with pytest.raises(ValueError, match="no unmasked token"):
        R.colbert_score(Q[:1], D, torch.zeros_like(mask).unsqueeze(-1))
