"""Generate training-path goldens with the REFERENCE's own functions and autograd (build container only).

    python tests/golden/make_golden_train.py

Calls, unmodified, from /root/reference/third_party/ColBERT/colbert/modeling/colbert.py:
    ColBERT.compute_ib_loss_new (:82-113)   on a stub `self` carrying only colbert_config and loss_fn
                                            (the method touches nothing else; no BERT weights needed)
    ColBERT.score (:217-224) -> colbert_score (:268-286), aligned form with repeat_interleave'd queries
and back-propagates through them with torch autograd, fp32 on CPU.  Inputs are bf16-representable, so the
CUDA path (which rounds its operands to bf16) sees exactly the same numbers.
Output: tests/golden/train_ib_loss.npz — inputs, the loss, the aligned scores and the gradients w.r.t. the
query and document embeddings.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import bf16_bits, import_reference  # noqa: E402


def main():
    ColBERTConfig, ColBERT = import_reference()[:2]
    cfg = ColBERTConfig(total_visible_gpus=0, nway=2)
    stub = types.SimpleNamespace(colbert_config=cfg, loss_fn=torch.nn.CrossEntropyLoss(), use_gpu=False)
    g = torch.Generator().manual_seed(31)
    B, nway, nq, nd = 4, 2, 40, 57
    Q = torch.nn.functional.normalize(torch.randn(B, nq, 128, generator=g), dim=-1).bfloat16().float()
    D = torch.nn.functional.normalize(torch.randn(B * nway, nd, 128, generator=g), dim=-1).bfloat16().float()
    # ColBERT.doc's mask: padding at the end AND punctuation holes inside the passage (colbert.py:199-203)
    lens = torch.randint(20, nd + 1, (B * nway,), generator=g)
    mask = torch.arange(nd)[None, :] < lens[:, None]
    mask &= torch.rand(B * nway, nd, generator=g) > 0.1
    mask[:, 0] = True
    D = D * mask[..., None]                                        # doc() zeroes masked tokens (:196)
    D_mask = mask.unsqueeze(-1)

    # in-batch negatives loss
    Q1, D1 = Q.clone().requires_grad_(True), D.clone().requires_grad_(True)
    loss = ColBERT.compute_ib_loss_new(stub, Q1, D1, D_mask)
    loss.backward()
    # aligned score of the training forward (colbert.py:71-73), arbitrary upstream weights
    Q2, D2 = Q.clone().requires_grad_(True), D.clone().requires_grad_(True)
    scores = ColBERT.score(stub, Q2.repeat_interleave(nway, dim=0).contiguous(), D2, D_mask)
    w = torch.linspace(-1.0, 1.0, scores.numel())
    (scores * w).sum().backward()

    np.savez_compressed(os.path.join(HERE, "train_ib_loss.npz"),
                        Q_bf16=bf16_bits(Q), D_bf16=bf16_bits(D), mask=mask.numpy(), nway=np.int64(nway),
                        ib_loss=np.float32(loss.item()), ib_dQ=Q1.grad.numpy(), ib_dD=D1.grad.numpy(),
                        scores=scores.detach().numpy(), score_weights=w.numpy(), score_dQ=Q2.grad.numpy(),
                        score_dD=D2.grad.numpy())
    print("ib_loss = %.6f  |dQ|max = %.3e  |dD|max = %.3e  scores[:4] = %s"
          % (loss.item(), Q1.grad.abs().max(), D1.grad.abs().max(), scores[:4].tolist()))
    print("wrote", os.path.join(HERE, "train_ib_loss.npz"),
          "%.0f KB" % (os.path.getsize(os.path.join(HERE, "train_ib_loss.npz")) / 1024))


if __name__ == "__main__":
    main()
