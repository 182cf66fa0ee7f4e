"""Generate goldens for ``config.interaction == 'flipr'`` with the REFERENCE's own functions (build container only).

    python tests/golden/make_golden_flipr.py

Calls, unmodified, /root/reference/third_party/ColBERT/colbert/modeling/colbert.py: ``ColBERT.score`` (:217-224) ->
``colbert_score`` (:268-286) -> ``colbert_score_reduce`` with its 'flipr' branch (:248-261; query_maxlen = 64, the
only value the reference allows), fp32 on CPU, and back-propagates through it.  Two query lengths: 96 (32 tokens
beyond query_maxlen: both terms) and 70 (6 beyond: first term only); both caller shapes (one query against all
documents, aligned pairs built with repeat_interleave).  Inputs are bf16-representable.
Output: tests/golden/flipr.npz.
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
    cfg = ColBERTConfig(total_visible_gpus=0, interaction="flipr", query_maxlen=64)
    stub = types.SimpleNamespace(colbert_config=cfg, use_gpu=False)
    out = {}
    for nq in (96, 70):
        g = torch.Generator().manual_seed(500 + nq)
        B, r, nd = 3, 4, 45
        Q = torch.nn.functional.normalize(torch.randn(B, nq, 128, generator=g), dim=-1).bfloat16().float()
        D = torch.nn.functional.normalize(torch.randn(B * r, nd, 128, generator=g), dim=-1).bfloat16().float()
        lens = torch.randint(10, nd + 1, (B * r,), generator=g)
        mask = torch.arange(nd)[None, :] < lens[:, None]
        mask &= torch.rand(B * r, nd, generator=g) > 0.1
        mask[:, 0] = True
        D = D * mask[..., None]
        D_mask = mask.unsqueeze(-1)
        # aligned pairs (colbert.py:71-73)
        Q1, D1 = Q.clone().requires_grad_(True), D.clone().requires_grad_(True)
        s_al = ColBERT.score(stub, Q1.repeat_interleave(r, dim=0).contiguous(), D1, D_mask)
        w = torch.linspace(-1.0, 1.0, s_al.numel())
        (s_al * w).sum().backward()
        # one query against all documents (colbert.py:282)
        Q2, D2 = Q.clone().requires_grad_(True), D.clone().requires_grad_(True)
        s_one = ColBERT.score(stub, Q2[:1], D2, D_mask)
        (s_one * w).sum().backward()
        k = "nq%d_" % nq
        out.update({k + "Q_bf16": bf16_bits(Q), k + "D_bf16": bf16_bits(D), k + "mask": mask.numpy(),
                    k + "aligned": s_al.detach().numpy(), k + "aligned_dQ": Q1.grad.numpy(),
                    k + "aligned_dD": D1.grad.numpy(), k + "one": s_one.detach().numpy(),
                    k + "one_dQ": Q2.grad.numpy(), k + "one_dD": D2.grad.numpy()})
        print("nq=%d aligned[:3]=%s one[:3]=%s" % (nq, s_al[:3].tolist(), s_one[:3].tolist()))
    out["weights"] = w.numpy()
    out["docs_per_query"] = np.int64(4)
    path = os.path.join(HERE, "flipr.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
