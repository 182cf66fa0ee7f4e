"""CPU: the 'flipr' reduction of this package's colbert_score (host logic around the arg-max entry points, which the
numpy oracle substitutes here) against the reference-generated golden tests/golden/flipr.npz."""
import os
import types

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, bf16_bits_to_f32


@pytest.mark.parametrize("nq", [96, 70])
def test_flipr_reduction_against_reference_golden(nq, monkeypatch):
    import oracle_backend
    import ravqa_b200 as R
    calls = oracle_backend.install(monkeypatch)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: False)
    z = np.load(os.path.join(GOLDEN_DIR, "flipr.npz"))
    k = "nq%d_" % nq
    cfg = types.SimpleNamespace(interaction="flipr", query_maxlen=64)
    r = int(z["docs_per_query"])
    w = torch.from_numpy(z["weights"])
    mask = torch.from_numpy(z[k + "mask"]).unsqueeze(-1)
    for shape in ("aligned", "one"):
        Q = torch.from_numpy(bf16_bits_to_f32(z[k + "Q_bf16"])).requires_grad_(True)
        D = torch.from_numpy(bf16_bits_to_f32(z[k + "D_bf16"])).requires_grad_(True)
        Qin = Q.repeat_interleave(r, dim=0).contiguous() if shape == "aligned" else Q[:1]
        s = R.colbert_score(Qin, D, mask, config=cfg)
        (s * w).sum().backward()
        np.testing.assert_allclose(s.detach().numpy(), z[k + shape], rtol=1e-5)
        np.testing.assert_allclose(Q.grad.numpy(), z[k + shape + "_dQ"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(D.grad.numpy(), z[k + shape + "_dD"], rtol=1e-4, atol=1e-6)
    assert calls["argmax_grouped"] == 1 and calls["argmax"] == 1     # one launch per call, selection on its output
    with pytest.raises(AssertionError):
        R.colbert_score(Q[:1], D, mask, config=types.SimpleNamespace(interaction="flipr", query_maxlen=32))
    with pytest.raises(AssertionError):
        R.colbert_score(Q[:1], D, mask, config=types.SimpleNamespace(interaction="other", query_maxlen=64))
