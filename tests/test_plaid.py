"""PLAID index reader + residual decode.

CPU: the oracle restatement of decompress_residuals.cpp / ResidualCodec tables against fixtures produced
by the reference's own codec (tests/golden/make_golden_plaid.py).
GPU: flmr_plaid_decode (through the C ABI) against the same fixtures, and search over a decoded index.
"""
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR
from oracle import maxsim_oracle as O

NBITS = [1, 2, 4, 8]


def _fixture(nbits):
    z = np.load(os.path.join(GOLDEN_DIR, "plaid_nbits%d.npz" % nbits))
    return {k: z[k] for k in z.files}, os.path.join(GOLDEN_DIR, "plaid_nbits%d" % nbits)


@pytest.mark.parametrize("nbits", NBITS)
def test_oracle_decode_matches_reference(nbits):
    g, _ = _fixture(nbits)
    raw = O.plaid_decode(g["codes"], g["residuals"], g["centroids"], g["bucket_weights"], nbits, normalize=False)
    np.testing.assert_array_equal(raw, g["decoded_cpp_raw"])            # decompress_residuals_cpp, bit-exact
    dec = O.plaid_decode(g["codes"], g["residuals"], g["centroids"], g["bucket_weights"], nbits)
    np.testing.assert_allclose(dec, g["decoded_ref"], rtol=0, atol=2e-7)  # + F.normalize


@pytest.mark.parametrize("nbits", NBITS)
def test_plaid_directory_metadata(nbits):
    from ravqa_b200.plaid import read_plaid_metadata
    _, path = _fixture(nbits)
    m = read_plaid_metadata(path)
    assert m["nbits"] == nbits and m["dim"] == 128 and m["num_chunks"] == 1


@pytest.mark.gpu
@pytest.mark.parametrize("nbits", NBITS)
def test_gpu_decode_matches_reference(nbits):
    from ravqa_b200.plaid import plaid_to_flat
    g, path = _fixture(nbits)
    tokens, doclens = plaid_to_flat(path)
    assert doclens.tolist() == g["doclens"].tolist() and tokens.dtype == torch.bfloat16
    got = tokens.float().cpu().numpy()
    ref = O.bf16_round(g["decoded_ref"])
    # fp32 decode is identical; the norm's summation order may differ by an ulp before bf16 rounding
    exact = (got == ref).mean()
    assert exact > 0.995, exact
    np.testing.assert_allclose(got, ref, rtol=2 ** -7, atol=1e-6)


@pytest.mark.gpu
def test_search_over_decoded_plaid_index():
    import ravqa_b200 as R
    g, path = _fixture(8)
    corpus = R.FlatCorpus.from_plaid(path)
    Q, _, _ = O.synth(1, 4, 3, 32, seed=3)
    ref = O.maxsim_scores(Q, O.bf16_round(g["decoded_ref"]), g["doclens"])
    got = R.maxsim_scores(corpus, torch.from_numpy(Q)).cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=1e-4)
    # against the reference's fp32 embeddings the only difference is the bf16 storage of D
    ref32 = O.maxsim_scores(Q, g["decoded_ref"], g["doclens"])
    np.testing.assert_allclose(got, ref32, rtol=1e-3)
    _, pids = R.maxsim_topk(corpus, torch.from_numpy(Q), 5)
    assert np.array_equal(pids.cpu().numpy(), O.topk(ref, 5)[1])


@pytest.mark.gpu
def test_decode_rejects_bad_input():
    from ravqa_b200 import _cabi
    from ravqa_b200.plaid import decode_chunk
    g, _ = _fixture(2)
    cent = torch.from_numpy(g["centroids"]).cuda()
    w = torch.from_numpy(g["bucket_weights"]).cuda()
    codes = torch.from_numpy(g["codes"]).clone()
    codes[3] = 10_000                                                  # corrupt centroid id
    out = torch.empty((codes.numel(), 128), dtype=torch.bfloat16, device="cuda")
    with pytest.raises(_cabi.FlmrError, match="centroid code"):
        decode_chunk(codes, torch.from_numpy(g["residuals"]), cent, w, 2, out)
    with pytest.raises(ValueError):
        decode_chunk(codes, torch.from_numpy(g["residuals"])[:, :4], cent, w, 2, out)
