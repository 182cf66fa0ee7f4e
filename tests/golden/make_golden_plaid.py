"""Generate PLAID-index fixtures with the REFERENCE's own codec (build container only).

    python tests/golden/make_golden_plaid.py

For nbits in {1, 2, 4, 8} builds a tiny index directory in the reference's on-disk format
 using, unmodified, from /root/reference/third_party/ColBERT:
    ResidualCodec (compress_into_codes / lookup_centroids / binarize / compress / save)
                                                 colbert/indexing/codecs/residual.py
    ResidualEmbeddings.save                       colbert/indexing/codecs/residual_embeddings.py
    bucket cutoffs / weights as in CollectionIndexer._compute_avg_residual
                                                 colbert/indexing/collection_indexer.py:290-314
and records what the reference DEcodes from it:
    decoded_ref      ResidualCodec.decompress (CPU branch, residual.py:242-278): centroid + bucket weight,
                     L2-normalised, fp32
    decoded_cpp_raw  decompress_residuals_cpp (colbert/search/decompress_residuals.cpp, JIT-built) for all
                     pids, before normalisation (the call of IndexScorer.score_pids, index_storage.py:160-172)
Output: tests/golden/plaid_nbits<N>/ (index files) + tests/golden/plaid_nbits<N>.npz (goldens).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402  (same import shims)


def main():
    ColBERTConfig = import_reference()[0]
    from colbert.indexing.codecs.residual import ResidualCodec
    from colbert.search.index_storage import IndexScorer
    IndexScorer.try_load_torch_extensions(False)       # JIT-builds filter_pids.cpp + decompress_residuals.cpp
    decompress_cpp = IndexScorer.decompress_residuals

    for nbits in (1, 2, 4, 8):
        g = torch.Generator().manual_seed(100 + nbits)
        K, n_passages = 48, 23
        doclens = torch.randint(3, 40, (n_passages,), generator=g)
        n_tok = int(doclens.sum())
        embs = torch.nn.functional.normalize(torch.randn(n_tok, 128, generator=g), dim=-1)
        centroids = torch.nn.functional.normalize(torch.randn(K, 128, generator=g), dim=-1).half().float()
        cfg = ColBERTConfig(nbits=nbits, dim=128, total_visible_gpus=0)
        # bucket cutoffs / weights exactly as CollectionIndexer._compute_avg_residual does on its held-out sample
        c0 = ResidualCodec(config=cfg, centroids=centroids, avg_residual=None)
        heldout = embs[torch.randperm(n_tok, generator=g)[: n_tok // 2]]
        recon = c0.lookup_centroids(c0.compress_into_codes(heldout, out_device="cpu"), out_device="cpu")
        res = heldout - recon
        avg_residual = torch.abs(res).mean(dim=0).mean()
        num_options = 2 ** nbits
        quantiles = torch.arange(0, num_options) * (1 / num_options)
        bucket_cutoffs = res.float().quantile(quantiles[1:])
        bucket_weights = res.float().quantile(quantiles + (0.5 / num_options))
        codec = ResidualCodec(config=cfg, centroids=centroids, avg_residual=avg_residual,
                              bucket_cutoffs=bucket_cutoffs, bucket_weights=bucket_weights)
        comp = codec.compress(embs)
        decoded_ref = codec.decompress(comp)                                   # normalised fp32
        offsets = torch.cat([torch.zeros(1, dtype=torch.long), doclens.cumsum(0)])
        pids = torch.arange(n_passages, dtype=torch.int32)
        decoded_cpp_raw = decompress_cpp(pids, doclens.long(), offsets.long(), codec.bucket_weights,
                                         codec.reversed_bit_map, codec.decompression_lookup_table,
                                         comp.residuals, comp.codes, codec.centroids, 128, nbits)
        # ---- index directory in the reference's format ----
        d = os.path.join(HERE, "plaid_nbits%d" % nbits)
        os.makedirs(d, exist_ok=True)
        codec.save(d)                                                          # centroids.pt, avg_residual.pt, buckets.pt
        comp.save(os.path.join(d, "0"))                                        # 0.codes.pt, 0.residuals.pt
        with open(os.path.join(d, "doclens.0.json"), "w") as f:
            json.dump(doclens.tolist(), f)
        with open(os.path.join(d, "0.metadata.json"), "w") as f:
            json.dump({"passage_offset": 0, "num_passages": n_passages, "num_embeddings": n_tok,
                       "embedding_offset": 0}, f)
        with open(os.path.join(d, "metadata.json"), "w") as f:
            json.dump({"config": {"nbits": nbits, "dim": 128}, "num_chunks": 1, "num_partitions": K,
                       "num_embeddings": n_tok, "avg_doclen": n_tok / n_passages}, f)
        np.savez_compressed(os.path.join(HERE, "plaid_nbits%d.npz" % nbits),
                            decoded_ref=decoded_ref.numpy(), decoded_cpp_raw=decoded_cpp_raw.numpy(),
                            doclens=doclens.numpy().astype(np.int32),
                            codes=comp.codes.numpy(), residuals=comp.residuals.numpy(),
                            centroids=codec.centroids.numpy(), bucket_weights=codec.bucket_weights.numpy())
        raw_n = torch.nn.functional.normalize(decoded_cpp_raw.float(), p=2, dim=-1)
        print("nbits=%d: tokens=%d  |decompress() - normalize(cpp)|max = %.2e  recon cos(min) = %.3f"
              % (nbits, n_tok, (decoded_ref - raw_n).abs().max().item(),
                 (decoded_ref * embs).sum(-1).min().item()))


if __name__ == "__main__":
    main()
