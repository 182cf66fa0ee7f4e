"""Generate the PLAID *search* golden with the REFERENCE's own IndexScorer (build container only).

    python tests/golden/make_golden_plaid_search.py

Builds a small clustered corpus (40 topic directions + per-dimension noise 0.06, i.e. token/topic cosine ~0.83), indexes it with — unmodified, from /root/reference/third_party/ColBERT —
    ResidualCodec.compress / save                 colbert/indexing/codecs/residual.py
    optimize_ivf                                  colbert/indexing/utils.py:8-56
(k-means by seeded Lloyd iterations: faiss is absent), loads it back with
    IndexScorer(index_path, use_gpu=False)        colbert/search/index_storage.py:17-64
and records, for several queries x search configurations, what the reference returns from
    IndexScorer.retrieve                          index_storage.py:66-80   (candidate pids, centroid scores)
    IndexScorer.filter_pids (filter_pids.cpp)     index_storage.py:153-156 (pids surviving centroid pruning)
    IndexScorer.rank                              index_storage.py:86-100  (final pids + scores)
Output: tests/golden/plaid_search.npz — the index tensors exactly as IndexScorer holds them, the original
embeddings (bf16 bits; pins oracle.plaid_search.PlaidIndex.build), queries, and the reference outputs.
"""
from __future__ import annotations

import json
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import bf16_bits, import_reference  # noqa: E402

CONFIGS = [  # (ncells, centroid_score_threshold, ndocs, query_maxlen)
    (2, 0.45, 16, 8),
    (1, 0.80, 8, 8),
    (4, 0.40, 32, 12),
]


def main():
    ColBERTConfig = import_reference()[0]
    from colbert.indexing.codecs.residual import ResidualCodec
    from colbert.indexing.utils import optimize_ivf
    from colbert.search.index_storage import IndexScorer

    g = torch.Generator().manual_seed(7)
    n, K, nbits, nq = 250, 128, 2, 12
    doclens = torch.randint(4, 29, (n,), generator=g)
    n_emb = int(doclens.sum())
    topics = torch.nn.functional.normalize(torch.randn(40, 128, generator=g), dim=-1)
    ptop = torch.randint(0, 40, (n, 3), generator=g)
    tok_pid = torch.repeat_interleave(torch.arange(n), doclens)
    tok_topic = ptop[tok_pid, torch.randint(0, 3, (n_emb,), generator=g)]
    D = torch.nn.functional.normalize(topics[tok_topic] + 0.06 * torch.randn(n_emb, 128, generator=g), dim=-1)
    D = D.bfloat16().float()
    off = torch.cat([torch.zeros(1, dtype=torch.long), doclens.cumsum(0)])
    n_queries = 6
    targets = torch.randint(0, n, (n_queries,), generator=g)
    Q = []
    for t in targets.tolist():
        rows = D[off[t]:off[t + 1]]
        rows = rows[torch.randint(0, rows.size(0), (nq,), generator=g)]
        Q.append(torch.nn.functional.normalize(rows + 0.04 * torch.randn(nq, 128, generator=g), dim=-1))
    Q = torch.stack(Q).bfloat16().float()

    # seeded Lloyd k-means (not part of the golden: the centroids are an INPUT of the restated build)
    c = D[torch.randperm(n_emb, generator=g)[:K]].clone()
    for _ in range(6):
        a = (D @ c.T).argmax(dim=1)
        s = torch.zeros_like(c).index_add_(0, a, D)
        cnt = torch.bincount(a, minlength=K).unsqueeze(1)
        c = torch.where(cnt > 0, torch.nn.functional.normalize(s / cnt.clamp_min(1), dim=-1), c)
    centroids = c.half().float()
    heldout_idx = torch.randperm(n_emb, generator=g)[: n_emb // 3]
    heldout = D[heldout_idx]

    index_path = tempfile.mkdtemp(prefix="plaid_search_golden_")
    cfg = ColBERTConfig(nbits=nbits, dim=128, total_visible_gpus=0, index_path=index_path)
    c0 = ResidualCodec(config=cfg, centroids=centroids, avg_residual=None)
    res = heldout - c0.lookup_centroids(c0.compress_into_codes(heldout, out_device="cpu"), out_device="cpu")
    num_options = 2 ** nbits
    quant = torch.arange(0, num_options) * (1 / num_options)
    codec = ResidualCodec(config=cfg, centroids=centroids, avg_residual=torch.abs(res).mean(dim=0).mean(),
                          bucket_cutoffs=res.float().quantile(quant[1:]),
                          bucket_weights=res.float().quantile(quant + 0.5 / num_options))
    comp = codec.compress(D)
    codec.save(index_path)
    comp.save(os.path.join(index_path, "0"))
    json.dump(doclens.tolist(), open(os.path.join(index_path, "doclens.0.json"), "w"))
    json.dump({"passage_offset": 0, "num_passages": n, "num_embeddings": n_emb, "embedding_offset": 0},
              open(os.path.join(index_path, "0.metadata.json"), "w"))
    exported = cfg.export()
    exported.pop("collection", None)
    json.dump({"config": exported, "num_chunks": 1, "num_partitions": K, "num_embeddings": n_emb,
               "avg_doclen": n_emb / n}, open(os.path.join(index_path, "metadata.json"), "w"))
    cs = comp.codes.long().sort()
    optimize_ivf(cs.indices, torch.bincount(cs.values, minlength=K), index_path)

    scorer = IndexScorer(index_path, use_gpu=False)
    out = dict(
        embs_bf16=bf16_bits(D), heldout_idx=heldout_idx.numpy(), queries=Q.numpy(),
        targets=targets.numpy(), configs=np.asarray(CONFIGS, dtype=np.float64),
        centroids=scorer.codec.centroids.numpy(), bucket_cutoffs=scorer.codec.bucket_cutoffs.numpy(),
        bucket_weights=scorer.codec.bucket_weights.numpy(),
        # the loader over-allocates 512 uninitialised rows (residual_embeddings.py:44-52); not part of the index
        codes=scorer.embeddings.codes[:n_emb].numpy(), residuals=scorer.embeddings.residuals[:n_emb].numpy(),
        doclens=scorer.doclens.numpy(),
        ivf=scorer.ivf.tensor[: int(scorer.ivf.lengths.sum())].numpy(), ivf_lengths=scorer.ivf.lengths.numpy(),
        nbits=np.int64(nbits),
    )
    for ci, (ncells, thr, ndocs, qmax) in enumerate(CONFIGS):
        scfg = ColBERTConfig(ncells=ncells, centroid_score_threshold=thr, ndocs=ndocs, total_visible_gpus=0,
                             query_maxlen=qmax)
        for qi in range(n_queries):
            q = Q[qi:qi + 1]
            with torch.inference_mode():
                cand, cscores = scorer.retrieve(scfg, q)
                assert cand.numel() >= ndocs, (ci, qi, cand.numel())     # filter_pids.cpp is undefined below that
                idx = cscores.max(-1).values >= thr
                kept = IndexScorer.filter_pids(cand, cscores, scorer.embeddings.codes, scorer.doclens,
                                               scorer.embeddings_strided.codes_strided.offsets, idx, ndocs)
            pids, scores = scorer.rank(scfg, q)
            key = "c%d_q%d_" % (ci, qi)
            out[key + "cand"] = cand.numpy()
            out[key + "cscores"] = cscores.numpy()
            out[key + "kept"] = kept.numpy()
            out[key + "pids"] = np.asarray(pids, dtype=np.int64)
            out[key + "scores"] = np.asarray(scores, dtype=np.float32)
            print("config %d query %d: %4d candidates -> %3d kept; top-3 %s  target %d"
                  % (ci, qi, cand.numel(), kept.numel(), pids[:3], int(targets[qi])))
    np.savez_compressed(os.path.join(HERE, "plaid_search.npz"), **out)
    shutil.rmtree(index_path, ignore_errors=True)
    print("wrote", os.path.join(HERE, "plaid_search.npz"),
          "%.0f KB" % (os.path.getsize(os.path.join(HERE, "plaid_search.npz")) / 1024))


if __name__ == "__main__":
    main()
