"""GPU: the reference's call sites, line by line, against the real kernels.

The reference package is absent on the GPU box, so ``Run`` / ``RunConfig`` / ``ColBERTConfig`` / ``Queries`` /
``Searcher`` are imported from ``ravqa_b200`` here; tests/test_reference_callsites.py runs the SAME lines in the
build container with the reference's own classes (and proves the stand-ins resolve paths identically).  The
index under ``tests/golden/callsites/`` was written by the reference's unmodified ``CollectionIndexer``
(tests/golden/make_golden_callsites.py); expected rankings are the reference's exact ``colbert_score`` over the
embeddings its own codec decompresses.
"""
import os
import random

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR

pytestmark = pytest.mark.gpu
CKPT_DIR = os.path.join(GOLDEN_DIR, "callsites", "ckpt")


@pytest.fixture()
def golden():
    z = np.load(os.path.join(GOLDEN_DIR, "callsites.npz"))
    return {k: z[k] for k in z.files}


def _exact_topk(golden, key, k):
    order = np.argsort(-golden[key], axis=1, kind="stable")[:, :k]
    return order, np.take_along_axis(golden[key], order, axis=1)


def test_executor_search_lines_verbatim(golden):
    """src/executors/FLMR_executor.py:774-792 on the reference-built PLAID directory."""
    from ravqa_b200 import ColBERTConfig, Queries, Run, RunConfig, Searcher

    class _Self:
        global_rank = 0
        device = torch.device("cuda")
        config = type("C", (), {"ckpt_dir": CKPT_DIR})()
        model_config = {"nbits": 8}
    self = _Self()
    dataloader_idx = 0
    question_ids = ["q%d" % i for i in range(golden["queries"].shape[0])]
    questions = ["question %d" % i for i in range(len(question_ids))]
    query_embeddings = torch.from_numpy(golden["queries"])             # CPU tensor, as the executor holds it
    Ks = [1, 5, int(golden["k"])]
    get_world_size = lambda: 1                                          # noqa: E731

    with Run().context(RunConfig(nranks=1, rank=self.global_rank, root=self.config.ckpt_dir, experiment=f"temp_index_{dataloader_idx}")):
        if self.device == torch.device('cpu'):
            total_visible_gpus = 0
        else:
            if get_world_size() > 1:
                total_visible_gpus = 0
            else:
                total_visible_gpus = 1 #torch.cuda.device_count()

        config = ColBERTConfig(
            total_visible_gpus=total_visible_gpus,
        )
        nbits = self.model_config.get("nbits", 2)
        searcher = Searcher(index=f"temp_index.nbits={nbits}", config=config)
        custom_quries = {question_id: question for question_id, question in zip(question_ids, questions)}
        queries = Queries(data=custom_quries)
        ranking = searcher._search_all_Q(queries, query_embeddings, k=max(Ks))

        ranking_dict = ranking.todict()
        ranking_nz = searcher._search_all_Q(queries, query_embeddings, k=max(Ks), remove_zero_tensors=True).todict()
        single = [searcher.dense_search(query_embeddings[i:i + 1], k=max(Ks), remove_zero_tensors=True)
                  for i in range(len(question_ids))]

        del searcher

    assert searcher_kind_is_plaid(ranking)
    ids, sc = _exact_topk(golden, "exact_scores_bf16", max(Ks))
    ids_nz, sc_nz = _exact_topk(golden, "exact_scores_zero_rows_dropped", max(Ks))
    ids32, sc32 = _exact_topk(golden, "exact_scores", max(Ks))
    assert list(ranking_dict.keys()) == question_ids
    for qi, qid in enumerate(question_ids):
        got = ranking_dict[qid]
        assert [r for _, r, _ in got] == list(range(1, max(Ks) + 1))
        # identical top-k ids vs the reference's exact MaxSim on identical (bf16) inputs ...
        assert [p for p, _, _ in got] == ids[qi].tolist()
        np.testing.assert_allclose([s for _, _, s in got], sc[qi], rtol=5e-4)
        # ... and within the bf16 storage error of its fp32 decode
        np.testing.assert_allclose([s for _, _, s in got], golden["exact_scores"][qi, ids[qi]], rtol=3e-3)
        # remove_zero_tensors: the reference drops the rows (searcher.py:120-126); same ranking either way
        assert [p for p, _, _ in ranking_nz[qid]] == ids_nz[qi].tolist() == ids[qi].tolist()
        np.testing.assert_allclose([s for _, _, s in ranking_nz[qid]], sc_nz[qi], rtol=5e-4)
        pids, ranks, scores = single[qi]
        assert pids == ids[qi].tolist() and ranks == list(range(1, max(Ks) + 1))
        np.testing.assert_allclose(scores, sc[qi], rtol=5e-4)
    # for the record: the reference's own (PLAID-pruned, approximate) search on this index found the same top-1
    assert (golden["plaid_pids"][:, 0] == ids[:, 0]).all()


def searcher_kind_is_plaid(ranking):
    prov = ranking.provenance()
    return prov["index_kind"] == "plaid" and prov["index"].endswith("temp_index_0/indexes/temp_index.nbits=8")


def test_gpu_plaid_decode_matches_reference_codec(golden):
    """The corpus the Searcher scans = the reference's ``ResidualCodec.decompress`` output, rounded to bf16."""
    import ravqa_b200 as R
    from helpers import bf16_bits_to_f32
    path = os.path.join(CKPT_DIR, "temp_index_0", "indexes", "temp_index.nbits=8")
    corpus = R.FlatCorpus.from_plaid(path)
    assert corpus.n_passages == len(golden["doclens"]) and np.array_equal(corpus.doclens, golden["doclens"])
    E, M = corpus.gather_padded(torch.arange(8))
    got = torch.cat([E[p][M[p].squeeze(-1)] for p in range(8)]).float().cpu().numpy()
    want = bf16_bits_to_f32(golden["decoded_bf16_sample"])[: got.shape[0]]
    # bf16 neighbours at most (the reference normalises in fp32 with a different summation order)
    assert np.abs(got - want).max() <= 2 ** -8 * np.abs(want).max()
    assert (got == want).mean() > 0.99
    # shards of the same directory
    a = R.FlatCorpus.from_plaid(path, rank=0, world_size=2)
    b = R.FlatCorpus.from_plaid(path, rank=1, world_size=2)
    assert a.n_passages + b.n_passages == corpus.n_passages and b.pid_base == a.n_passages
    Q = torch.from_numpy(golden["queries"])
    sa, pa = R.maxsim_topk(a, Q, 5)
    sb, pb = R.maxsim_topk(b, Q, 5)
    ms, mp = R.topk_merge(torch.stack([sa, sb]), torch.stack([pa, pb]), 5)
    s, p = R.maxsim_topk(corpus, Q, 5)
    assert torch.equal(mp, p) and torch.equal(ms, s)


def test_rag_retrieval_lines_verbatim(golden):
    """src/models/rag/rag_model_blip.py:288-301 and :388-441 — index through the Run context derived from
    ``index_path``, batched search, per-question re-score through ``question_encoder.score`` — vs a torch fp32
    restatement of the reference's ``colbert_score`` on the same retrieved passages (values and gradient)."""
    from ravqa_b200 import ColBERTConfig, FLMRModelForRetrieval, Queries, Run, RunConfig, Searcher
    doclens = golden["doclens"]
    rng = np.random.default_rng(5)
    nd_max = int(doclens.max())
    item_embeddings = {}
    for pid, n in enumerate(doclens):
        e = np.zeros((nd_max, 128), dtype=np.float32)
        e[:n] = rng.standard_normal((n, 128)).astype(np.float32)
        e[:n] /= np.linalg.norm(e[:n], axis=1, keepdims=True)
        m = np.zeros((nd_max, 1), dtype=np.float32)
        m[:n] = 1
        item_embeddings[pid] = (torch.from_numpy(e).bfloat16().float().numpy(), m)

    class _Self:
        global_rank = 0
        device = torch.device("cuda")
        question_encoder = FLMRModelForRetrieval()
    self = _Self()
    self.question_encoder.colbert_config = ColBERTConfig()
    self.item_embeddings = item_embeddings
    index_path = os.path.join(CKPT_DIR, "temp_index_0")

    index_root = os.path.dirname(index_path)
    index_name = os.path.basename(index_path)
    if self.device == torch.device('cpu'):
        total_visible_gpus = 0
    else:
        total_visible_gpus = 1
    with Run().context(RunConfig(nranks=1, rank=self.global_rank, root=index_root, experiment=index_name)):
        config = ColBERTConfig(
            total_visible_gpus=total_visible_gpus,
        )
        self.index = Searcher(index=f"temp_index.nbits=8", config=config)

    question_hidden_states = torch.from_numpy(golden["queries"]).cuda().requires_grad_(True)
    input_text_sequences = ["question %d" % i for i in range(question_hidden_states.size(0))]
    n_docs = 3
    random.seed(11)
    custom_quries = {i: query for i, query in enumerate(input_text_sequences)}
    queries = Queries(data=custom_quries)
    if n_docs < 5:
        n_docs_retrieve = 5
    else:
        n_docs_retrieve = n_docs
    ranking = self.index._search_all_Q(queries, question_hidden_states.cpu().detach(), k=n_docs_retrieve, progress=False)
    retrieval_results = ranking.todict()
    doc_scores = []
    all_retrieved_doc_indices = []
    for query_index, retrieved_docs in retrieval_results.items():
        retrieved_doc_indices = []
        retrieved_doc_scores = []
        if n_docs != n_docs_retrieve:
            retrieved_docs = random.sample(retrieved_docs, n_docs)
        for doc_index, _, doc_score in retrieved_docs:
            retrieved_doc_indices.append(doc_index)
            retrieved_doc_scores.append(doc_score)
        retrieved_item_embeddings = []
        retrieved_item_embeding_mask = []
        for i in retrieved_doc_indices:
            emb_tuple = self.item_embeddings[i]
            retrieved_item_embeddings.append(torch.Tensor(emb_tuple[0]))
            retrieved_item_embeding_mask.append(torch.Tensor(emb_tuple[1]))
        retrieved_item_embeddings = torch.stack(retrieved_item_embeddings).to(self.device)
        retrieved_item_embeding_mask = torch.stack(retrieved_item_embeding_mask).to(self.device)
        retrieved_query_embedding = question_hidden_states[[query_index]]
        self.question_encoder.colbert_config.nway = len(retrieved_doc_indices)
        Q_duplicated = retrieved_query_embedding.repeat_interleave(self.question_encoder.colbert_config.nway, dim=0).contiguous()
        scores = self.question_encoder.score(Q_duplicated, retrieved_item_embeddings, retrieved_item_embeding_mask)
        doc_scores.append(scores)
        all_retrieved_doc_indices.append(retrieved_doc_indices)
    doc_scores = torch.stack(doc_scores)
    ids = np.array(all_retrieved_doc_indices)
    doc_scores.sum().backward()

    top5 = _exact_topk(golden, "exact_scores_bf16", 5)[0]
    for b in range(ids.shape[0]):
        assert set(ids[b]).issubset(set(top5[b])) and len(set(ids[b])) == n_docs
    Qr = torch.from_numpy(golden["queries"]).cuda().requires_grad_(True)
    want = []
    for b in range(ids.shape[0]):
        E = torch.stack([torch.Tensor(item_embeddings[i][0]) for i in ids[b]]).cuda()
        M = torch.stack([torch.Tensor(item_embeddings[i][1]) for i in ids[b]]).cuda().bool().squeeze(-1)
        S = (E @ Qr[b].T).masked_fill(~M[:, :, None], -9999.0).max(1).values.sum(-1)     # colbert.py:284 + 235-263
        want.append(S)
    want = torch.stack(want)
    want.sum().backward()
    np.testing.assert_allclose(doc_scores.detach().cpu().numpy(), want.detach().cpu().numpy(), rtol=2e-5)
    np.testing.assert_allclose(question_hidden_states.grad.cpu().numpy(), Qr.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)


def test_exhaustive_eval_score_shape_is_block_diagonal():
    """FLMR_executor.py:826-833: ``score(query_embeddings.repeat_interleave(4), items.repeat(n_q), ...)`` — U
    distinct queries x 4 items each.  One grouped launch scores the U*4 aligned pairs only (the kernel launch
    count shows it), results equal the all-pairs matrix's block diagonal."""
    import ravqa_b200 as R
    from ravqa_b200 import _cabi
    g = torch.Generator().manual_seed(9)
    U, r, nq, nd = 37, 4, 320, 60
    Q = torch.nn.functional.normalize(torch.randn(U, nq, 128, generator=g), dim=-1).cuda()
    items = torch.nn.functional.normalize(torch.randn(r, nd, 128, generator=g), dim=-1).cuda()
    imask = (torch.rand(r, nd, 1, generator=g) > 0.2).cuda()
    imask[:, 0] = True
    Q_duplicated = Q.repeat_interleave(r, dim=0).contiguous()
    _cabi.lib().flmr_launch_count(1)
    scores = R.colbert_score(Q_duplicated, items.repeat(U, 1, 1), imask.repeat(U, 1, 1)).reshape(U, -1)
    assert _cabi.lib().flmr_launch_count(1) == 1
    full = R.all_pairs_maxsim(Q, items, imask)                         # [U, r]: every query x the 4 items
    np.testing.assert_allclose(scores.cpu().numpy(), full.cpu().numpy(), rtol=1e-6)
    # irregular runs (lengths 2, 1, 3): still only the aligned pairs, one document per query
    Qi = torch.cat([Q[:1].expand(2, -1, -1), Q[1:2], Q[2:3].expand(3, -1, -1)]).contiguous()
    Di = torch.nn.functional.normalize(torch.randn(6, nd, 128, generator=g), dim=-1).cuda()
    Mi = torch.ones(6, nd, 1, dtype=torch.bool, device="cuda")
    got = R.colbert_score(Qi, Di, Mi)
    want = torch.stack([R.all_pairs_maxsim(Qi[i:i + 1], Di[i:i + 1], Mi[i:i + 1])[0, 0] for i in range(6)])
    np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-6)


def test_grouped_maxsim_values_and_gradients():
    import ravqa_b200 as R
    g = torch.Generator().manual_seed(4)
    B, r, nq, nd = 5, 3, 70, 45
    Q0 = torch.nn.functional.normalize(torch.randn(B, nq, 128, generator=g), dim=-1).bfloat16().float().cuda()
    D0 = torch.nn.functional.normalize(torch.randn(B * r, nd, 128, generator=g), dim=-1).bfloat16().float().cuda()
    M = (torch.rand(B * r, nd, 1, generator=g) > 0.3).cuda()
    M[:, 3] = True
    w = torch.linspace(-1, 2, B * r, device="cuda").view(B, r)
    Q, D = Q0.clone().requires_grad_(True), D0.clone().requires_grad_(True)
    S = R.grouped_maxsim(Q, D, M, r)
    (S * w).sum().backward()
    Qr, Dr = Q0.clone().requires_grad_(True), D0.clone().requires_grad_(True)
    Sr = (torch.einsum("bqd,brkd->brqk", Qr, Dr.view(B, r, nd, 128))
          .masked_fill(~M.view(B, r, 1, nd), -9999.0).max(-1).values.sum(-1))
    (Sr * w).sum().backward()
    np.testing.assert_allclose(S.detach().cpu().numpy(), Sr.detach().cpu().numpy(), rtol=2e-6)
    np.testing.assert_allclose(Q.grad.cpu().numpy(), Qr.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(D.grad.cpu().numpy(), Dr.grad.cpu().numpy(), rtol=1e-4, atol=1e-5)


def test_searcher_corner_cases(tmp_path):
    """filter_fn that keeps nothing -> empty result lists (the reference would return fewer than k too);
    a Ranking survives save/cast; flat indexes are found through the same Run addressing."""
    import ravqa_b200 as R
    from oracle import maxsim_oracle as O
    Q, D, dl = O.synth(50, 12, 2, 32, seed=1, ragged=True)
    root = str(tmp_path)
    with R.Run().context(R.RunConfig(root=root, experiment="exp")):
        path = R.resolve_index_path("flat.idx", R.ColBERTConfig())
        R.save_flat_index(path, torch.from_numpy(D), dl)
        s = R.Searcher(index="flat.idx", config=R.ColBERTConfig(total_visible_gpus=1))
    assert s.index_kind == "flat" and s.index == os.path.join(root, "exp", "indexes/", "flat.idx")
    rk = s._search_all_Q({7: "x", 9: "y"}, torch.from_numpy(Q), k=3, filter_fn=lambda pids: pids[:0])
    assert rk.todict() == {7: [], 9: []}
    assert s.dense_search(torch.from_numpy(Q[:1]), k=3, filter_fn=lambda pids: pids[:0]) == ([], [], [])
    rk = s._search_all_Q({7: "x", 9: "y"}, torch.from_numpy(Q), k=3)        # (the TSV loader wants numeric qids)
    out = rk.save(os.path.join(root, "r.tsv"))
    back = R.Ranking.cast(out)
    assert back.todict().keys() == rk.todict().keys()
    assert [(p, r) for p, r, _ in back.todict()[7]] == [(p, r) for p, r, _ in rk.todict()[7]]
    exact = O.topk(O.maxsim_scores(Q, D, dl), 3)[1]
    assert [p for p, _, _ in rk.todict()[9]] == exact[1].tolist()


@pytest.mark.parametrize("aligned", [True, False])
def test_streamed_index_load_equals_in_memory_corpus(tmp_path, aligned):
    """FlatCorpus.from_index (C-level builder: pread -> pinned ring -> H2D -> padded layout; scatter kernel when
    doclens are not multiples of 4) gives the same corpus as handing the tensors over — whole index and shards,
    single-file and chunked layouts."""
    import ravqa_b200 as R
    from oracle import maxsim_oracle as O
    from ravqa_b200.index_io import finalize_chunked_index, save_flat_chunk
    Q, D, dl = O.synth(700, 40, 3, 64, seed=31, ragged=not aligned)
    if aligned:
        assert (dl % 4 == 0).all()
    Dt = torch.from_numpy(D).to(torch.bfloat16)
    ref = R.FlatCorpus(Dt, dl)
    want = R.maxsim_scores(ref, torch.from_numpy(Q))
    one = str(tmp_path / "one")
    R.save_flat_index(one, Dt, dl)
    chunked = str(tmp_path / "chunked")
    off = np.concatenate([[0], np.cumsum(dl)])
    for c, (a, b) in enumerate([(0, 250), (250, 251), (251, 700)]):
        save_flat_chunk(chunked, c, a, Dt[off[a]:off[b]], dl[a:b])
    finalize_chunked_index(chunked, 3)
    for path in (one, chunked):
        got = R.FlatCorpus.from_index(path)
        assert got.load_stats["gigabytes"] > 0 and got.info.adopted == 0
        assert torch.equal(R.maxsim_scores(got, torch.from_numpy(Q)), want)
        parts = [R.FlatCorpus.from_index(path, rank=r, world_size=3) for r in range(3)]
        assert sum(p.n_passages for p in parts) == 700 and parts[1].pid_base == parts[0].n_passages
        cat = torch.cat([R.maxsim_scores(p, torch.from_numpy(Q)) for p in parts], dim=1)
        assert torch.equal(cat, want)
    # more ranks than passages: the surplus ranks get no corpus
    tiny = str(tmp_path / "tiny")
    R.save_flat_index(tiny, Dt[: off[2]], dl[:2])
    shards = [R.FlatCorpus.from_index(tiny, rank=r, world_size=4) for r in range(4)]
    assert sum(s.n_passages for s in shards if s is not None) == 2 and any(s is None for s in shards)
