"""GPU: size-independent properties at BASELINE.json's FULL size (configs[2]: 1,000,000 passages x 180
tokens x 128 dims = 46 GB of bf16, Nq = 320) where no CPU oracle can run, and the NCCL sharded path."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make_corpus(R, n_p, nd, dev="cuda"):
    D = torch.empty((n_p * nd, 128), dtype=torch.bfloat16, device=dev)
    step = 25_000
    for c0 in range(0, n_p, step):
        c1 = min(n_p, c0 + step)
        g = torch.Generator(device=dev).manual_seed(977 * c0 + 5)
        D[c0 * nd:c1 * nd] = torch.nn.functional.normalize(
            torch.randn(((c1 - c0) * nd, 128), device=dev, generator=g), dim=-1).to(torch.bfloat16)
    return D, R.FlatCorpus(D, np.full(n_p, nd, dtype=np.int32))


def test_full_size_1m_passages_properties():
    import ravqa_b200 as R
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~60 GB of free HBM")
    n_p, nd, nq = 1_000_000, 180, 320
    D, corpus = _make_corpus(R, n_p, nd)
    assert corpus.info.adopted == 1 and corpus.info.n_tokens == n_p * nd
    g = torch.Generator(device="cuda").manual_seed(3)
    Q = torch.nn.functional.normalize(torch.randn((3, nq, 128), device="cuda", generator=g), dim=-1).to(torch.bfloat16)
    # planted positives: 24 token rows of a known passage copied into each query
    targets = [17, 500_000, 999_999]                      # first CTA, middle, very last passage
    for b, t in enumerate(targets):
        Q[b, :24] = D[t * nd: t * nd + 24]
    s_all = R.maxsim_scores(corpus, Q)
    for k in (5, 100):
        ts, tp = R.maxsim_topk(corpus, Q, k)
        rs, rp = torch.sort(s_all, dim=1, descending=True, stable=True)
        assert torch.equal(tp, rp[:, :k]) and torch.equal(ts, rs[:, :k])     # fused top-k == sort of all scores
    _, tp = R.maxsim_topk(corpus, Q, 5)
    assert [int(tp[b, 0]) for b in range(3)] == targets                     # Recall@1 of planted positives
    # torch fp32 restatement of colbert_score on the top hits + 3000 random passages + the corpus ends
    pick = torch.cat([tp.reshape(-1), torch.randint(0, n_p, (3000,), device="cuda"),
                      torch.tensor([0, 1, n_p - 2, n_p - 1], device="cuda")])
    Dp = D.view(n_p, nd, 128)[pick].float()
    for b in range(3):
        ref = (Dp @ Q[b].float().T).max(dim=1).values.sum(dim=-1)
        np.testing.assert_allclose(s_all[b, pick].cpu().numpy(), ref.cpu().numpy(), rtol=2e-5)
    # independent SIMT kernel on the whole corpus for one query
    s_simt = R.debug_scores_simt(corpus, Q[:1])
    rel = ((s_all[:1] - s_simt).abs() / s_simt.abs().clamp_min(1e-6)).max().item()
    assert rel < 2e-5, rel
    assert torch.equal(s_all, R.maxsim_scores(corpus, Q))                    # deterministic
    # four queries = one CTA-pair pass (the headline's path): same scores as the normal passes above up to the last
    # bits a different tile partition can move, same ranking, fused top-k == sort of its own scores
    Q4 = torch.cat([Q, Q[:1].flip(1)])
    s4 = R.maxsim_scores(corpus, Q4)
    np.testing.assert_allclose(s4[:3].cpu().numpy(), s_all.cpu().numpy(), rtol=2e-6)
    ts4, tp4 = R.maxsim_topk(corpus, Q4, 100)
    rs4, rp4 = torch.sort(s4, dim=1, descending=True, stable=True)
    assert torch.equal(tp4, rp4[:, :100]) and torch.equal(ts4, rs4[:, :100])
    assert torch.equal(tp4[:3, :5], tp)
    corpus.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, ret):
    import torch.distributed as dist
    import ravqa_b200 as R
    from oracle import maxsim_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    Q, D, dl = O.synth(3000, 60, 4, 64, seed=5, ragged=True)
    off = np.concatenate([[0], np.cumsum(dl)])
    p0, p1 = R.shard_ranges(dl, world)[rank]
    corpus = R.FlatCorpus(torch.from_numpy(D[off[p0]:off[p1]]).to(torch.bfloat16), dl[p0:p1], device=rank, pid_base=p0)
    s, p = R.ShardedSearcher.from_corpus(corpus).search(torch.from_numpy(Q).cuda(rank), 10)
    rs, rp = O.topk(O.maxsim_scores(Q, D, dl), 10)
    ok = bool(np.array_equal(p.cpu().numpy(), rp) and np.allclose(s.cpu().numpy(), rs, rtol=2e-5))
    # the same through the C-level exchange (flmr_comm_create / flmr_maxsim_topk_sharded: the library's own NCCL
    # communicator, one grouped all-gather + merge enqueued by one C call)
    from ravqa_b200.sharded import NcclExchange
    ex = NcclExchange(torch.device("cuda", rank))
    s2, p2 = ex.search(corpus, torch.from_numpy(Q).cuda(rank), 10)
    ok = ok and torch.equal(p2, p) and torch.equal(s2, s)
    s3, p3 = ex.exchange(*R.maxsim_topk(corpus, torch.from_numpy(Q).cuda(rank), 10), 10)
    ok = ok and torch.equal(p3, p) and torch.equal(s3, s)
    # the reference-facing Searcher in sharded mode, from a flat index and from the reference-built PLAID directory
    import tempfile
    from helpers import GOLDEN_DIR
    root = os.environ["FLMR_TEST_SHARED_TMP"]
    if rank == 0:
        R.save_flat_index(os.path.join(root, "e", "indexes", "flat"), torch.from_numpy(D), dl)
    dist.barrier()
    with R.Run().context(R.RunConfig(root=root, experiment="e")):
        sr = R.Searcher(index="flat", config=R.ColBERTConfig(total_visible_gpus=1), shard_across_ranks=True)
    rk = sr._search_all_Q(list(range(4)), torch.from_numpy(Q), k=10).todict()
    ok = ok and all([pid for pid, _, _ in rk[b]] == rp[b].tolist() for b in range(4))
    # k larger than the whole corpus and a filter: every rank returns the same short, merged lists
    keep_fn = lambda pids: pids[(pids % 7 == 0)]                             # noqa: E731
    rk = sr._search_all_Q(list(range(4)), torch.from_numpy(Q), k=5, filter_fn=keep_fn).todict()
    full = O.maxsim_scores(Q, D, dl)
    for b in range(4):
        cand = np.arange(0, len(dl), 7)
        want = cand[np.lexsort((cand, -full[b, cand].astype(np.float64)))[:5]]
        ok = ok and [pid for pid, _, _ in rk[b]] == want.tolist()
    z = np.load(os.path.join(GOLDEN_DIR, "callsites.npz"))
    with R.Run().context(R.RunConfig(root=os.path.join(GOLDEN_DIR, "callsites", "ckpt"), experiment="temp_index_0")):
        sp = R.Searcher(index="temp_index.nbits=8", config=R.ColBERTConfig(), shard_across_ranks=True)
    rk = sp._search_all_Q(list(range(z["queries"].shape[0])), torch.from_numpy(z["queries"]), k=128).todict()
    want = np.argsort(-z["exact_scores_bf16"], axis=1, kind="stable")[:, :128]
    # k beyond every shard's size (160 passages over `world` shards): short local lists, full merged lists
    ok = ok and all([pid for pid, _, _ in rk[b]] == want[b].tolist() for b in range(want.shape[0]))
    ret[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_nccl():
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import tempfile
    ret = mp.Manager().dict()
    with tempfile.TemporaryDirectory() as tmp:
        os.environ["FLMR_TEST_SHARED_TMP"] = tmp
        mp.spawn(_nccl_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {r: True for r in range(world)}
