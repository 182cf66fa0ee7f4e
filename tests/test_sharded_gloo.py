"""CPU, world_size-2 gloo: the sharded search plumbing (token-balanced shard ranges, global pid
offsets, ONE all-gather of per-shard top-k, merge) with oracle-backed local scorers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import maxsim_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ravqa_b200.sharded import ShardedSearcher, shard_ranges
    Q, D, dl = O.synth(300, 40, 3, 32, seed=7, ragged=True)
    off = np.concatenate([[0], np.cumsum(dl)])
    p0, p1 = shard_ranges(dl, world)[rank]
    Dl, dll = D[off[p0]:off[p1]], dl[p0:p1]

    def local_topk(Qt, kk):
        s = O.maxsim_scores(Qt.numpy(), Dl, dll)
        ts, tp = O.topk(s, kk, pid_base=p0)
        return torch.from_numpy(ts), torch.from_numpy(tp)

    def merge(gs, gp, kk):
        R, B, kin = gs.shape
        out_s = torch.empty(B, kk)
        out_p = torch.empty(B, kk, dtype=torch.int64)
        for b in range(B):
            s = gs[:, b].reshape(-1).numpy()
            p = gp[:, b].reshape(-1).numpy()
            keep = p >= 0
            s, p = s[keep], p[keep]
            order = np.lexsort((p, -s.astype(np.float64)))[:kk]
            out_s[b, :len(order)] = torch.from_numpy(s[order])
            out_p[b, :len(order)] = torch.from_numpy(p[order])
        return out_s, out_p

    s, p = ShardedSearcher(local_topk, merge).search(torch.from_numpy(Q), k)
    ref_s, ref_p = O.topk(O.maxsim_scores(Q, D, dl), k)
    ok = np.array_equal(p.numpy(), ref_p) and np.allclose(s.numpy(), ref_s, rtol=1e-6)
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("k", [5, 40])
def test_sharded_search_two_ranks_gloo(k):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), k, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_ranges_balanced_and_contiguous():
    from ravqa_b200.sharded import shard_ranges
    rng = np.random.default_rng(0)
    dl = rng.integers(1, 300, size=10_000)
    for world in (1, 2, 4, 8):
        rs = shard_ranges(dl, world)
        assert rs[0][0] == 0 and rs[-1][1] == len(dl)
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        tok = [dl[a:b].sum() for a, b in rs]
        assert max(tok) - min(tok) <= 2 * dl.max()
