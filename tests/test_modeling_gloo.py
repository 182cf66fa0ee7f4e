"""CPU, world_size-2 gloo: cross-rank in-batch negatives (the gather the reference leaves
disabled at colbert.py:68-69).  The CUDA scorer is replaced by a torch fp32 restatement of the all-pairs
MaxSim (colbert.py:89-92) — this test covers the exchange logic: differentiable document all-gather,
ragged padded lengths across ranks, label offsets, gradient of the GLOBAL loss reaching local rows."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def torch_all_pairs(Q, D, D_mask):
    """scores[b, p] = sum_i max_{j: mask[p, j]} <Q[b, i], D[p, j]>  (differentiable torch restatement)."""
    S = torch.einsum("bqd,pkd->bpqk", Q.float(), D.float())
    S = S.masked_fill(~D_mask.reshape(D.size(0), 1, D.size(1)).bool().unsqueeze(0), float("-inf"))
    return S.max(dim=-1).values.sum(dim=-1)


def _make(rank, B, nway, nq, nd):
    g = torch.Generator().manual_seed(100 + rank)
    Q = torch.nn.functional.normalize(torch.randn(B, nq, 128, generator=g), dim=-1)
    D = torch.nn.functional.normalize(torch.randn(B * nway, nd, 128, generator=g), dim=-1)
    lens = torch.randint(2, nd + 1, (B * nway,), generator=g)
    mask = (torch.arange(nd)[None, :] < lens[:, None]).unsqueeze(-1)
    return Q, D, mask


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ravqa_b200.modeling import in_batch_negatives_loss
    B, nway, nq = 3, 2, 5
    nds = [7, 11]                                             # ranks pad their documents differently
    Q, D, mask = _make(rank, B, nway, nq, nds[rank])
    Q.requires_grad_(True)
    D.requires_grad_(True)
    loss, S = in_batch_negatives_loss(Q, D, mask, nway, return_scores=True, cross_rank_negatives=True,
                                      all_pairs_fn=torch_all_pairs)
    loss.backward()

    # single-process restatement of the global batch
    parts = [_make(r, B, nway, nq, nds[r]) for r in range(world)]
    Qs = [p[0].clone().requires_grad_(True) for p in parts]
    Ds = [p[1].clone().requires_grad_(True) for p in parts]
    nd_max = max(nds)
    Dg = torch.cat([torch.nn.functional.pad(d, (0, 0, 0, nd_max - d.size(1))) for d in Ds])
    Mg = torch.cat([torch.nn.functional.pad(p[2], (0, 0, 0, nd_max - p[2].size(1))) for p in parts])
    losses = []
    for r in range(world):
        Sr = torch_all_pairs(Qs[r], Dg, Mg)
        labels = r * B * nway + torch.arange(B) * nway
        losses.append(torch.nn.functional.cross_entropy(Sr, labels))
        if r == rank:
            S_ref = Sr.detach()
    sum(losses).backward()                                   # every rank back-propagates its own local loss
    ok = (S.shape == (B, world * B * nway)
          and torch.allclose(S.detach(), S_ref, atol=1e-5)
          and torch.allclose(loss.detach(), losses[rank].detach(), atol=1e-6)
          and torch.allclose(Q.grad, Qs[rank].grad, atol=1e-5)
          and torch.allclose(D.grad, Ds[rank].grad, atol=1e-5))   # includes the OTHER rank's queries
    # and the local-only variant differs (the other rank's negatives matter)
    only_local = in_batch_negatives_loss(Q.detach(), D.detach(), mask, nway, all_pairs_fn=torch_all_pairs)
    ret[rank] = bool(ok) and abs(float(only_local) - float(loss)) > 1e-4
    dist.barrier()
    dist.destroy_process_group()


def test_cross_rank_negatives_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_cross_rank_negatives_requires_process_group():
    import pytest
    from ravqa_b200.modeling import in_batch_negatives_loss
    Q, D, mask = _make(0, 2, 2, 4, 6)
    with pytest.raises(RuntimeError):
        in_batch_negatives_loss(Q, D, mask, 2, cross_rank_negatives=True, all_pairs_fn=torch_all_pairs)


def test_equal_run_length_matches_unique_consecutive():
    """modeling._equal_run_length (vectorised run detection; replaces torch.unique_consecutive(dim=0), 82 ms per
    call on CUDA for 832-token queries) on CPU tensors: r for a repeat_interleave(r) batch, 0 for anything else."""
    from ravqa_b200.modeling import _equal_run_length
    g = torch.Generator().manual_seed(3)
    base = torch.randn(4, 6, 128, generator=g)
    for reps in ([3, 1, 2, 2], [1, 1, 1, 1], [5], [1, 4], [2, 2, 2], [4, 4], [3, 3, 3, 3], [2, 2, 1]):
        Q = torch.repeat_interleave(base[: len(reps)], torch.tensor(reps), dim=0)
        _, counts = torch.unique_consecutive(Q, dim=0, return_counts=True)
        want = int(counts[0]) if (counts == counts[0]).all() and counts[0] > 1 else 0
        assert _equal_run_length(Q) == want, reps
    Q = torch.stack([base[0], base[0], base[1], base[1], base[0], base[0]])   # equal but not adjacent runs
    assert _equal_run_length(Q) == 2
    assert _equal_run_length(base[:1]) == 0
