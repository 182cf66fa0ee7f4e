"""Passage-sharded search across the GPUs of one box.

The reference does not shard: under DDP every rank repeats the whole search on CPU
(src/executors/FLMR_executor.py:778-796).  Here rank r keeps a contiguous range of passages
resident (token-balanced), every rank scans its shard with the fused kernel, and ONE collective —
an all-gather of the per-shard top-k ``[B, k]`` (fp32 score, int64 pid) — precedes a tiny merge
kernel on every rank.  No data-path collective touches the token matrix.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_ranges(doclens: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Contiguous passage ranges balanced by TOKEN count (ragged corpora), one per rank."""
    doclens = np.asarray(doclens, dtype=np.int64)
    n = len(doclens)
    off = np.concatenate([[0], np.cumsum(doclens)])
    bounds = [0]
    for r in range(1, world_size):
        target = off[-1] * r // world_size
        p = int(np.searchsorted(off, target, side="left"))
        bounds.append(min(max(p, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


class ShardedSearcher:
    """``local_topk(Q, k) -> (scores [B,k], pids [B,k])`` runs on this rank's shard (global pids);
    ``merge(scores [R,B,k], pids [R,B,k], k) -> (scores [B,k], pids [B,k])`` merges the gathered
    lists.  The product wiring (``from_corpus``) uses the CUDA scan + merge kernels; the CPU/gloo
    tests inject oracle-backed callables to exercise the sharding and exchange logic without a GPU.
    """

    def __init__(self, local_topk: Callable, merge: Callable, group=None):
        self.local_topk = local_topk
        self.merge = merge
        self.group = group

    @classmethod
    def from_corpus(cls, corpus, group=None, relu: bool = False) -> "ShardedSearcher":
        from .maxsim import maxsim_topk, topk_merge
        return cls(lambda Q, k: maxsim_topk(corpus, Q, k, relu=relu),
                   lambda s, p, k: topk_merge(s, p, k), group)

    def search(self, Q: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        s, p = self.local_topk(Q, k)
        return self.exchange(s, p, k)

    def exchange(self, s: torch.Tensor, p: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """The one exchange step: all-gather of every rank's ``[B, k_local]`` list, then the merge.  Scores
        travel as their bit patterns next to the pids in ONE pre-shaped int64 buffer (one collective, no
        per-rank tensor list); a shard with fewer than ``k`` passages pads its list with (-inf, -1) entries,
        which the merge ignores."""
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return self.merge(s.unsqueeze(0), p.unsqueeze(0), k)
        B, kl = s.shape
        if kl < k:                                  # short shard (k > its passages, or an empty filter)
            s = torch.cat([s, s.new_full((B, k - kl), float("-inf"))], dim=1)
            p = torch.cat([p, p.new_full((B, k - kl), -1)], dim=1)
        payload = torch.empty((2, B, k), dtype=torch.int64, device=s.device)
        payload[0] = s.contiguous().view(torch.int32)
        payload[1] = p
        gathered = torch.empty((world, 2, B, k), dtype=torch.int64, device=s.device)
        if dist.get_backend(self.group) == "nccl":
            dist.all_gather_into_tensor(gathered, payload, group=self.group)
        else:                                        # gloo (CPU tests) has no flat all-gather
            dist.all_gather(list(gathered.unbind(0)), payload, group=self.group)
        gs = gathered[:, 0].to(torch.int32).view(torch.float32)
        return self.merge(gs, gathered[:, 1], k)
