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


class NcclExchange:
    """This rank's end of the C-level exchange (``flmr_comm_*`` / ``flmr_maxsim_topk_sharded`` /
    ``flmr_topk_exchange``): the library's own NCCL communicator, created from an id that rank 0 obtains and
    ``torch.distributed`` broadcasts once.  Per search: the fused scan, ONE grouped all-gather and the merge
    kernel, all enqueued by one C call on the current stream — no tensor packing or casting on the way."""

    def __init__(self, device: torch.device, group=None):
        import ctypes as C
        from . import _cabi
        L = _cabi.lib()
        self.device = device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_char * 128)()
            _cabi.check(L.flmr_comm_unique_id(C.cast(buf, C.c_void_p)))
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        backend = dist.get_backend(group)
        uid = uid.to(device) if backend == "nccl" else uid
        dist.broadcast(uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(uid.cpu().numpy().tobytes())
        self._h = C.c_void_p()
        with torch.cuda.device(device):
            _cabi.check(L.flmr_comm_create(C.c_char_p(raw), self.rank, self.world, int(device.index), C.byref(self._h)))

    def search(self, corpus, Q: torch.Tensor, k: int, relu: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """``flmr_maxsim_topk_sharded``: (scores [B, k], global pids [B, k]) of the MERGED ranking, on every rank."""
        import ctypes as C
        from . import _cabi
        from .maxsim import _prep_queries
        Qd = _prep_queries(corpus, Q)
        B, nq = Qd.size(0), Qd.size(1)
        s = torch.empty((B, k), dtype=torch.float32, device=self.device)
        p = torch.empty((B, k), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().flmr_maxsim_topk_sharded(
                corpus.handle, corpus.workspace(), self._h, C.c_void_p(Qd.data_ptr()), B, nq, k,
                _cabi.FLAG_RELU if relu else 0, C.c_void_p(s.data_ptr()), C.c_void_p(p.data_ptr()),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return s, p

    def exchange(self, s: torch.Tensor, p: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """``flmr_topk_exchange`` for lists the caller already has (``[B, k_in]``, pid < 0 = empty entry)."""
        import ctypes as C
        from . import _cabi
        s = s.to(self.device, torch.float32).contiguous()
        p = p.to(self.device, torch.int64).contiguous()
        B, k_in = s.shape
        out_s = torch.empty((B, k), dtype=torch.float32, device=self.device)
        out_p = torch.empty((B, k), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.lib().flmr_topk_exchange(
                self._h, C.c_void_p(s.data_ptr()), C.c_void_p(p.data_ptr()), B, k_in, k,
                C.c_void_p(out_s.data_ptr()), C.c_void_p(out_p.data_ptr()),
                C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out_s, out_p

    def close(self):
        from . import _cabi
        if self._h is not None:
            torch.cuda.synchronize(self.device)
            _cabi.lib().flmr_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


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
