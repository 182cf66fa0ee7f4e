"""CPU: the flat-store Indexer: colbert.Indexer.index semantics (overwrite / reuse /
resume), chunk ownership across ranks, shard-range loading."""
import json
import os
import zlib

import numpy as np
import pytest
import torch

from ravqa_b200.index_io import chunk_exists, load_flat_index, save_flat_index
from ravqa_b200.indexer import Indexer


def fake_encoder(calls):
    def encode(passages):
        calls.append(len(passages))
        doclens = [3 + (len(p) % 5) for p in passages]
        g = torch.Generator().manual_seed(sum(doclens))
        embs = torch.nn.functional.normalize(torch.randn(sum(doclens), 128, generator=g), dim=-1)
        # make embeddings a pure function of the passage text so chunks are reproducible
        rows = []
        for p, n in zip(passages, doclens):
            gp = torch.Generator().manual_seed(zlib.crc32(p.encode()))    # (str hashes differ between processes)
            rows.append(torch.nn.functional.normalize(torch.randn(n, 128, generator=gp), dim=-1))
        return torch.cat(rows), doclens
    return encode


def test_index_chunks_resume_and_load(tmp_path):
    passages = ["passage number %d %s" % (i, "x" * (i % 7)) for i in range(1037)]
    calls = []
    ix = Indexer(encode_fn=fake_encoder(calls), index_root=str(tmp_path), chunksize=250)
    path = ix.index("temp_index.nbits=8", passages, overwrite=True)
    meta = json.load(open(os.path.join(path, "metadata.json")))
    assert meta["n_passages"] == 1037 and meta["num_chunks"] == 5 and calls == [250, 250, 250, 250, 37]
    tokens, doclens, _ = load_flat_index(path)
    ref_t, ref_d = fake_encoder([])(passages)
    assert doclens.tolist() == ref_d and torch.equal(tokens, ref_t.to(torch.bfloat16))
    # shard loading reads exactly the requested passages
    t2, d2, _ = load_flat_index(path, passage_range=(240, 777))
    off = np.concatenate([[0], np.cumsum(ref_d)])
    assert d2.tolist() == ref_d[240:777] and torch.equal(t2, tokens[off[240]:off[777]])
    # overwrite=False refuses an existing index, 'reuse' keeps it without re-encoding
    with pytest.raises(AssertionError):
        ix.index("temp_index.nbits=8", passages, overwrite=False)
    calls.clear()
    assert ix.index("temp_index.nbits=8", passages, overwrite="reuse") == path and calls == []
    # resume: delete one chunk -> only that chunk is re-encoded
    os.remove(os.path.join(path, "2.metadata.json"))
    assert not chunk_exists(path, 2)
    ix.index("temp_index.nbits=8", passages, overwrite="resume")
    assert calls == [250]
    t3, d3, _ = load_flat_index(path)
    assert torch.equal(t3, tokens) and d3.tolist() == ref_d


def _two_rank_worker(rank, port, root, ret):
    import time
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=2)
    passages = ["p%d" % i for i in range(90)]
    if rank == 0:
        time.sleep(1.0)          # rank 1 reaches index() first: it must wait for rank 0's erase, not be erased by it
    ix = Indexer(encode_fn=fake_encoder([]), index_root=root, chunksize=20, rank=rank, nranks=2)
    path = ix.index("idx", passages, overwrite=True)
    ret[rank] = os.path.exists(os.path.join(path, "metadata.json"))      # nobody returns before the index is final
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_ranks_and_single_file_format(tmp_path):
    """Two ranks build one index (round-robin chunk ownership) over a directory that already holds a stale index:
    rank 0 erases and finalizes, rank 1 waits for the erase and rank 0 for rank 1's chunks (the reference does the
    erase in the parent before launching its workers, colbert/indexer.py:58-84)."""
    import socket
    import torch.multiprocessing as mp
    passages = ["p%d" % i for i in range(90)]
    stale = os.path.join(str(tmp_path), "idx")
    save_flat_index(stale, torch.zeros(7, 128), [3, 4])                   # what overwrite=True must remove
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ret = mp.Manager().dict()
    mp.spawn(_two_rank_worker, args=(port, str(tmp_path), ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
    with pytest.raises(RuntimeError, match="process group"):              # ranks cannot be faked one after the other
        Indexer(encode_fn=fake_encoder([]), index_root=str(tmp_path), rank=1, nranks=2).index("y", passages)
    tokens, doclens, meta = load_flat_index(os.path.join(str(tmp_path), "idx"))
    ref_t, ref_d = fake_encoder([])(passages)
    assert meta["num_chunks"] == 5 and doclens.tolist() == ref_d and torch.equal(tokens, ref_t.to(torch.bfloat16))
    p = save_flat_index(os.path.join(str(tmp_path), "single"), ref_t, ref_d)
    t, d, m = load_flat_index(p, passage_range=(10, 20))
    off = np.concatenate([[0], np.cumsum(ref_d)])
    assert m["num_chunks"] == 0 and torch.equal(t, ref_t.to(torch.bfloat16)[off[10]:off[20]])
    with pytest.raises(RuntimeError):
        Indexer(index_root=str(tmp_path)).index("x", passages)
