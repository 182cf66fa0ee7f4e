"""GPU: the end-to-end drop-in flow of examples/retrieval_eval_flow.py (index -> batched search -> Recall@K)."""
import importlib.util
import os

import pytest

from helpers import ROOT

pytestmark = pytest.mark.gpu


def test_retrieval_eval_flow(tmp_path):
    spec = importlib.util.spec_from_file_location("flow", os.path.join(ROOT, "examples", "retrieval_eval_flow.py"))
    flow = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(flow)
    import numpy as np
    from oracle import maxsim_oracle as O
    from ravqa_b200.index_io import load_flat_index
    recall, ranking, (path, Q, gold) = flow.run(n_passages=3000, n_queries=16, index_root=str(tmp_path), verbose=False)
    assert recall["Recall@10"] >= 0.85 and recall["Recall@100"] >= recall["Recall@10"] >= recall["Recall@1"]
    assert all(len(v) == 100 and [r for _, r, _ in v] == list(range(1, 101)) for v in ranking.values())
    # the ranking IS the exhaustive MaxSim ranking of the stored (bf16) embeddings
    tokens, doclens, _ = load_flat_index(path)
    ref = O.topk(O.maxsim_scores(O.bf16_round(Q[:3].numpy()), tokens.float().numpy(), doclens), 10)[1]
    for i in range(3):
        assert [pid for pid, _, _ in ranking["q%d" % i][:10]] == ref[i].tolist()
