"""A few launches of the arg-max forward at the global-batch shape (8 queries x 832 rows against 128 ragged
documents of up to 512 tokens), for `ncu -k regex:flmr_argmax_tc` / compute-sanitizer captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ravqa_b200 import _cabi  # noqa: E402
from ravqa_b200.maxsim import maxsim_argmax  # noqa: E402
from train_step_probe import make  # noqa: E402

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 128
Q, D, mask, tok = make(8, n_docs, 832, 512, 0, torch.device("cuda", 0))
Qb, Db = Q.bfloat16(), D.bfloat16()
_cabi.lib().flmr_debug_set_argmax_path(2)
for _ in range(3):
    arg, rowmax = maxsim_argmax(Qb, Db, mask, return_rowmax=True)
torch.cuda.synchronize()
print("ok", tok, float(rowmax.sum()))
