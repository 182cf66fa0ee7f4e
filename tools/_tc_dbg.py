import sys, torch
sys.path.insert(0, '/root/repo')
import ravqa_b200 as R
from ravqa_b200 import _cabi
from ravqa_b200.maxsim import maxsim_argmax
sys.path.insert(0, '/root/repo/tools')
from train_step_probe import make
L = _cabi.lib()
dev = torch.device('cuda', 0)
for (B, n) in ((8, 24), (8, 32), (8, 64), (8, 128)):
    Q, D, mask, tok = make(B, n, 832, 512, 0, dev)
    Qb, Db = Q.bfloat16(), D.bfloat16()
    L.flmr_debug_set_argmax_path(1)
    a1, m1 = maxsim_argmax(Qb, Db, mask, return_rowmax=True)
    L.flmr_debug_set_argmax_path(2)
    try:
        a2, m2 = maxsim_argmax(Qb, Db, mask, return_rowmax=True)
        torch.cuda.synchronize()
        print(B, n, 'ok', (a1 == a2).float().mean().item(), (m1 - m2).abs().max().item(), flush=True)
    except Exception as e:
        print(B, n, 'ERR', repr(e)[:500], flush=True)
        break
