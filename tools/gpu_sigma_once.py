"""A few qq_verify_update_account_dlog_batch calls (N proofs over 9 accounts, random valid points; for ncu launch lists)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
k = 9
eng = g.load_package().Engine(0)
rng = np.random.default_rng(9)


def scal(m):
    s = rng.integers(0, 256, size=(m, 32), dtype=np.uint8)
    s[:, 31] &= 0x0f
    return s


pts = [eng.fixed_base(0, scal(n * k))[0] for _ in range(8)]
ia = np.concatenate(pts[:4], axis=1).copy()
da = np.concatenate(pts[4:], axis=1).copy()
z, x = scal(n * k), scal(n)
import torch  # noqa: E402
for r in range(reps):
    if r == reps - 1:
        torch.cuda.profiler.start()
    st = eng.verify_update_account_dlog(ia, da, z, x, k)
torch.cuda.profiler.stop()
print("ok", n, int(st.sum()), eng.last_kernel_breakdown())
eng.close()
