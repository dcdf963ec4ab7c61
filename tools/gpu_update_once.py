"""One qq_update_account_batch call on N synthetic accounts (for ncu captures of k_varbase_split and friends)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = g.load_package().Engine(0)
rng = np.random.default_rng(5)


def sc(k):
    a = rng.integers(0, 256, size=(k, 32), dtype=np.uint8)
    a[:, 31] &= 0x0f
    return a


acc = np.concatenate([eng.fixed_base(0, sc(n))[0] for _ in range(4)], axis=1).copy()
bl, u, c = sc(n), sc(n), sc(n)
for _ in range(reps):
    out, st = eng.update_account(acc, bl, u, c)
assert not st.any()
print("ok", n, eng.last_kernel_breakdown())
eng.close()
