"""One update_public_key and one generate_commitment call on n accounts (host API; for ncu launch lists)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
eng = g.load_package().Engine(0)
rng = np.random.default_rng(5)


def scal(m):
    s = rng.integers(0, 256, size=(m, 32), dtype=np.uint8)
    s[:, 31] &= 0x0f
    return s


pk = np.concatenate([eng.fixed_base(0, scal(n))[0] for _ in range(2)], axis=1).copy()
r, v = scal(n), scal(n)
import torch  # noqa: E402
for rep in range(2):
    if rep == 1:
        torch.cuda.profiler.start()
    o1, s1 = eng.update_public_key(pk, r)
    o2, s2 = eng.generate_commitment(pk, r, v)
torch.cuda.profiler.stop()
assert not s1.any() and not s2.any()
print("ok", n)
eng.close()
