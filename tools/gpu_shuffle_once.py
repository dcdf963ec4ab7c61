"""One aggregate-form qq_verify_shuffle_batch call on N tiled golden proofs (for ncu captures of the transcript kernels)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "shuffle_proofs.bin"), dtype=np.uint8).reshape(-1, 6432)
rec = np.tile(raw, ((n + raw.shape[0] - 1) // raw.shape[0], 1))[:n]
arrs = [np.ascontiguousarray(rec[:, a:b]) for a, b in ((0, 1152), (1152, 2304), (2304, 2656), (2656, 6432))]
eng = g.load_package().Engine(0)
for _ in range(reps):
    st, sg, det = eng.verify_shuffle(*arrs)
assert not st.any()
print("ok", n, eng.last_kernel_breakdown())
eng.close()
