"""A few qq_verify_range_proof_batch calls on N tiled golden proofs of m values (for ncu captures)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
per = m * 32 + (9 + 2 * ((64 * m).bit_length() - 1)) * 32
raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "range_proofs_m%d.bin" % m), dtype=np.uint8).reshape(-1, per)
rec = np.tile(raw, ((n + raw.shape[0] - 1) // raw.shape[0], 1))[:n]
cm, pr = np.ascontiguousarray(rec[:, :m * 32]), np.ascontiguousarray(rec[:, m * 32:])
eng = g.load_package().Engine(0)
import torch  # noqa: E402
for r in range(reps):
    if r == reps - 1:
        torch.cuda.profiler.start()      # ncu --profile-from-start off: only the last repetition is captured
    st = eng.verify_range_proofs(cm, pr, m)
torch.cuda.profiler.stop()
assert not st.any()
print("ok", n, m, eng.last_kernel_breakdown())
eng.close()
