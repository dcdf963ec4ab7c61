"""Config 3 shape: 4096 shuffle proofs x ~42 small MSMs (2-9 terms) -> timing of qq_msm_segmented."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

# per-proof MSM job list (terms per MSM), from the App. C tally: commit4 x ~12, pc2 x 6, vt(2) x 2, vt(3) x 3, const(3),
# vt(6) x 6, vt(7), const(9) x 6, const(2), commit3  -> 42 MSMs, ~205 terms
JOBS = [4] * 12 + [2] * 6 + [2] * 2 + [3] * 3 + [3] + [6] * 6 + [7] + [9] * 6 + [2] + [3] + [4] * 3


def main():
    pkg = g.load_package()
    eng = pkg.Engine(0)
    rng = np.random.default_rng(3)
    proofs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    ks = np.array(JOBS * proofs, dtype=np.uint32)
    offs = np.zeros(ks.size + 1, np.uint32)
    offs[1:] = np.cumsum(ks)
    nt = int(offs[-1])
    raw = rng.integers(0, 256, size=(nt, 32), dtype=np.uint8)
    raw[:, 31] &= 0x0f
    pts, _ = eng.fixed_base(0, raw)
    sc = rng.integers(0, 256, size=(nt, 32), dtype=np.uint8)
    sc[:, 31] &= 0x0f
    for rep in range(3):
        t = time.time()
        out, st = eng.msm_segmented(sc, pts, offs)
        dt = time.time() - t
    assert not st.any()
    print(json.dumps({"probe": "shuffle_msm_jobs", "proofs": proofs, "msms": int(ks.size), "terms": nt, "wall_s": dt,
                      "proofs_per_s": proofs / dt, "terms_per_s": nt / dt, "kernel_ms": eng.last_kernel_ms,
                      "breakdown_ms": eng.last_kernel_breakdown()}))


if __name__ == "__main__":
    main()
