"""Small invocations of the kernels added last (k_rp_fold, k_rp_fold_sum, k_msm_decompress with the overlapped sort, the
shuffle verifier's batches) as one small-case workload (odd counts, a tampered proof, both MSM paths); compute-sanitizer is closed on this pool, so
the checks are the verdicts and the equality of the two MSM paths."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    pkg = g.load_package()
    eng = pkg.Engine(0)
    gold = os.path.join(ROOT, "tests", "golden")
    rng = np.random.default_rng(3)
    for m in (1, 4, 16):
        per = m * 32 + eng.range_proof_bytes(m)
        rr = np.tile(np.fromfile(os.path.join(gold, "range_proofs_m%d.bin" % m), dtype=np.uint8).reshape(-1, per), (10, 1))[:37].copy()
        rr[5, m * 32 + 5 * 32] ^= 1
        st = eng.verify_range_proofs(rr[:, :m * 32].copy(), rr[:, m * 32:].copy(), m)
        assert int(st[5]) == 6 and int(st.astype(bool).sum()) == 1
    rec = np.tile(np.fromfile(os.path.join(gold, "shuffle_proofs.bin"), dtype=np.uint8).reshape(-1, 6432), (3, 1))[:9]
    st = eng.verify_shuffle(rec[:, :1152].copy(), rec[:, 1152:2304].copy(), rec[:, 2304:2656].copy(), rec[:, 2656:].copy())[0]
    assert not st.any()
    sc = rng.integers(0, 256, size=(5000, 32), dtype=np.uint8)
    sc[:, 31] &= 0x0f
    pts = eng.fixed_base(0, sc)[0]
    eng.msm_set_overlap(1 << 10, 30, 3)
    o1, s1 = eng.msm(sc, pts)
    eng.msm_set_overlap(1 << 10, 0, 3)
    o2, s2 = eng.msm(sc, pts)
    assert s1 == s2 == 0 and o1.tobytes() == o2.tobytes()
    eng.close()
    print("small-case workload done")


if __name__ == "__main__":
    main()
