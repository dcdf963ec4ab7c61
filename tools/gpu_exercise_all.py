"""Small-size pass over every C-ABI entry point (odd sizes, tiny tables): a quick crash / CUDA-error check, also usable
under `compute-sanitizer --tool memcheck` where that tool is available (it is closed on the build pool).  Results are
only checked for plausibility here; parity lives in tests/.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    pkg = g.load_package()
    eng = pkg.Engine(0)
    rng = np.random.default_rng(11)

    def scal(n):
        r = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        r[:, 31] &= 0x0f
        return r
    n = 37
    eng.fixed_base_set_window(0, 9)
    eng.fixed_base_set_window(1, 9)
    cols = [eng.fixed_base(0, scal(n))[0] for _ in range(4)]
    acc = np.concatenate(cols, axis=1).copy()
    big = eng.fixed_base(1, scal(3000))[0]          # big-table path (n >= 2048)
    assert big.any()
    out, st = eng.update_account(acc, scal(n), scal(n), scal(n))
    assert not st.any()
    eng.verify_account(acc, scal(n), scal(n))
    eng.update_public_key(acc[:, :64], scal(n))
    eng.verify_public_key_update(acc[:, :64], acc[:, :64], scal(n))
    eng.generate_commitment(acc[:, :64], scal(n), scal(n))
    eng.add_commitments(acc[:, 64:], acc[:, 64:], negate_b=True)
    eng.mul_commitment(acc[:, 64:], scal(n))
    eng.delta_epsilon(acc, scal(n), scal(n), np.frombuffer(pkg.RistrettoPublicKey.generate_base_pk().as_bytes(), np.uint8))
    eng.delta_identity_check(acc)
    eng.decommit(acc[:, 64:], scal(n))
    eng.decommit_value(acc[:3, 64:], scal(3), 21)
    for m in (1, 5, 300, 3000):
        pts = eng.fixed_base(0, scal(m))[0]
        o, s = eng.msm(scal(m), pts)
        assert s == 0
        h = eng.msm_points_prepare(pts)
        o2, s2 = eng.msm_prepared(scal(m), h)
        eng.msm_points_free(h)
    ks = np.array([2, 3, 9, 4, 12, 1, 0, 2] * 5, dtype=np.uint32)
    offs = np.zeros(ks.size + 1, np.uint32)
    offs[1:] = np.cumsum(ks)
    nt = int(offs[-1])
    eng.msm_segmented(scal(nt), eng.fixed_base(0, scal(nt))[0], offs)
    eng.from_uniform_bytes(rng.integers(0, 256, size=(9, 64), dtype=np.uint8))
    eng.vector_pedersen_gens(4)
    eng.bulletproof_gens(4, 2)
    xyzt = np.zeros((2, 128), np.uint8)
    xyzt[:, 32] = 1
    xyzt[:, 64] = 1
    eng.points_sum(xyzt)
    # the batched verifiers on the committed golden proofs (odd counts), and the MSM tuning knob
    gold = os.path.join(ROOT, "tests", "golden")
    rec = np.fromfile(os.path.join(gold, "shuffle_proofs.bin"), dtype=np.uint8).reshape(-1, 6432)[:3]
    st = eng.verify_shuffle(rec[:, :1152].copy(), rec[:, 1152:2304].copy(), rec[:, 2304:2656].copy(), rec[:, 2656:].copy())[0]
    assert not st.any()
    for m in (1, 4, 16):
        per = m * 32 + eng.range_proof_bytes(m)
        rr = np.fromfile(os.path.join(gold, "range_proofs_m%d.bin" % m), dtype=np.uint8).reshape(-1, per)[:3]
        assert not eng.verify_range_proofs(rr[:, :m * 32].copy(), rr[:, m * 32:].copy(), m).any()
    eng.msm_set_overlap(1 << 10, 40, 2)
    pts = eng.fixed_base(0, scal(3000))[0]
    assert eng.msm(scal(3000), pts)[1] == 0
    eng.msm_set_overlap()
    eng.close()
    print("exercise pass done: every entry point returned without a CUDA error")


if __name__ == "__main__":
    main()
