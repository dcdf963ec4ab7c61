"""Quick on-GPU probe: integer-pipe peak, per-kernel-family timing of the account pipeline. Writes JSON lines."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

L = 2**252 + 27742317777372353535851937790883648493


def rand_scalars(rng, n):
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    raw[:, 31] &= 0x0f  # < 2^252 < l : canonical, uniform enough for timing
    return raw


def make_accounts(eng, rng, n):
    """valid accounts from fixed-base mults: pk = (rho*B, (sk*rho)*B), comm = (k*B, k2*B) (any valid points)."""
    cols = [eng.fixed_base(0, rand_scalars(rng, n))[0] for _ in range(4)]
    return np.concatenate(cols, axis=1).copy()


def main():
    pkg = g.load_package()
    eng = pkg.Engine(0)
    print(json.dumps({"probe": "imad_peak", **eng.measure_imad_peak(), "sms": eng.sm_count}), flush=True)
    rng = np.random.default_rng(1)
    sizes = [int(x) for x in (sys.argv[1:] or ["16384", "131072"])]
    for n in sizes:
        acc = make_accounts(eng, rng, n)
        bl, u, c = rand_scalars(rng, n), rand_scalars(rng, n), rand_scalars(rng, n)
        for rep in range(2):
            t = time.time()
            out, st = eng.update_account(acc, bl, u, c)
            dt = time.time() - t
        assert not st.any()
        print(json.dumps({"probe": "update_account_host", "n": n, "wall_s": dt, "accounts_per_s": n / dt,
                          "kernel_ms": eng.last_kernel_ms, "breakdown_ms": eng.last_kernel_breakdown()}), flush=True)
        t = time.time()
        o2, st2 = eng.generate_commitment(acc[:, :64].copy(), u, bl)
        dt = time.time() - t
        print(json.dumps({"probe": "generate_commitment_host", "n": n, "wall_s": dt, "per_s": n / dt,
                          "kernel_ms": eng.last_kernel_ms, "breakdown_ms": eng.last_kernel_breakdown()}), flush=True)
        t = time.time()
        o3, st3 = eng.fixed_base(0, u)
        dt = time.time() - t
        print(json.dumps({"probe": "fixed_base_host", "n": n, "wall_s": dt, "per_s": n / dt,
                          "kernel_ms": eng.last_kernel_ms, "breakdown_ms": eng.last_kernel_breakdown()}), flush=True)
        t = time.time()
        st4 = eng.verify_account(acc, u, bl)
        dt = time.time() - t
        print(json.dumps({"probe": "verify_account_host", "n": n, "wall_s": dt, "per_s": n / dt,
                          "kernel_ms": eng.last_kernel_ms, "breakdown_ms": eng.last_kernel_breakdown()}), flush=True)
    n = 4096
    pts, _ = eng.fixed_base(0, rand_scalars(rng, n))
    t = time.time()
    o, s = eng.msm(rand_scalars(rng, n), pts)
    dt = time.time() - t
    print(json.dumps({"probe": "msm_naive", "n": n, "wall_s": dt, "kernel_ms": eng.last_kernel_ms,
                      "breakdown_ms": eng.last_kernel_breakdown()}), flush=True)
    print(json.dumps({"probe": "launches", "count": eng.launch_count}))


if __name__ == "__main__":
    main()
