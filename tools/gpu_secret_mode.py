"""Throughput cost of qq_set_secret_mode (masked full-table scans instead of index-addressed table reads): update_account at
2^18 accounts and fixed base at 2^20 scalars, both modes, outputs compared."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    eng = g.load_package().Engine(0)
    rng = np.random.default_rng(3)

    def sc(n):
        a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        a[:, 31] &= 0x0f
        return a
    n = 1 << 18
    acc = np.concatenate([eng.fixed_base(0, sc(n))[0] for _ in range(4)], axis=1).copy()
    bl, u, c = sc(n), sc(n), sc(n)
    nf = 1 << 20
    fs = sc(nf)
    res = {}
    for mode in (False, True):
        eng.set_secret_mode(mode)
        best = 1e30
        for rep in range(3):
            out, st = eng.update_account(acc, bl, u, c)
            if rep:
                best = min(best, eng.last_kernel_ms)
        bd = eng.last_kernel_breakdown()
        bestf = 1e30
        for rep in range(3):
            fo, fst = eng.fixed_base(0, fs)
            if rep:
                bestf = min(bestf, eng.last_kernel_ms)
        res[mode] = (out.copy(), fo.copy())
        print(json.dumps({"probe": "secret_mode", "secret": mode, "update_account_n": n, "update_account_kernel_ms": best,
                          "accounts_per_s": n / (best * 1e-3), "breakdown_ms": bd, "fixed_base_n": nf, "fixed_base_kernel_ms": bestf,
                          "fixed_base_per_s": nf / (bestf * 1e-3)}), flush=True)
    eng.set_secret_mode(False)
    print(json.dumps({"probe": "secret_mode", "outputs_identical": bool(np.array_equal(res[False][0], res[True][0]) and
                                                                        np.array_equal(res[False][1], res[True][1]))}))
    eng.close()


if __name__ == "__main__":
    main()
