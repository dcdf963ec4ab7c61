"""On-GPU probe of small batches (BASELINE.json configs[0] and a block's worth of transactions): update_account,
verify_account, update_public_key through the host API for n = 9 ... 16384 with the four-lane cooperative
variable-base kernel off / default / forced.  Median wall time per call and the kernel-family breakdown; outputs of
the three settings must be byte-identical.  Writes JSON lines."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
from tools.gpu_probe import make_accounts, rand_scalars  # noqa: E402


def median_ms(fn, reps):
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t) * 1e3)
    return float(np.median(ts))


def main():
    pkg = g.load_package()
    eng = pkg.Engine(0)
    rng = np.random.default_rng(3)
    sizes = [int(x) for x in (sys.argv[1:] or ["9", "100", "1000", "2368", "4736", "9472", "16384"])]
    for n in sizes:
        acc = make_accounts(eng, rng, n)
        bl, u, c = rand_scalars(rng, n), rand_scalars(rng, n), rand_scalars(rng, n)
        ref = None
        for name, limit in (("off", 0), ("default", -1), ("forced", 1 << 40)):
            eng.varbase_set_coop_limit(limit)
            out, st = eng.update_account(acc, bl, u, c)
            assert not st.any()
            if ref is None:
                ref = out.copy()
            same = bool((out == ref).all())
            reps = 20 if n <= 4736 else 8
            ms = median_ms(lambda: eng.update_account(acc, bl, u, c), reps)
            bd = eng.last_kernel_breakdown()
            ms_v = median_ms(lambda: eng.verify_account(acc, u, bl), reps)
            ms_k = median_ms(lambda: eng.update_public_key(acc[:, :64].copy(), u), reps)
            pk = acc[:, :64].copy()
            ms_g = median_ms(lambda: eng.generate_commitment(pk, c, bl), reps)
            print(json.dumps({"probe": "small_batch", "n": n, "coop": name, "update_account_ms": ms,
                              "breakdown_ms": bd, "verify_account_ms": ms_v, "update_public_key_ms": ms_k,
                              "generate_commitment_ms": ms_g, "same_output": same}), flush=True)
        eng.varbase_set_coop_limit(-1)


if __name__ == "__main__":
    main()
