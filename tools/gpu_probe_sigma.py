"""Throughput of the batched sigma-proof verifier (qq_verify_update_account_dlog_batch): B proofs x n accounts, random
valid points (the verdict is "reject" for all of them; the work is identical to that of valid proofs)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 9
    pkg = g.load_package()
    eng = pkg.Engine(0)
    rng = np.random.default_rng(4)

    def sc(m):
        r = rng.integers(0, 256, size=(m, 32), dtype=np.uint8)
        r[:, 31] &= 0x0f
        return r
    items = B * n
    ia = np.concatenate([eng.fixed_base(0, sc(items))[0] for _ in range(4)], axis=1).copy()
    da = np.concatenate([eng.fixed_base(0, sc(items))[0] for _ in range(4)], axis=1).copy()
    z, x = sc(items), sc(B)
    for mode in ("host", "device"):
        eng.verify_set_transcripts(mode == "device")
        ts = []
        for rep in range(6):
            t = time.perf_counter()
            st = eng.verify_update_account_dlog(ia, da, z, x, n)
            ts.append(time.perf_counter() - t)
        dt = float(np.median(ts[1:]))
        print(json.dumps({"probe": "verify_update_account_dlog", "transcripts": mode, "proofs": B, "accounts_per_proof": n,
                          "msms": 2 * items, "wall_s": dt, "proofs_per_s": B / dt, "kernel_ms": eng.last_kernel_ms,
                          "all_rejected": bool((st == 6).all())}))
    # one proof at a time (the latency a single verification sees), small-batch paths on / off
    for name, limit in (("off", 0), ("default", -1)):
        eng.varbase_set_coop_limit(limit)
        ts = []
        for rep in range(12):
            t = time.perf_counter()
            st1 = eng.verify_update_account_dlog(ia[:n], da[:n], z[:n], x[:1], n)
            ts.append((time.perf_counter() - t) * 1e3)
        print(json.dumps({"probe": "verify_update_account_dlog_single", "small_batch_paths": name, "accounts_per_proof": n,
                          "msms": 2 * n, "ms_median": float(np.median(ts)), "kernel_ms": eng.last_kernel_ms,
                          "same_verdict": bool(st1[0] == st[0])}))
    eng.varbase_set_coop_limit(-1)
    eng.close()


if __name__ == "__main__":
    main()
