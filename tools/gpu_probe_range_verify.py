"""Throughput of the batched Bulletproofs range-proof verification (qq_verify_range_proof_batch) on the committed valid
proofs (tests/golden/range_proofs_m*.bin), tiled to N proofs (BASELINE.json configs[3]).  Every proof must be accepted; a
second pass with one tampered proof in the middle must reject exactly that one (bisection)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def load(m, n):
    per = m * 32 + (9 + 2 * ((64 * m).bit_length() - 1)) * 32
    raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "range_proofs_m%d.bin" % m), dtype=np.uint8).reshape(-1, per)
    rec = np.tile(raw, ((n + raw.shape[0] - 1) // raw.shape[0], 1))[:n]
    return np.ascontiguousarray(rec[:, :m * 32]), np.ascontiguousarray(rec[:, m * 32:])


def main():
    pkg = g.load_package()
    eng = pkg.Engine(0)
    sizes = [int(x) for x in (sys.argv[1:] or ["1", "64", "4096"])]
    for m in (1, 4, 16):
        for n in sizes:
            import torch
            cm, pr = (torch.from_numpy(a).pin_memory().numpy() for a in load(m, n))
            eng.verify_set_transcripts(False)
            ts = []
            for rep in range(3):
                t = time.perf_counter()
                st = eng.verify_range_proofs(cm, pr, m)
                ts.append(time.perf_counter() - t)
            host_ms = min(ts[1:]) * 1e3
            eng.verify_set_transcripts(True)
            ts = []
            for rep in range(4):
                t = time.perf_counter()
                st = eng.verify_range_proofs(cm, pr, m)
                ts.append(time.perf_counter() - t)
            assert not st.any(), st
            best = min(ts[1:])
            terms = 2 * 64 * m + 2 + n * (4 + 2 * ((64 * m).bit_length() - 1) + m)
            line = {"probe": "verify_range_proofs", "m": m, "proofs": n, "wall_ms": best * 1e3, "proofs_per_s": n / best,
                    "values_per_s": n * m / best, "last_call_kernel_ms": eng.last_kernel_ms,
                    "breakdown_ms": eng.last_kernel_breakdown(), "wall_ms_host_transcripts": host_ms, "aggregate_msm_terms": terms,
                    "per_proof_msm_terms_in_the_reference": 2 * 64 * m + 2 * ((64 * m).bit_length() - 1) + m + 6, "all_accepted": True}
            eng.verify_set_aggregation(False)      # every transcript's own MSM (the reference's per-proof form) through the grouped MSM
            ts = []
            for rep in range(3):
                t = time.perf_counter()
                st = eng.verify_range_proofs(cm, pr, m)
                ts.append(time.perf_counter() - t)
            eng.verify_set_aggregation(True)
            assert not st.any()
            line["wall_ms_per_transcript_form"] = min(ts[1:]) * 1e3
            if n >= 3:
                pr2 = torch.from_numpy(pr.copy()).pin_memory().numpy()
                pr2[n // 2, 5 * 32 + 1] ^= 1
                ts = []
                for rep in range(3):      # the first repetition grows the failure path's scratch (cudaMalloc)
                    t = time.perf_counter()
                    st = eng.verify_range_proofs(cm, pr2, m)
                    ts.append(time.perf_counter() - t)
                line["one_bad_proof_wall_ms"] = min(ts) * 1e3
                assert st[n // 2] == 6 and int(st.astype(bool).sum()) == 1
                if n >= 1024:             # sixteen more, spread over the batch
                    for k in range(16):
                        pr2[(n // 16) * k + 3, 5 * 32 + 1] ^= 1
                    ts = []
                    for rep in range(2):
                        t = time.perf_counter()
                        st = eng.verify_range_proofs(cm, pr2, m)
                        ts.append(time.perf_counter() - t)
                    line["seventeen_bad_proofs_wall_ms"] = min(ts) * 1e3
                    assert int(st.astype(bool).sum()) == 17
            print(json.dumps(line), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
