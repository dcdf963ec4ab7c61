"""Throughput of the batched ShuffleProof::verify (qq_verify_shuffle_batch) on the committed valid proofs
(tests/golden/shuffle_proofs.bin), tiled to N proofs (BASELINE.json configs[2]: 4096).  Every proof must be accepted."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def load(n):
    raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "shuffle_proofs.bin"), dtype=np.uint8).reshape(-1, 6432)
    rec = np.tile(raw, ((n + raw.shape[0] - 1) // raw.shape[0], 1))[:n]
    return (np.ascontiguousarray(rec[:, :1152]), np.ascontiguousarray(rec[:, 1152:2304]),
            np.ascontiguousarray(rec[:, 2304:2656]), np.ascontiguousarray(rec[:, 2656:]))


def main():
    pkg = g.load_package()
    eng = pkg.Engine(0)
    import torch
    for n in [int(x) for x in (sys.argv[1:] or ["1", "64", "512", "4096", "16384"])]:
        # page-locked host buffers, as a caller that feeds serialized transactions would hold them
        arrs = [torch.from_numpy(a).pin_memory().numpy() for a in load(n)]
        for mode in ("aggregate", "device", "host"):
            eng.verify_set_transcripts(mode != "host")
            eng.verify_set_aggregation(mode == "aggregate")
            ts = []
            for rep in range(5):
                t = time.perf_counter()
                st, sg, det = eng.verify_shuffle(*arrs)
                ts.append(time.perf_counter() - t)
            assert not st.any()
            best = min(ts[1:])
            print(json.dumps({"probe": "verify_shuffle", "transcripts": mode, "proofs": n, "wall_ms": best * 1e3,
                              "proofs_per_s": n / best, "kernel_ms": eng.last_kernel_ms, "breakdown_ms": eng.last_kernel_breakdown(),
                              "msms": 32 * n, "terms": 239 * n, "all_accepted": True}), flush=True)
        eng.verify_set_transcripts(True)
        eng.verify_set_aggregation(True)
        if n >= 64:      # one tampered proof (an output account): the aggregate fails, the slice is verified in the exact form
            bad = [torch.from_numpy(a.copy()).pin_memory().numpy() for a in arrs]
            bad[1][n // 2, 5] ^= 1
            ts = []
            for rep in range(3):      # the first repetition grows the workspace (cudaMalloc)
                t = time.perf_counter()
                st, sg, det = eng.verify_shuffle(*bad)
                ts.append(time.perf_counter() - t)
            assert np.nonzero(st)[0].tolist() == [n // 2]
            print(json.dumps({"probe": "verify_shuffle", "transcripts": "aggregate, one tampered proof (grouped check + exact form for its group)",
                              "proofs": n, "wall_ms": min(ts) * 1e3, "first_call_ms": ts[0] * 1e3}), flush=True)
            if n >= 1024:             # sixteen tampered proofs spread over the batch
                for k in range(16):
                    bad[1][(n // 16) * k + 3, 5] ^= 1
                ts = []
                for rep in range(2):
                    t = time.perf_counter()
                    st, sg, det = eng.verify_shuffle(*bad)
                    ts.append(time.perf_counter() - t)
                assert int(np.count_nonzero(st)) == 17
                print(json.dumps({"probe": "verify_shuffle", "transcripts": "aggregate, 17 tampered proofs", "proofs": n, "wall_ms": min(ts) * 1e3}),
                      flush=True)
    eng.close()


if __name__ == "__main__":
    main()
