"""Host / device overlap for the batched proof verifiers with the C ABI as it is ("one qq_ctx per caller thread"): two contexts
on the same GPU, two caller threads, each verifying slices of the batch - while one thread waits for its GPU batch the other
runs its Merlin transcripts on the host cores.  Prints one JSON line per configuration."""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__ as g  # noqa: E402
from gpu_probe_shuffle_verify import load as load_shuffle  # noqa: E402
from gpu_probe_range_verify import load as load_range  # noqa: E402


def run(engines, work, nslices):
    """work(engine, lo, hi) for nslices slices, round-robin over the engines' threads; returns wall seconds."""
    n = work.n
    bounds = [(n * i // nslices, n * (i + 1) // nslices) for i in range(nslices)]
    errs = []

    def worker(k):
        try:
            for i in range(k, nslices, len(engines)):
                work(engines[k], *bounds[i])
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    t = time.perf_counter()
    th = [threading.Thread(target=worker, args=(k,)) for k in range(len(engines))]
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t
    if errs:
        raise errs[0]
    return dt


def main():
    pkg = g.load_package()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    engines = [pkg.Engine(0), pkg.Engine(0)]
    si, so, stm, pr = load_shuffle(n)
    cm, rp = load_range(16, n)

    def shuffle_work(e, lo, hi):
        st = e.verify_shuffle(si[lo:hi], so[lo:hi], stm[lo:hi], pr[lo:hi])[0]
        assert not st.any()

    def range_work(e, lo, hi):
        st = e.verify_range_proofs(cm[lo:hi], rp[lo:hi], 16)
        assert not st.any()
    shuffle_work.n = range_work.n = n
    for name, work in (("verify_shuffle", shuffle_work), ("verify_range_proofs_m16", range_work)):
        for nctx, nslices in ((1, 1), (1, 4), (2, 2), (2, 4), (2, 8)):
            best = min(run(engines[:nctx], work, nslices) for _ in range(4))
            print(json.dumps({"probe": name, "proofs": n, "contexts": nctx, "slices": nslices, "wall_ms": best * 1e3,
                              "proofs_per_s": n / best}), flush=True)
    for e in engines:
        e.close()


if __name__ == "__main__":
    main()
