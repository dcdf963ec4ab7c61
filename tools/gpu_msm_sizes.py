import os, sys, json, time
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT)
import __graft_entry__ as g
eng = g.load_package().Engine(0)
for n in (20000, 32768, 50000, 65536, 100000, 131072, 165890, 200000, 262144):
    rng = np.random.default_rng(n)
    hs = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); hs[:, 31] &= 0x0f
    a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); a[:, 31] &= 0x0f
    pts, st = eng.fixed_base(0, hs)
    best = 1e9
    for r in range(5):
        out, s = eng.msm(a, pts)
        bd = eng.last_kernel_breakdown()
        best = min(best, sum(bd.values()))
    print(json.dumps({"n": n, "kernel_ms": round(best, 4), "out": out.tobytes().hex()[:16]}))
