"""Fixed-base throughput on the GPU as a function of the table window width (fixedbase_big.cuh).

Scalars resident in device memory, 32-byte compressed outputs written to device memory (qq_fixed_base_batch_dev);
time from the library's CUDA events.  One JSON line per (window, distribution).  Usage:
    python tools/gpu_fixed_base.py [log2_n] [W ...]        (W = 0 -> shared-memory 6-bit table only)
"""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

IMAD_FIXED_COMPRESSED = 91_400   # FIXED(4) + ENC, the cost-model figure for one compressed fixed-base mult


def main():
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    ws = [int(x) for x in sys.argv[2:]] or [0, 16, 20, 22, 24, 26]
    n = 1 << lg
    pkg = g.load_package()
    eng = pkg.Engine(0)
    peak = eng.measure_imad_peak()
    rng = np.random.default_rng(7)
    vp = ctypes.c_void_p
    s_full = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    s_full[:, 31] &= 0x0f
    s_bal = s_full.copy()
    s_bal[:, 8:] = 0                                   # 64-bit balances (the `bl` of update_account)
    d_s = eng.dev_alloc(n * 32)
    d_out = eng.dev_alloc(n * 32)
    d_st = eng.dev_alloc(n)
    ref = None
    for W in ws:
        t0 = time.time()
        try:
            eng.fixed_base_set_window(0, W)
        except Exception as e:  # out of memory for the widest tables on a busy device
            print(json.dumps({"window_bits": W, "error": str(e)}), flush=True)
            continue
        build_s = time.time() - t0
        for name, s in (("uniform252", s_full), ("balance64", s_bal)):
            eng.dev_upload(d_s, s)
            best = 1e30
            for rep in range(4):
                eng.call_dev("qq_fixed_base_batch_dev", ctypes.c_int(0), vp(d_s.value), vp(d_out.value), vp(d_st.value),
                             ctypes.c_size_t(n))
                if rep:
                    best = min(best, eng.last_kernel_ms)
            bd = eng.last_kernel_breakdown()
            out = eng.dev_download(d_out, n * 32)
            if name == "uniform252":
                if ref is None:
                    ref = out.copy()
                same = bool((out == ref).all())
            else:
                same = None
            rate = n / (best * 1e-3)
            print(json.dumps({"window_bits": W, "scalars": name, "n": n, "ms": best, "mults_per_s": rate,
                              "imad_model_frac": rate * IMAD_FIXED_COMPRESSED / peak["imad_lo_per_s"],
                              "table_build_s": build_s, "breakdown_ms": bd, "same_bytes_as_first_window": same}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
