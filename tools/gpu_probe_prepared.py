"""qq_msm_prepared over the shifted form (one bucket set for all windows, no Horner chain) against the plain prepared form, by size."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    import torch
    eng = g.load_package().Engine(0)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(1)
    vp = ctypes.c_void_p
    for lg in [int(x) for x in (sys.argv[1:] or ["10", "11", "12", "14", "16", "18", "20"])]:
        n = 1 << lg
        sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        sc[:, 31] &= 0x0f
        pts, _ = eng.fixed_base(0, sc)
        a = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
        a[:, 31] &= 0x0f
        pts_d = torch.from_numpy(pts.reshape(-1)).to(dev)
        a_d = torch.from_numpy(a.reshape(-1)).to(dev)
        small = torch.zeros(256, dtype=torch.uint8, device=dev)
        h = ctypes.c_void_p()
        eng._ck(eng.lib.qq_msm_points_prepare_dev(eng.h, vp(pts_d.data_ptr()), ctypes.c_size_t(n), ctypes.byref(h)), "prepare")
        line = {"probe": "msm_prepared", "points": n, "shifted_bytes": int(eng.lib.qq_msm_points_shifted_bytes(h))}
        outs = {}
        for use in (True, False):
            eng.msm_set_shifted(use_it=use)
            for _ in range(3):
                eng.call_dev("qq_msm_prepared_dev", vp(a_d.data_ptr()), h, ctypes.c_size_t(n), vp(small.data_ptr()), vp(small.data_ptr() + 64))
            eng.event_record(0)
            reps = 10
            for _ in range(reps):
                eng.call_dev("qq_msm_prepared_dev", vp(a_d.data_ptr()), h, ctypes.c_size_t(n), vp(small.data_ptr()), vp(small.data_ptr() + 64))
            eng.event_record(1)
            line["ms_shifted" if use else "ms_plain"] = eng.event_elapsed_ms(0, 1) / reps
            line["breakdown_shifted" if use else "breakdown_plain"] = {k: round(v, 3) for k, v in eng.last_kernel_breakdown().items() if v}
            outs[use] = small.cpu().numpy()[:32].copy()
        line["same_result"] = bool((outs[True] == outs[False]).all())
        eng.msm_set_shifted(use_it=True)
        eng.lib.qq_msm_points_free(eng.h, h)
        print(json.dumps(line), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
