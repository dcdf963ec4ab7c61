#!/usr/bin/env python
"""Benchmark of the quisquis hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, libqq_b200.so)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

A "step" = Account::update_account over one batch of 2^20 synthetic accounts per GPU (BASELINE.json configs[1]);
the JSON line also carries the 2^20-point MSM (configs[3]) as `msm`.  Metric: account updates / s (whole job).

Timing: W warm-up steps, then exactly K steps bracketed by barrier + synchronize; device time is taken with CUDA
events recorded on the library's own stream (qq_event_record), max over ranks.  Inputs (224 MB) and outputs (128 MB)
per step exceed the 126 MB L2, so no L2 flush is needed between steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work per unit (SURVEY.md App. B / BASELINE.md section 4), in thread-level IMAD instructions
# (a 32x32->64 product = 2: IMAD.WIDE and IMAD.HI issue at half the IMAD rate on sm_100a, measured)
IMAD_PER_UPDATE_ACCOUNT = 1_438_000
IMAD_PER_VARBASE = 289_000
# what k_varbase_split actually executes per account (2 points x 2 scalars through 4 quarter tables, scalarmult.cuh):
# 2 x (1248 S + 2253 M) with M = 144, S = 88 IMAD units -- 75 % of the 4 x 289 000 the cost model charges
IMAD_EXECUTED_VARBASE_PER_ACCOUNT = 2 * (1248 * 88 + 2253 * 144)
NCU_TRAFFIC_BYTES_PER_LAUNCH = 37.96e9
IMAD_PER_FIXED_COMPRESSED = 91_400  # FIXED(4) + ENC
IMAD_PER_MSM_POINT = 43_900        # compressed input, n = 2^20, c = 16
BYTES_PER_UPDATE_ACCOUNT = 224 + 128 + 1
L = 2**252 + 27742317777372353535851937790883648493


def rand_scalars(rng, n):
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    raw[:, 31] &= 0x0f          # < 2^252 < l: canonical, uniform over 252 bits ("worst case" distribution A)
    return raw


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region (one persistent process)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(3)
            except Exception:
                self.proc.kill()

    def summary(self, t0=None, t1=None):
        samples = []
        for ts, line in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.15):
                continue
            parts = [x.strip() for x in line.strip().split(",")]
            if len(parts) >= 7:
                try:
                    float(parts[0])
                    samples.append(parts)
                except ValueError:
                    pass
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(s[0]) for s in samples)
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in samples:
            for nme, v in zip(names, s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": float(samples[0][1]),
                "power_w_max": max(float(s[2]) for s in samples), "reasons": sorted(reasons), "samples": len(samples)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# =====================================================================================================================
# reference arm: the reference's own CPU implementation of the path = curve25519-dalek's algorithms, restated in
# oracle/qq_oracle.c (dalek cannot be built here: no cargo/rustc, not vendored), all host threads.
# =====================================================================================================================
def run_reference(args):
    world, rank, local = dist_setup(args.gpus)
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle as C
    C.set_threads(len(os.sched_getaffinity(0)))   # torchrun exports OMP_NUM_THREADS=1; the baseline uses every host core
    cores = C.threads()
    rng = np.random.default_rng(1234)
    # calibrate a bounded sample: ~2 s of all-core work per step
    m0 = 64 * cores
    cols = [C.fixed_base(0, rand_scalars(rng, m0))[0] for _ in range(4)]
    acc0 = np.concatenate(cols, axis=1).copy()
    t = time.time()
    C.update_account(acc0, rand_scalars(rng, m0), rand_scalars(rng, m0), rand_scalars(rng, m0))
    rate0 = m0 / (time.time() - t)
    m = int(max(m0, min(1 << 20, rate0 * args.ref_seconds_per_step)))
    reps = (m + m0 - 1) // m0
    acc = np.tile(acc0, (reps, 1))[:m].copy()
    bl, u, c = rand_scalars(rng, m), rand_scalars(rng, m), rand_scalars(rng, m)
    for _ in range(args.warmup):
        C.update_account(acc[:m0], bl[:m0], u[:m0], c[:m0])
    t = time.time()
    for _ in range(args.steps):
        out, st = C.update_account(acc, bl, u, c)
    dt = time.time() - t
    assert not st.any()
    value = m * args.steps / dt
    # the second half of the metric: Ristretto MSM points/s at 2^20 (compressed points in, decompression included), all cores
    msm_ref = None
    if args.msm_points > 0:
        mm = args.msm_points
        pts = np.concatenate([C.fixed_base(0, rand_scalars(rng, min(1 << 16, mm - lo)))[0] for lo in range(0, mm, 1 << 16)])
        a = rand_scalars(rng, mm)
        C.msm(a[:4096], pts[:4096])
        t = time.time()
        mo, ms_ = C.msm(a, pts)
        dtm = time.time() - t
        assert ms_ == 0
        msm_ref = {"points": mm, "ms": dtm * 1e3, "points_per_sec": mm / dtm, "cores": cores, "kind": "port",
                   "sample": "one %d-point MSM (oracle/qq_oracle.c oq_msm: decode + Pippenger per thread slice + sum), %d host threads" % (mm, cores)}
    line = {
        "impl": "reference", "metric": "account_updates_per_sec", "value": value, "unit": "accounts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64 (radix-2^51 limbs, 128-bit products)",
        "data": "synthetic",
        "config": {"workload": "Account::update_account, 2^20 accounts per GPU (BASELINE.json configs[1]); this arm times "
                               "a bounded sample of %d accounts per step" % m,
                   "scalars": "uniform 252-bit"},
        "cpu_baseline": {"value": value, "unit": "accounts/s", "cores": cores, "kind": "port",
                         "sample": "%d accounts x %d steps, OpenMP over %d host threads; oracle/qq_oracle.c restates "
                                   "curve25519-dalek 3.2.1's algorithms (dalek itself is not buildable here: no "
                                   "cargo/rustc, crate not vendored)" % (m, args.steps, cores)},
        "e2e": {"value": value, "unit": "accounts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if msm_ref:
        line["msm"] = msm_ref
    emit(line)


# =====================================================================================================================
# our arm
# =====================================================================================================================
def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    world, rank, local = dist_setup(args.gpus)
    if world > 1:
        # NCCL_DEBUG=VERSION (the image default) prints "NCCL version ..." on stdout; rank 0's stdout carries one JSON line
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # CPU-side barrier for the legs in which ONE process drives several GPUs: an NCCL barrier parks a spinning kernel of
        # another process on those GPUs, and two processes time-slice a GPU
        cpu_group = dist.new_group(backend="gloo")
    pkg = g.load_package()
    eng = pkg.Engine(local)
    dev = torch.device("cuda", local)
    n = args.accounts
    rng = np.random.default_rng(1000 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def maxr_early(x):
        if world == 1:
            return x
        t_ = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t_, op=dist.ReduceOp.MAX)
        return float(t_.item())

    # ---- synthetic inputs: valid accounts with known discrete logs, built with the fixed-base kernel ------------
    cols = [eng.fixed_base(0, rand_scalars(rng, n))[0] for _ in range(4)]
    acc_h = np.concatenate(cols, axis=1).copy()
    bl_h, u_h, c_h = rand_scalars(rng, n), rand_scalars(rng, n), rand_scalars(rng, n)
    # pinned host buffers for the end-to-end leg
    pin = {k: torch.from_numpy(v).pin_memory() for k, v in (("acc", acc_h), ("bl", bl_h), ("u", u_h), ("c", c_h))}
    out_pin = torch.empty(n * 128, dtype=torch.uint8).pin_memory()
    st_pin = torch.empty(n, dtype=torch.uint8).pin_memory()
    # device-resident copies for the kernel-only leg (torch owns the memory; the library gets raw pointers)
    d = {k: v.to(dev) for k, v in pin.items()}
    out_d = torch.empty(n * 128, dtype=torch.uint8, device=dev)
    st_d = torch.empty(n, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    import ctypes
    vp = ctypes.c_void_p

    def step_dev():
        eng.call_dev("qq_update_account_batch_dev", vp(d["acc"].data_ptr()), vp(d["bl"].data_ptr()),
                     vp(d["u"].data_ptr()), vp(d["c"].data_ptr()), vp(out_d.data_ptr()), vp(st_d.data_ptr()),
                     ctypes.c_size_t(n))

    def step_e2e():
        rc = eng.lib.qq_update_account_batch(eng.h, vp(pin["acc"].data_ptr()), vp(pin["bl"].data_ptr()),
                                             vp(pin["u"].data_ptr()), vp(pin["c"].data_ptr()), vp(out_pin.data_ptr()),
                                             vp(st_pin.data_ptr()), ctypes.c_size_t(n))
        if rc != 0:
            raise RuntimeError("qq_update_account_batch failed: %d" % rc)

    peak = eng.measure_imad_peak()

    # ---- kernel-only leg (inputs resident in HBM) -----------------------------------------------------------------
    for _ in range(args.warmup):
        step_dev()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    launches0 = eng.launch_count
    eng.event_record(0)
    t0 = time.time()
    t_region0 = t0
    vb_ms = 0.0
    breakdown = {}
    for _ in range(args.steps):
        step_dev()
        bd = eng.last_kernel_breakdown()
        vb_ms += bd["varbase"]
        for k_, v_ in bd.items():
            breakdown[k_] = breakdown.get(k_, 0.0) + v_
    eng.event_record(1)
    dev_ms = eng.event_elapsed_ms(0, 1)
    barrier()
    t_region1 = time.time()
    wall_ms = (t_region1 - t0) * 1e3
    launches = eng.launch_count - launches0
    assert int(st_d.max().item()) == 0

    # ---- end-to-end leg: host buffers in, host buffers out, through the public C ABI ----------------------------
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.time()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_ms = (time.time() - t0) * 1e3
    assert int(st_pin.max().item()) == 0
    same = bool(torch.equal(out_pin, out_d.cpu()))
    sampler.stop()
    # SHA-256 of this rank's 128 n output bytes.  Rank r always works on block r of the global batch (inputs seeded with
    # 1000 + r whatever the world size), so block r's digest is the same in every run that has a rank r: the 1 / 2 / 4 / 8-GPU
    # outputs are compared block by block, and block 0 against the CPU port below.
    import hashlib
    my_hash = hashlib.sha256(out_pin.numpy().tobytes()).digest()
    if world > 1:
        ht = torch.from_numpy(np.frombuffer(my_hash, np.uint8).copy()).to(dev)
        hall = [torch.zeros(32, dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(hall, ht)
        block_hashes = [bytes(h.cpu().numpy().tobytes()).hex() for h in hall]
    else:
        block_hashes = [my_hash.hex()]

    # ---- config 2, distribution (B): protocol-like balances, bl in {0 (7 of 9), +v, l - v} with v a uniform u64
    # (transaction.rs:68-72, accounts.rs:419-429: a transfer is [-v, +v, 0 x 7]); u, c stay uniform.  The balance only feeds
    # the fixed-base term (1.7 % of the step), so the rate is that of distribution (A).
    proto = None
    if n >= 9:
        v64 = rng.integers(0, 2**63, size=n, dtype=np.int64).astype(object)
        kind = np.arange(n) % 9
        blB = np.zeros((n, 32), np.uint8)
        idx_p, idx_m = np.nonzero(kind == 1)[0], np.nonzero(kind == 0)[0]
        for i in idx_p[:4096]:
            blB[i] = np.frombuffer(int(v64[i]).to_bytes(32, "little"), np.uint8)
        for i in idx_m[:4096]:
            blB[i] = np.frombuffer((L - int(v64[i])).to_bytes(32, "little"), np.uint8)
        # beyond the first 4096 of each kind the same values repeat (python big-int conversion is slow; the kernels do not care)
        if idx_p.size > 4096:
            blB[idx_p[4096:]] = blB[idx_p[:4096]][np.arange(idx_p.size - 4096) % 4096]
            blB[idx_m[4096:]] = blB[idx_m[:4096]][np.arange(idx_m.size - 4096) % 4096]
        blB_d = torch.from_numpy(blB.reshape(-1)).to(dev)
        outB = torch.empty(n * 128, dtype=torch.uint8, device=dev)

        def step_B():
            eng.call_dev("qq_update_account_batch_dev", vp(d["acc"].data_ptr()), vp(blB_d.data_ptr()), vp(d["u"].data_ptr()),
                         vp(d["c"].data_ptr()), vp(outB.data_ptr()), vp(st_d.data_ptr()), ctypes.c_size_t(n))
        step_B()
        barrier()
        eng.event_record(6)
        for _ in range(max(2, args.steps // 2)):
            step_B()
        eng.event_record(7)
        msB = eng.event_elapsed_ms(6, 7) / max(2, args.steps // 2)
        assert int(st_d.max().item()) == 0
        # where bl == 0 the commitment's d differs from (A) only through bl: spot-check pk' unchanged between the two runs
        pk_same = bool(torch.equal(outB.view(n, 128)[:, :64], out_d.view(n, 128)[:, :64]))
        proto = {"distribution": "bl in {0 (7/9), +v, l - v}, v uniform u64; u, c uniform 252-bit", "ms_per_step": msB,
                 "accounts_per_sec_per_gpu": n / (msB * 1e-3), "pk_half_equals_distribution_A": pk_same}
        del blB_d, outB

    # ---- the other three functions the north star names, at the same batch size (inputs resident in HBM; 32-byte encodings
    # out): RistrettoPublicKey::update_public_key, ElGamalCommitment::generate_commitment (fixed base + two variable-base
    # multiplications per commitment) and Account::create_delta_and_epsilon_accounts (host API, r supplied by the caller)
    named = None
    if n >= 1024 and not args.no_named_functions:
        pk_d = d["acc"].view(n, 128)[:, :64].contiguous()
        out64 = torch.empty(n * 64, dtype=torch.uint8, device=dev)
        reps_n = max(2, args.steps // 2)

        def timed(fn):
            fn()
            barrier()
            eng.event_record(6)
            for _ in range(reps_n):
                fn()
            eng.event_record(7)
            return eng.event_elapsed_ms(6, 7) / reps_n
        ms_pk = timed(lambda: eng.call_dev("qq_update_public_key_batch_dev", vp(pk_d.data_ptr()), vp(d["u"].data_ptr()),
                                           vp(out64.data_ptr()), vp(st_d.data_ptr()), ctypes.c_size_t(n)))
        assert int(st_d.max().item()) == 0
        pk_same = bool(torch.equal(out64.view(n, 64), out_d.view(n, 128)[:, :64]))      # update_account's pk half is update_public_key(pk, u)
        ms_gc = timed(lambda: eng.call_dev("qq_generate_commitment_batch_dev", vp(pk_d.data_ptr()), vp(d["c"].data_ptr()),
                                           vp(d["bl"].data_ptr()), vp(out64.data_ptr()), vp(st_d.data_ptr()), ctypes.c_size_t(n)))
        assert int(st_d.max().item()) == 0
        nde = min(n, 1 << 18)
        t_de = time.time()
        # BASE_PK_BTC_COMPRESSED (reference src/ristretto/constants.rs:12-21)
        base_pk = np.frombuffer(bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76"
                                              "8c9240b456a9e6dc65c377a1048d745f94a08cdb7f44cbcd7b46f34048871134"), np.uint8)
        dl, ep, st_de = eng.delta_epsilon(acc_h[:nde], bl_h[:nde], c_h[:nde], base_pk)
        ms_de = (time.time() - t_de) * 1e3
        assert not st_de.any()
        named = {"update_public_key": {"n": n, "ms": ms_pk, "per_sec_per_gpu": n / (ms_pk * 1e-3), "equals_pk_half_of_update_account": pk_same},
                 "generate_commitment": {"n": n, "ms": ms_gc, "per_sec_per_gpu": n / (ms_gc * 1e-3)},
                 "create_delta_and_epsilon_accounts": {"n": nde, "ms_host_api": ms_de, "per_sec_per_gpu": nde / (ms_de * 1e-3),
                                                       "note": "host buffers in and out (pageable), one call"}}
        del pk_d, out64

    # ---- configs[0]: the 9-account anonymity set (3x3 shuffle), latency of one update_account + verify_account call pair
    # through the host API (the reference runs this case on one CPU core; reported for information)
    anon9 = None
    if rank == 0:
        a9, b9, u9, c9 = acc_h[:9].copy(), bl_h[:9].copy(), u_h[:9].copy(), c_h[:9].copy()
        lat = []
        for rep in range(12):
            t0_ = time.time()
            o9, s9 = eng.update_account(a9, b9, u9, c9)
            eng.verify_account(o9, u9, b9)      # verdict not used: times the call
            lat.append((time.time() - t0_) * 1e3)
        lat.sort()
        anon9 = {"accounts": 9, "update_plus_verify_ms_median": lat[len(lat) // 2], "ms_min": lat[0],
                 "matches_batch_output": bool((o9.reshape(-1) == out_d.cpu().numpy()[:9 * 128]).all())}

    # ---- fixed-base throughput (north-star target: 1e9 / s): scalars resident, 32-byte encodings out ----------------
    fixed = None
    if args.fixed_points > 0:
        nf = args.fixed_points
        fs = torch.from_numpy(rand_scalars(rng, nf).reshape(-1)).to(dev)
        fo = torch.empty(nf * 32, dtype=torch.uint8, device=dev)
        fst = torch.empty(nf, dtype=torch.uint8, device=dev)
        fixed = {"n": nf, "scalars": "uniform 252-bit", "output": "32-byte compressed points", "windows": []}
        w_default = eng.fixed_base_window(0)
        ref_sum = None
        for W in ([w_default] + ([args.fixed_window] if args.fixed_window not in (0, w_default) else [])):
            eng.fixed_base_set_window(0, W)
            best = 1e30
            for rep in range(4):
                eng.call_dev("qq_fixed_base_batch_dev", ctypes.c_int(0), vp(fs.data_ptr()), vp(fo.data_ptr()),
                             vp(fst.data_ptr()), ctypes.c_size_t(nf))
                if rep:
                    best = min(best, eng.last_kernel_ms)
            chk = int(fo.to(torch.int64).sum().item())
            if ref_sum is None:
                ref_sum = chk
            ent = (1 << (W - 1)) + 1
            # executed work per scalar: (NW - 1) mixed additions of 7 field products + the lift of the first entry (1) + the
            # double-and-compress encoder (4 squarings + 19 products + 3 for its share of the batch inversion); a product is
            # 72 32x32->64 multiplies (64 + 8 for the fold), a squaring 44.  Table traffic: NW gathers of 96 bytes.
            nw_ = (255 + W - 1) // W
            wide_per_scalar = ((nw_ - 1) * 7 + 1 + 19 + 3) * 72 + 4 * 44
            hbm_peak = (json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0)
            fixed["windows"].append({"window_bits": W, "table_bytes": nw_ * ent * 96, "ms": best,
                                     "mults_per_sec_per_gpu": nf / (best * 1e-3),
                                     "executed_wide_multiplies_per_scalar": wide_per_scalar,
                                     "wide_multiply_frac": nf * wide_per_scalar / (best * 1e-3) / peak["imad_wide_per_s"],
                                     "table_gather_GBps": nf * nw_ * 96 / (best * 1e-3) / 1e9,
                                     "table_gather_frac_of_hbm_peak": nf * nw_ * 96 / (best * 1e-3) / 1e9 / hbm_peak,
                                     "table_resident_in": "L2 (126 MB)" if nw_ * ent * 96 <= 100e6 else "HBM",
                                     "same_output_as_default_window": chk == ref_sum})
        # signed 64-bit values (balances, the `bl` of update_account / generate_commitment): only the low windows are walked
        fv = torch.from_numpy(rng.integers(-2**63, 2**63 - 1, size=nf, dtype=np.int64)).to(dev)
        best = 1e30
        for rep in range(4):
            eng.call_dev("qq_fixed_base_i64_batch_dev", ctypes.c_int(0), vp(fv.data_ptr()), vp(fo.data_ptr()), ctypes.c_size_t(nf))
            if rep:
                best = min(best, eng.last_kernel_ms)
        fixed["i64_values"] = {"window_bits": eng.fixed_base_window(0), "ms": best, "mults_per_sec_per_gpu": nf / (best * 1e-3),
                               "scalars": "uniform signed 64-bit values (Scalar::from(u64) balances and their negations)"}
        eng.fixed_base_set_window(0, w_default)
        del fs, fo, fst, fv

    # ---- MSM 2^20 (configs[3]): known-dlog points, last scalar solved so the sum is the identity ------------------
    msm = None
    if args.msm_points > 0:
        m = args.msm_points
        hs = rand_scalars(rng, m)
        pts_h, _ = eng.fixed_base(0, hs)
        a = rand_scalars(rng, m)
        pts_d = torch.from_numpy(pts_h.reshape(-1)).to(dev)
        a_d = torch.from_numpy(a.reshape(-1)).to(dev)
        small = torch.zeros(256, dtype=torch.uint8, device=dev)
        for _ in range(2):
            eng.call_dev("qq_msm_dev", vp(a_d.data_ptr()), vp(pts_d.data_ptr()), ctypes.c_size_t(m),
                         vp(small.data_ptr()), vp(small.data_ptr() + 64))
        barrier()
        eng.event_record(2)
        reps = max(2, args.steps)
        bdm = {}
        for _ in range(reps):
            eng.call_dev("qq_msm_dev", vp(a_d.data_ptr()), vp(pts_d.data_ptr()), ctypes.c_size_t(m),
                         vp(small.data_ptr()), vp(small.data_ptr() + 64))
            for k_, v_ in eng.last_kernel_breakdown().items():
                bdm[k_] = bdm.get(k_, 0.0) + v_ / reps
        eng.event_record(3)
        msm_ms = eng.event_elapsed_ms(2, 3) / reps
        barrier()
        res = small.cpu().numpy()
        msm_out32 = res[:32].copy()
        # correctness at full size: sum a_i h_i * B computed on the host with big ints, one fixed-base mult on the GPU
        tot = 0
        ai = [int.from_bytes(a[i].tobytes(), "little") for i in range(0, m)] if m <= (1 << 20) else None
        if ai is not None:
            hi_ = [int.from_bytes(hs[i].tobytes(), "little") for i in range(0, m)]
            tot = sum(x * y for x, y in zip(ai, hi_)) % L
            exp, _ = eng.fixed_base(0, np.frombuffer(tot.to_bytes(32, "little"), np.uint8))
            msm_ok = bool((exp[0] == res[:32]).all()) and int(res[64]) == 0
        else:
            msm_ok = None
        msm = {"points": m, "ms": msm_ms, "points_per_sec_per_gpu": m / (msm_ms * 1e-3),
               "breakdown_ms": bdm, "matches_known_dlog": msm_ok,
               "imad_frac": m * IMAD_PER_MSM_POINT / (msm_ms * 1e-3) / peak["imad_lo_per_s"]}
        # ---- the same MSM over a point set decompressed once (Bulletproofs generators are fixed and cacheable) --------
        hprep = ctypes.c_void_p()
        eng._ck(eng.lib.qq_msm_points_prepare_dev(eng.h, vp(pts_d.data_ptr()), ctypes.c_size_t(m), ctypes.byref(hprep)),
                "qq_msm_points_prepare_dev")
        for _ in range(2):
            eng.call_dev("qq_msm_prepared_dev", vp(a_d.data_ptr()), hprep, ctypes.c_size_t(m), vp(small.data_ptr() + 128),
                         vp(small.data_ptr() + 192))
        eng.event_record(4)
        for _ in range(reps):
            eng.call_dev("qq_msm_prepared_dev", vp(a_d.data_ptr()), hprep, ctypes.c_size_t(m), vp(small.data_ptr() + 128),
                         vp(small.data_ptr() + 192))
        eng.event_record(5)
        prep_ms = eng.event_elapsed_ms(4, 5) / reps
        res2 = small.cpu().numpy()
        eng.lib.qq_msm_points_free(eng.h, hprep)
        msm["prepared_points"] = {"ms": prep_ms, "points_per_sec_per_gpu": m / (prep_ms * 1e-3),
                                  "same_result": bool((res2[128:160] == res[:32]).all()) and int(res2[192]) == 0,
                                  "imad_frac": m * 16_128 / (prep_ms * 1e-3) / peak["imad_lo_per_s"],
                                  "note": "points decompressed once with qq_msm_points_prepare (not timed), 16 128 IMAD units per point"}
        # ---- ONE point set split over the ranks (strong scaling; BASELINE configs[4]: sweep 2^10 .. 2^24 at 1 / 2 / 4 / 8 GPUs).
        # Every rank derives the same global scalars (seed 77) and builds only its contiguous slice of the points; Pippenger
        # on the slice (qq_msm_partial_dev), the 144-byte partial records all-gathered over NCCL into DEVICE memory and added
        # and encoded there by one kernel (qq_points_sum_dev) - no host bounce of the partials.  Timed on the host clock around
        # the device-synchronous calls, max over ranks; world = 1 is the plain single-GPU MSM of the same set.
        sweep = []
        grng = np.random.default_rng(77)
        rec = torch.zeros(144, dtype=torch.uint8, device=dev)
        recs = torch.zeros(144 * world, dtype=torch.uint8, device=dev)
        sizes = [s_ for s_ in (1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24) if s_ <= args.msm_sweep_max]
        for tot_n in sizes:
            # global scalars in blocks of 2^16 so that every rank draws the same stream without holding 2^24 x 64 bytes twice
            lo, hi = tot_n * rank // world, tot_n * (rank + 1) // world
            hs_l = np.empty((hi - lo, 32), np.uint8)
            as_l = np.empty((hi - lo, 32), np.uint8)
            acc_dot = 0
            for b0 in range(0, tot_n, 1 << 16):
                b1 = min(tot_n, b0 + (1 << 16))
                hb, ab = rand_scalars(grng, b1 - b0), rand_scalars(grng, b1 - b0)
                if tot_n <= (1 << 14):      # expected result from the known discrete logs (big ints on the host; small sets only)
                    acc_dot += sum(int.from_bytes(hb[k].tobytes(), "little") * int.from_bytes(ab[k].tobytes(), "little") for k in range(b1 - b0))
                a_, b_ = max(b0, lo), min(b1, hi)
                if a_ < b_:
                    hs_l[a_ - lo:b_ - lo] = hb[a_ - b0:b_ - b0]
                    as_l[a_ - lo:b_ - lo] = ab[a_ - b0:b_ - b0]
            cnt = hi - lo
            if cnt:
                hs_t = torch.from_numpy(hs_l.reshape(-1)).to(dev)
                pt_t = torch.empty(cnt * 32, dtype=torch.uint8, device=dev)
                fst_t = torch.empty(cnt, dtype=torch.uint8, device=dev)
                eng.call_dev("qq_fixed_base_batch_dev", ctypes.c_int(0), vp(hs_t.data_ptr()), vp(pt_t.data_ptr()), vp(fst_t.data_ptr()),
                             ctypes.c_size_t(cnt))
                as_t = torch.from_numpy(as_l.reshape(-1)).to(dev)
                del hs_t, fst_t
            else:
                pt_t = torch.zeros(32, dtype=torch.uint8, device=dev)
                as_t = torch.zeros(32, dtype=torch.uint8, device=dev)

            def strong_once():
                eng.call_dev("qq_msm_partial_dev", vp(as_t.data_ptr()), vp(pt_t.data_ptr()), ctypes.c_size_t(cnt), vp(rec.data_ptr()),
                             vp(rec.data_ptr() + 128))
                if world > 1:
                    dist.all_gather_into_tensor(recs, rec)
                    return eng.points_sum_dev(recs.data_ptr(), world, 144)
                return eng.points_sum_dev(rec.data_ptr(), 1, 144)
            strong_once()
            strong_once()
            barrier()
            t0s = time.time()
            nrep = 3 if tot_n >= (1 << 22) else 6
            for _ in range(nrep):
                out_s, ident_s, st_s = strong_once()
            torch.cuda.synchronize(dev)
            ms_s = maxr_early((time.time() - t0s) * 1e3 / nrep)
            # exchange alone: gather + device-side sum of records that are already there
            barrier()
            t0s = time.time()
            for _ in range(10):
                if world > 1:
                    dist.all_gather_into_tensor(recs, rec)
                    eng.points_sum_dev(recs.data_ptr(), world, 144)
                else:
                    eng.points_sum_dev(rec.data_ptr(), 1, 144)
            torch.cuda.synchronize(dev)
            ex_ms = maxr_early((time.time() - t0s) * 1e3 / 10)
            t0s = time.time()
            for _ in range(10):
                eng.points_sum_dev(rec.data_ptr(), 1, 144)
            enc_ms = maxr_early((time.time() - t0s) * 1e3 / 10)
            ent = {"points_total": tot_n, "points_per_gpu": cnt, "ms": ms_s, "points_per_sec": tot_n / (ms_s * 1e-3),
                   "exchange_ms": ex_ms, "encode_only_ms": enc_ms, "exchange_over_single_gpu_ms": max(0.0, ex_ms - enc_ms),
                   "status": int(st_s),
                   "imad_frac_per_gpu": cnt * IMAD_PER_MSM_POINT / (ms_s * 1e-3) / peak["imad_lo_per_s"]}
            if tot_n <= (1 << 14):
                exp_s, _ = eng.fixed_base(0, np.frombuffer((acc_dot % L).to_bytes(32, "little"), np.uint8))
                ent["matches_known_dlog"] = bool((exp_s[0] == out_s).all())
            ent["result_sha16"] = out_s.tobytes().hex()[:16]      # equal across world sizes: the same global point set
            sweep.append(ent)
            del pt_t, as_t
        msm["strong"] = {"sweep": sweep,
                         "exchange": "NCCL all_gather_into_tensor of 144-byte partial records into device memory + one kernel that adds "
                                     "and encodes them (qq_points_sum_dev); no host bounce" if world > 1 else "single GPU",
                         "note": "one global point set (seed 77) split into contiguous slices, one per rank; ms = max over ranks"}

    # ---- configs[2]: 4096 shuffle proofs, sharded over the ranks (independent proofs, no collective) ----------------
    # ---- configs[3]: Bulletproofs 64-bit range proofs, 16 aggregated values each: ONE aggregated MSM per batch -----------
    # Workload = the committed valid proofs of tests/golden (made by the oracle's prover restatements), tiled.  Host buffers
    # in, verdicts out, through the public C ABI (host Merlin transcripts + GPU batches inside the timed region).
    proofs_sec = None
    if args.proofs > 0:
        gold = os.path.join(ROOT, "tests", "golden")
        per_rank = max(1, args.proofs // world)

        def tiled(path, rec_bytes, count):
            raw = np.fromfile(path, dtype=np.uint8).reshape(-1, rec_bytes)
            return np.tile(raw, ((count + raw.shape[0] - 1) // raw.shape[0], 1))[:count]

        def best_of(fn, reps=3):
            fn()
            barrier()
            best = 1e30
            for _ in range(reps):
                t_ = time.time()
                fn()
                best = min(best, (time.time() - t_) * 1e3)
            barrier()
            return best
        def pinned(a):      # page-locked host buffers, as a caller feeding serialised transactions would hold them
            return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        rec_all = tiled(os.path.join(gold, "shuffle_proofs.bin"), 6432, per_rank * world)
        cols_ = ((0, 1152), (1152, 2304), (2304, 2656), (2656, 6432))
        rec = rec_all[rank * per_rank:(rank + 1) * per_rank]
        si, so, stm, prf = (pinned(rec[:, a:b]) for a, b in cols_)
        res = {}

        def run_shuffle():
            res["st"] = eng.verify_shuffle(si, so, stm, prf)[0]
        sh_ms = best_of(run_shuffle)
        sh_ok = not res["st"].any()
        # the same batch through TWO contexts on this GPU from two caller threads ("one qq_ctx per caller thread"): while one
        # thread waits for its GPU batch the other runs its transcripts on the host cores
        import threading
        eng2 = pkg.Engine(local)
        half = per_rank // 2
        res2 = {}

        def run_shuffle_two():
            def part(e, lo, hi, key):
                res2[key] = e.verify_shuffle(si[lo:hi], so[lo:hi], stm[lo:hi], prf[lo:hi])[0]
            th = [threading.Thread(target=part, args=(eng, 0, half, "a")), threading.Thread(target=part, args=(eng2, half, per_rank, "b"))]
            for x in th:
                x.start()
            for x in th:
                x.join()
        sh2_ms = best_of(run_shuffle_two) if half >= 1 else None
        sh2_ok = half >= 1 and not res2["a"].any() and not res2["b"].any()
        eng2.close()
        prf_bad = prf.copy()
        prf_bad[per_rank // 2, 3776 - 1] ^= 0x01            # top byte of the DDH response: wrong, perhaps not even canonical
        sh_rej = eng.verify_shuffle(si, so, stm, prf_bad)[0]
        # one proof with a tampered output account (only the group equations see it): the aggregate fails, grouped MSMs locate the
        # failing group of 64 proofs, the exact form names the proof
        so_bad = pinned(so)
        so_bad[per_rank // 2, 5] ^= 1

        def run_shuffle_bad():
            res["st_bad"] = eng.verify_shuffle(si, so_bad, stm, prf)[0]
        sh_bad_ms = best_of(run_shuffle_bad, reps=2)
        sh_bad_ok = np.nonzero(res["st_bad"])[0].tolist() == [per_rank // 2]
        # strong scaling inside this run: rank 0 alone verifies the WHOLE batch (the other ranks wait), against the sharded time
        sh_n1 = None
        if world > 1:
            barrier()
            if rank == 0:
                full = [pinned(rec_all[:, a:b]) for a, b in cols_]
                eng.verify_shuffle(*full)
                t_ = time.time()
                for _ in range(3):
                    st_full = eng.verify_shuffle(*full)[0]
                sh_n1 = (time.time() - t_) * 1e3 / 3
                assert not st_full.any()
                del full
            barrier()
        # the exact form (every group equation an MSM of its own) on the same batch, for the record
        eng.verify_set_aggregation(False)
        sh_exact_ms = best_of(run_shuffle, reps=2)
        eng.verify_set_aggregation(True)
        proofs_sec = {"shuffle": {"proofs_per_gpu": per_rank, "ms": sh_ms, "all_accepted": bool(sh_ok),
                                  "ms_two_contexts": sh2_ms, "all_accepted_two_contexts": bool(sh2_ok),
                                  "ms_exact_form": sh_exact_ms, "n1_ms_whole_batch_on_rank0": sh_n1,
                                  "ms_one_tampered_proof": sh_bad_ms, "one_tampered_proof_isolated": bool(sh_bad_ok),
                                  "tampered_proof_rejected_alone": bool(sh_rej[per_rank // 2] != 0 and int(sh_rej.astype(bool).sum()) == 1),
                                  "msms_per_proof": 32, "terms_per_proof": 239, "exact_msms_per_proof": 4, "aggregated_terms_per_proof": 166,
                                  "api": "qq_verify_shuffle_batch (ShuffleProof::verify, 9 accounts): proof bytes uploaded, Merlin transcripts "
                                         "and Z/l algebra in transcript kernels, G / H / g_r / h_r as exact MSMs, the other 28 group equations "
                                         "of all proofs in one weighted Pippenger MSM; pinned host buffers in, verdict bytes out"}}
        m_rp = 16
        rp_rec = m_rp * 32 + eng.range_proof_bytes(m_rp)
        rp_list = []
        for count in sorted(set([max(1, 512 // world), per_rank])):
            rr = tiled(os.path.join(gold, "range_proofs_m%d.bin" % m_rp), rp_rec, count)
            cm, rpf = pinned(rr[:, :m_rp * 32]), pinned(rr[:, m_rp * 32:])

            def run_range():
                res["st"] = eng.verify_range_proofs(cm, rpf, m_rp)
            rp_ms = best_of(run_range)
            ok_ = not res["st"].any()
            rej, rp_bad_ms = None, None
            if count >= 3:
                rpf_bad = pinned(rpf)
                rpf_bad[count // 2, 5 * 32 + 1] ^= 1

                def run_range_bad():
                    res["st_bad"] = eng.verify_range_proofs(cm, rpf_bad, m_rp)
                rp_bad_ms = best_of(run_range_bad, reps=2)
                r_ = res["st_bad"]
                rej = bool(r_[count // 2] == 6 and int(r_.astype(bool).sum()) == 1)
            rp_n1 = None
            if world > 1 and count == per_rank:
                barrier()
                if rank == 0:
                    rr_all = tiled(os.path.join(gold, "range_proofs_m%d.bin" % m_rp), rp_rec, count * world)
                    cm_a, rp_a = pinned(rr_all[:, :m_rp * 32]), pinned(rr_all[:, m_rp * 32:])
                    eng.verify_range_proofs(cm_a, rp_a, m_rp)
                    t_ = time.time()
                    for _ in range(3):
                        st_full = eng.verify_range_proofs(cm_a, rp_a, m_rp)
                    rp_n1 = (time.time() - t_) * 1e3 / 3
                    assert not st_full.any()
                    del rr_all, cm_a, rp_a
                barrier()
            rp_list.append({"proofs_per_gpu": count, "values_per_proof": m_rp, "bits": 64, "ms": rp_ms, "all_accepted": bool(ok_),
                            "n1_ms_whole_batch_on_rank0": rp_n1,
                            "tampered_proof_rejected_alone": rej, "ms_one_tampered_proof": rp_bad_ms,
                            "aggregated_msm_terms": 2 * 64 * m_rp + 2 + count * (4 + 20 + m_rp),
                            "reference_msm_points": count * (2 * 64 * m_rp + 20 + m_rp + 6)})
        # sigma verifier (Verifier::verify_update_account_verifier): per_rank DLOG proofs over 9 accounts each, random valid points
        # (every proof is rejected by its challenge; the work is that of valid proofs)
        srng = np.random.default_rng(4 + rank)
        sitems = per_rank * 9
        s_ia = np.concatenate([eng.fixed_base(0, rand_scalars(srng, sitems))[0] for _ in range(4)], axis=1).copy()
        s_da = np.concatenate([eng.fixed_base(0, rand_scalars(srng, sitems))[0] for _ in range(4)], axis=1).copy()
        s_ia, s_da, s_z, s_x = pinned(s_ia), pinned(s_da), pinned(rand_scalars(srng, sitems)), pinned(rand_scalars(srng, per_rank))

        def run_sigma():
            res["sg"] = eng.verify_update_account_dlog(s_ia, s_da, s_z, s_x, 9)
        sg_ms = best_of(run_sigma)
        proofs_sec["sigma_dlog"] = {"proofs_per_gpu": per_rank, "accounts_per_proof": 9, "ms": sg_ms, "msms": 2 * sitems,
                                    "all_rejected_by_challenge": bool((res["sg"] == 6).all()),
                                    "api": "qq_verify_update_account_dlog_batch (Verifier::verify_update_account_verifier): accounts, responses "
                                           "and challenges uploaded, 3-term MSMs + Merlin transcripts on the device, one status byte per proof back"}
        proofs_sec["range_proofs"] = {"batches": rp_list,
                                      "api": "qq_verify_range_proof_batch (RangeProof::verify_multiple; the reference verifies "
                                             "each proof with its own 2090-term MSM, verifier.rs:517)"}

    # ---- the in-library multi-device handle (qq_init_multi): ONE process drives every GPU of this run; the other ranks wait.
    # One 2^20-point MSM split over the devices (slices resident), partial sums pulled to the root by peer copies.
    multi_leg = None
    if world > 1 and args.msm_points > 0:
        barrier()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                me = pkg.MultiEngine(list(range(world)))
                tot_n = args.msm_points
                mrng = np.random.default_rng(78)
                hs_g, as_g = rand_scalars(mrng, tot_n), rand_scalars(mrng, tot_n)
                pts_g, _ = eng.fixed_base(0, hs_g)
                sp, pp, cn = [], [], []
                for d_ in range(world):
                    lo, hi = tot_n * d_ // world, tot_n * (d_ + 1) // world
                    c_ = me.ctx(d_)
                    ptrs = []
                    for arr in (as_g[lo:hi], pts_g[lo:hi]):
                        p_ = ctypes.c_void_p()
                        assert eng.lib.qq_dev_alloc(c_, ctypes.byref(p_), ctypes.c_size_t(arr.size)) == 0
                        assert eng.lib.qq_dev_upload(c_, p_, vp(np.ascontiguousarray(arr).ctypes.data), ctypes.c_size_t(arr.size)) == 0
                        ptrs.append(p_)
                    sp.append(ptrs[0].value)
                    pp.append(ptrs[1].value)
                    cn.append(hi - lo)
                me.msm_dev(sp, pp, cn)
                t_ = time.time()
                for _ in range(5):
                    mo_, ms__ = me.msm_dev(sp, pp, cn)
                mm_ms = (time.time() - t_) * 1e3 / 5
                one_o, one_s = eng.msm(as_g, pts_g)
                multi_leg = {"devices": world, "msm_points_total": tot_n, "ms": mm_ms, "points_per_sec": tot_n / (mm_ms * 1e-3),
                             "equals_single_gpu_result": bool(ms__ == 0 and one_s == 0 and (mo_ == one_o).all()),
                             "api": "qq_multi_msm_dev: one worker thread per GPU, partial sums gathered on the root with "
                                    "cudaMemcpyPeerAsync and added by one kernel"}
                for d_ in range(world):
                    eng.lib.qq_dev_free(me.ctx(d_), ctypes.c_void_p(sp[d_]))
                    eng.lib.qq_dev_free(me.ctx(d_), ctypes.c_void_p(pp[d_]))
                me.close()
            except Exception as ex:      # reported, not fatal: the headline legs are done
                multi_leg = {"error": repr(ex)}
        dist.barrier(group=cpu_group)      # the other ranks wait on the CPU: their GPUs are rank 0's for this leg
        barrier()

    # ---- reduce over ranks ----------------------------------------------------------------------------------------
    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dev_ms_max, wall_ms_max, e2e_ms_max = maxr(dev_ms), maxr(wall_ms), maxr(e2e_ms)
    msm_ms_max = maxr(msm["ms"]) if msm else None
    if proofs_sec:
        sh = proofs_sec["shuffle"]
        sh["ms"] = maxr(sh["ms"])
        sh["proofs_total"] = sh["proofs_per_gpu"] * world
        sh["proofs_per_sec"] = sh["proofs_total"] / (sh["ms"] * 1e-3)
        if sh.get("n1_ms_whole_batch_on_rank0"):
            sh["strong_scaling"] = {"gpus": world, "proofs_total": sh["proofs_total"], "ms_1_gpu": sh["n1_ms_whole_batch_on_rank0"],
                                    "ms": sh["ms"], "speedup": sh["n1_ms_whole_batch_on_rank0"] / sh["ms"],
                                    "efficiency": sh["n1_ms_whole_batch_on_rank0"] / sh["ms"] / world}
        if sh.get("ms_two_contexts"):
            sh["ms_two_contexts"] = maxr(sh["ms_two_contexts"])
            sh["proofs_per_sec_two_contexts"] = sh["proofs_total"] / (sh["ms_two_contexts"] * 1e-3)
        for b_ in proofs_sec["range_proofs"]["batches"]:
            b_["ms"] = maxr(b_["ms"])
            b_["proofs_per_sec"] = b_["proofs_per_gpu"] * world / (b_["ms"] * 1e-3)
            b_["values_per_sec"] = b_["proofs_per_sec"] * b_["values_per_proof"]
            b_["reference_msm_points_per_sec"] = b_["reference_msm_points"] * world / (b_["ms"] * 1e-3)
            if b_.get("n1_ms_whole_batch_on_rank0"):
                b_["strong_scaling"] = {"gpus": world, "proofs_total": b_["proofs_per_gpu"] * world, "ms_1_gpu": b_["n1_ms_whole_batch_on_rank0"],
                                        "ms": b_["ms"], "speedup": b_["n1_ms_whole_batch_on_rank0"] / b_["ms"],
                                        "efficiency": b_["n1_ms_whole_batch_on_rank0"] / b_["ms"] / world}
    total_accounts = n * world * args.steps
    value = total_accounts / (wall_ms_max * 1e-3)
    e2e_value = total_accounts / (e2e_ms_max * 1e-3)

    # ---- CPU baseline on rank 0 (bounded sample) ------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import c_oracle as C
        C.set_threads(len(os.sched_getaffinity(0)))
        cores = C.threads()
        m0 = 64 * cores
        t = time.time()
        eo, es = C.update_account(acc_h[:m0], bl_h[:m0], u_h[:m0], c_h[:m0])
        r0 = m0 / (time.time() - t)
        ms_ = int(max(m0, min(n, r0 * args.cpu_seconds)))
        t = time.time()
        eo, es = C.update_account(acc_h[:ms_], bl_h[:ms_], u_h[:ms_], c_h[:ms_])
        dtc = time.time() - t
        same_cpu = bool((eo.reshape(-1) == out_pin.numpy()[:ms_ * 128]).all())
        cpu = {"value": ms_ / dtc, "unit": "accounts/s", "cores": cores, "kind": "port",
               "sample": "first %d accounts of the same batch, OpenMP over %d host threads; oracle/qq_oracle.c restates "
                         "curve25519-dalek 3.2.1's algorithms (radix-2^51 field, radix-16 variable-base, table "
                         "fixed-base); dalek itself cannot be built here" % (ms_, cores),
               "gpu_output_matches_cpu_on_sample": same_cpu}
        if msm:
            # the MSM half of the metric on the host cores: the same 2^20 compressed points and scalars through the C port
            t = time.time()
            cmo, cms = C.msm(a, pts_h)
            dtm = time.time() - t
            msm["cpu_baseline"] = {"value": msm["points"] / dtm, "unit": "points/s", "ms": dtm * 1e3, "cores": cores, "kind": "port",
                                   "sample": "the same %d-point MSM (oracle/qq_oracle.c oq_msm), %d host threads" % (msm["points"], cores),
                                   "gpu_output_matches_cpu": bool(cms == 0 and (cmo == msm_out32).all())}
        if proofs_sec:
            # one aggregated range proof (16 values) through the oracle's verifier restatement: Merlin in Python, the 2090-term
            # MSM in the C port with all host threads -- what the reference does per proof
            import rangeproof_ref as RP
            from merlin_ref import Transcript as OT
            rr = np.fromfile(os.path.join(ROOT, "tests", "golden", "range_proofs_m16.bin"), dtype=np.uint8).reshape(-1, 16 * 32 + 928)
            gens16 = RP.BulletproofGens(64, 16)
            lat = []
            for i in range(min(3, rr.shape[0])):
                b_ = rr[i].tobytes()
                tr_ = OT(b"SenderAccountProof")
                tr_.domain_sep(b"BulletProof")
                tr_.domain_sep(b"AggregateBulletProof")
                t = time.time()
                okp = RP.verify_multiple(tr_, b_[512:], [b_[32 * j:32 * j + 32] for j in range(16)], 64, bp_gens=gens16)
                lat.append((time.time() - t) * 1e3)
                assert okp
            proofs_sec["range_proofs"]["cpu_port_ms_per_proof"] = sorted(lat)[len(lat) // 2]
            proofs_sec["range_proofs"]["cpu_port_note"] = "oracle/rangeproof_ref.py verify_multiple, MSM in oracle/qq_oracle.c on %d host threads" % cores
        if anon9 is not None:
            # configs[0] as the reference runs it: the same 9 accounts on ONE host core (update_account + verify_account)
            C.set_threads(1)
            lat = []
            for rep in range(7):
                t = time.time()
                o9c, _ = C.update_account(acc_h[:9], bl_h[:9], u_h[:9], c_h[:9])
                C.verify_account(o9c, u_h[:9], bl_h[:9])
                lat.append((time.time() - t) * 1e3)
            lat.sort()
            anon9["cpu_port_one_core_ms_median"] = lat[len(lat) // 2]
            C.set_threads(cores)

    if rank == 0:
        clocks = sampler.summary(t_region0, t_region1)
        vb_avg_ms = vb_ms / args.steps
        vb_work = 4 * IMAD_PER_VARBASE * n                       # 4 variable-base mults per account in one launch
        achieved = vb_work / (vb_avg_ms * 1e-3)
        step_frac = (IMAD_PER_UPDATE_ACCOUNT * n) / (dev_ms / args.steps * 1e-3) / peak["imad_lo_per_s"]
        line = {
            "metric": "account_updates_per_sec", "value": value, "unit": "accounts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (8 saturated 32-bit limbs, 32x32->64 IMAD.WIDE carry chains)", "data": "synthetic",
            "config": {"workload": "Account::update_account over 2^20 accounts per GPU (BASELINE.json configs[1])"
                       if n == (1 << 20) else "Account::update_account over %d accounts per GPU" % n,
                       "accounts_per_gpu": n, "scalars": "uniform 252-bit (worst case)",
                       "l2": "inputs+outputs per step (%.0f MB) exceed the 126 MB L2; no flush needed" % (n * 352 / 1e6),
                       "sharding": "independent contiguous account slices per GPU, no data-path collective"},
            "device_ms_per_step": dev_ms_max / args.steps,
            "e2e": {"value": e2e_value, "unit": "accounts/s", "h2d_bytes_per_step": n * 224, "d2h_bytes_per_step": n * 129,
                    "ms_per_step": e2e_ms_max / args.steps, "api": "qq_update_account_batch (host pointers, pinned)",
                    "matches_device_path": same},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "imad", "kernel": "k_varbase_split (the 4 variable-base scalar mults of every account)",
                         "achieved": achieved / 1e12, "peak": peak["imad_lo_per_s"] / 1e12,
                         "unit": "Tera thread-IMAD/s (32x32->64 product = 2)", "frac": achieved / peak["imad_lo_per_s"],
                         "peak_source": "measured live by qq_measure_imad_peak (independent mad.lo.u32 chains, all SMs)",
                         "peak_theoretical": 148 * 64 * 1.965e9 / 1e12,
                         "numerator": "SURVEY App. B algorithmic figure, 4 x 289 000 IMAD units per account; the kernel "
                                      "reaches the same results with 75 % of those multiplies (312 instead of 504 "
                                      "doublings per point), see executed_frac",
                         "executed_frac": IMAD_EXECUTED_VARBASE_PER_ACCOUNT * n / (vb_avg_ms * 1e-3) / peak["imad_lo_per_s"],
                         "imad_wide_peak_per_s": peak["imad_wide_per_s"],
                         "whole_step_frac": step_frac, "kernel_ms_per_launch": vb_avg_ms,
                         "kernel_share_of_step": vb_avg_ms / (dev_ms / args.steps),
                         "breakdown_ms_per_step": {k_: v_ / args.steps for k_, v_ in breakdown.items()},
                         # dram__bytes_read.sum + dram__bytes_write.sum of one k_varbase_split launch at 2^20 accounts, from
                         # the ncu --set full capture summarised in profiles/ncu_varbase_split_r01_summary.json (per-thread
                         # window tables spilling past L2, 8.5 % of HBM bandwidth; the same capture shows the FMA-heavy
                         # integer-multiply pipe active 88.9 % of elapsed cycles)
                         "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH * (n / (1 << 20)),
                         "traffic_unit": "bytes per launch (ncu, 2^20-account launch, scaled linearly to this batch)",
                         "ncu_fmaheavy_pipe_active_pct": 88.9,
                         "hbm": {"algorithmic_bytes_per_step": n * BYTES_PER_UPDATE_ACCOUNT,
                                 "achieved_GBps": n * BYTES_PER_UPDATE_ACCOUNT / (dev_ms / args.steps * 1e-3) / 1e9,
                                 "peak_GBps": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
                                 if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0}},
            "cpu_baseline": cpu,
        }
        line["output_sha256_per_block"] = block_hashes
        line["config"]["blocks"] = "rank r processes block r of the global batch (inputs seeded 1000 + r): block digests are comparable across world sizes"
        if proto:
            proto["accounts_per_sec"] = proto["accounts_per_sec_per_gpu"] * world
            line["update_account_protocol_like"] = proto
        if named is not None:
            line["named_functions"] = named
        if multi_leg:
            line["multi_device_handle"] = multi_leg
        if anon9:
            line["anonymity_set_9"] = anon9
        if fixed:
            for w_ in fixed["windows"] + [fixed["i64_values"]]:
                w_["mults_per_sec"] = w_["mults_per_sec_per_gpu"] * world
            line["fixed_base"] = fixed
        if proofs_sec:
            line["proof_verification"] = proofs_sec
        if msm:
            msm["points_per_sec"] = msm["points"] * world / (msm_ms_max * 1e-3)
            msm["note"] = "each GPU runs a full %d-point MSM (weak scaling); compressed input, decompression included" % msm["points"]
            line["msm"] = msm
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    eng.close()


_JSON_FD = None


def claim_stdout():
    """Rank 0's stdout carries exactly one JSON line: file descriptor 1 is pointed at stderr for the whole run (NCCL prints its
    version banner on fd 1 from C, whatever NCCL_DEBUG says in this image) and the line is written to the saved descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--accounts", type=int, default=1 << 20, help="accounts per GPU per step")
    ap.add_argument("--msm-points", type=int, default=1 << 20)
    ap.add_argument("--msm-sweep-max", type=int, default=1 << 24, help="largest point set of the strong-scaling MSM sweep (0 = skip)")
    ap.add_argument("--fixed-points", type=int, default=1 << 22, help="fixed-base batch per GPU (0 = skip)")
    ap.add_argument("--fixed-window", type=int, default=22,
                    help="also time the fixed-base batch with this table window (22 bits = 2.4 GB in HBM; 0 = default table only)")
    ap.add_argument("--proofs", type=int, default=4096, help="shuffle / range proofs verified in total over the ranks (0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds-per-step", type=float, default=4.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-named-functions", action="store_true", help="skip the update_public_key / generate_commitment / delta-epsilon timings")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
