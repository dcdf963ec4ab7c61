"""Multi-GPU plumbing: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in CPU tests).

Account batches and segmented MSMs shard into independent contiguous slices -- no data-path collective.
One large MSM has a single exchange step: every rank's partial sum (extended point, 4 x 32 canonical bytes = 128 B)
is all-gathered and the ranks' points are added (Edwards addition is not an NCCL reduce op, so
"all-reduce-to-root" = all-gather of 128 B + a local sum).  SURVEY.md section 8e.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of n items owned by `rank`."""
    lo = n * rank // world
    hi = n * (rank + 1) // world
    return lo, hi


def gather_partials(partial_xyzt, status, device=None):
    """All-gather the 128-byte partial sums and the per-rank status byte.  Returns (world x 128 uint8, world uint8)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    buf = np.zeros(132, np.uint8)
    buf[:128] = np.asarray(partial_xyzt, dtype=np.uint8).reshape(-1)
    buf[128] = int(status)
    if world == 1:
        return buf[:128].reshape(1, 128).copy(), buf[128:129].copy()
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    allb = torch.stack(out).cpu().numpy()
    return allb[:, :128].copy(), allb[:, 128].copy()


def msm_sharded(partial_fn, sum_fn, scalars, points, device=None):
    """Large MSM across ranks.  `partial_fn(scalars, points) -> (xyzt 128 B, status)` runs on this rank's slice,
    `sum_fn(k x 128 B) -> (32 B compressed, is_identity)` adds the gathered partial sums.
    Every rank returns the same (compressed point, status)."""
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    scalars = np.asarray(scalars, dtype=np.uint8).reshape(-1, 32)
    points = np.asarray(points, dtype=np.uint8).reshape(-1, 32)
    lo, hi = shard_range(scalars.shape[0], rank, world)
    part, st = partial_fn(scalars[lo:hi], points[lo:hi])
    parts, sts = gather_partials(part, st, device)
    bad = [int(s) for s in sts if s]
    if bad:
        return np.zeros(32, np.uint8), bad[0]   # lowest rank = earliest slice = first failing term
    out, ident = sum_fn(parts.reshape(-1))
    return out, 0


def verify_sharded(verify_fn, arrays, nproofs, device=None):
    """Independent proofs across ranks (BASELINE configs[2]: 4 096 shuffle proofs over 8 GPUs): every rank verifies its
    contiguous slice, the status bytes are all-gathered (the only exchange: one byte per proof).
    `arrays`: per-proof arrays (nproofs x bytes each); `verify_fn(*slices) -> status (uint8 per proof)`.
    Every rank returns the full nproofs status array."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    arrays = [np.asarray(a, dtype=np.uint8).reshape(nproofs, -1) for a in arrays]
    lo, hi = shard_range(nproofs, rank, world)
    local = np.asarray(verify_fn(*[a[lo:hi] for a in arrays]), dtype=np.uint8).reshape(-1)
    if local.size != hi - lo:
        raise ValueError("verify_fn returned %d verdicts for %d proofs" % (local.size, hi - lo))
    if world == 1:
        return local.copy()
    width = (nproofs + world - 1) // world          # slices differ by at most one proof: pad to the widest
    buf = np.zeros(width, np.uint8)
    buf[:local.size] = local
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    full = np.zeros(nproofs, np.uint8)
    for r in range(world):
        a, b = shard_range(nproofs, r, world)
        full[a:b] = out[r].cpu().numpy()[:b - a]
    return full


def msm_sharded_engine(engine, scalars, points, device):
    """The same exchange with the CUDA engine on both sides and the partial sums kept in device memory: this rank's slice through
    qq_msm_partial (result record = 128 B X | Y | Z | T + status byte, 144 B), the records all-gathered into ONE device tensor
    (NCCL: all_gather_into_tensor, no host bounce; under a gloo group the record travels through host memory) and added and
    encoded by one kernel on every rank (qq_points_sum_dev).  Returns (32-byte compressed point, status) on every rank."""
    import ctypes
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    scalars = np.asarray(scalars, dtype=np.uint8).reshape(-1, 32)
    points = np.asarray(points, dtype=np.uint8).reshape(-1, 32)
    lo, hi = shard_range(scalars.shape[0], rank, world)
    s_d = torch.from_numpy(np.ascontiguousarray(scalars[lo:hi]).reshape(-1).copy()).to(device)
    p_d = torch.from_numpy(np.ascontiguousarray(points[lo:hi]).reshape(-1).copy()).to(device)
    if hi == lo:      # empty slice: the kernels still want valid pointers
        s_d = torch.zeros(32, dtype=torch.uint8, device=device)
        p_d = torch.zeros(32, dtype=torch.uint8, device=device)
    rec = torch.zeros(144, dtype=torch.uint8, device=device)
    vp = ctypes.c_void_p
    engine.call_dev("qq_msm_partial_dev", vp(s_d.data_ptr()), vp(p_d.data_ptr()), ctypes.c_size_t(hi - lo), vp(rec.data_ptr()),
                    vp(rec.data_ptr() + 128))
    if world == 1:
        recs = rec
    elif dist.get_backend() == "nccl":
        recs = torch.zeros(144 * world, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(recs, rec)
    else:
        parts = [torch.zeros(144, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(parts, rec.cpu())
        recs = torch.cat(parts).to(device)
    out, _, st = engine.points_sum_dev(recs.data_ptr(), world, 144)
    return out, st
