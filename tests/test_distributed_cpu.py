"""CPU test of the N > 1 path (gloo, world_size 2): sharding + all-gather of the 128-byte MSM partial sums.
The partial/sum callbacks are played by the oracle here (no GPU in this container); on the GPU box the same
plumbing is exercised with the CUDA engine by bench.py --gpus N."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import __graft_entry__ as g
    import ristretto_ref as R
    from qq_testlib import Stream, sb
    g.load_package()
    from quisquis_rust_b200 import distributed as D
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    st = Stream(b"dist")
    hs = [st.scalar() for _ in range(n)]
    a = [st.scalar() for _ in range(n)]
    pts = np.frombuffer(b"".join(R.compress(R.mul(h, R.BASEPOINT)) for h in hs), np.uint8)
    sc = np.frombuffer(b"".join(sb(x) for x in a), np.uint8)

    def partial_fn(s, p):  # oracle stands in for qq_msm_partial
        acc = R.IDENTITY
        for i in range(s.shape[0]):
            q_ = R.decompress(p[i].tobytes())
            if q_ is None:
                return np.zeros(128, np.uint8), 1
            acc = R.add(acc, R.mul(int.from_bytes(s[i].tobytes(), "little"), q_))
        return np.frombuffer(b"".join(c.to_bytes(32, "little") for c in acc), np.uint8), 0

    def sum_fn(parts):  # oracle stands in for qq_points_sum
        acc = R.IDENTITY
        for k in range(parts.size // 128):
            c = [int.from_bytes(parts[k * 128 + 32 * j:k * 128 + 32 * j + 32].tobytes(), "little") for j in range(4)]
            acc = R.add(acc, tuple(c))
        return np.frombuffer(R.compress(acc), np.uint8), R.is_identity(acc)
    out, status = D.msm_sharded(partial_fn, sum_fn, sc, pts)
    exp = R.compress(R.mul(sum(x * h for x, h in zip(a, hs)) % R.L, R.BASEPOINT))
    lo, hi = D.shard_range(n, rank, world)
    # a bad point in rank 1's slice must fail the whole MSM on every rank
    pts2 = pts.copy().reshape(-1, 32)
    pts2[n - 1] = 0xff
    out2, status2 = D.msm_sharded(partial_fn, sum_fn, sc, pts2.reshape(-1))
    q.put((rank, out.tobytes() == exp, status, (lo, hi), status2, out2.tobytes() == bytes(32)))
    dist.destroy_process_group()


def test_msm_sharded_gloo_world2():
    world, n = 2, 9
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(60)
    assert [r[0] for r in res] == [0, 1]
    assert all(r[1] and r[2] == 0 for r in res)
    assert res[0][3] == (0, 4) and res[1][3] == (4, 9)
    assert all(r[4] == 1 and r[5] for r in res)


def test_shard_range_covers_everything():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.load_package()
    from quisquis_rust_b200.distributed import shard_range
    for n in (0, 1, 7, 9, 1 << 20, (1 << 20) + 3):
        for w in (1, 2, 4, 8):
            cuts = [shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))


def _verify_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import __graft_entry__ as g
    import rangeproof_ref as RP
    from merlin_ref import Transcript
    g.load_package()
    from quisquis_rust_b200 import distributed as D
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    m = 1
    per = m * 32 + (9 + 12) * 32
    raw = np.fromfile(os.path.join(ROOT, "tests", "golden", "range_proofs_m1.bin"), dtype=np.uint8).reshape(-1, per)
    n = 5                                                   # uneven slices: 2 + 3
    rec = np.tile(raw, (2, 1))[:n].copy()
    rec[1, 32 + 5 * 32] ^= 1                                # rank 0's slice
    rec[4, 32 + 6 * 32] ^= 1                                # rank 1's slice
    seen = []

    def verify_fn(cm, pr):  # the oracle's verifier stands in for qq_verify_range_proof_batch
        seen.append(cm.shape[0])
        out = []
        for i in range(cm.shape[0]):
            tr = Transcript(b"SenderAccountProof")
            tr.domain_sep(b"BulletProof")
            tr.domain_sep(b"AggregateBulletProof")
            ok = RP.verify_multiple(tr, pr[i].tobytes(), [cm[i].tobytes()], 64, bp_gens=RP.BulletproofGens(64, 1))
            out.append(0 if ok else 6)
        return np.array(out, np.uint8)
    full = D.verify_sharded(verify_fn, [rec[:, :32], rec[:, 32:]], n)
    q.put((rank, full.tolist(), seen))
    dist.destroy_process_group()


def test_verify_sharded_gloo_world2():
    """Proof batches shard into contiguous slices per rank (no data-path collective), the verdicts are all-gathered: both ranks
    end with the same five verdicts, each having verified only its own slice."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_verify_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(60)
    assert [r[0] for r in res] == [0, 1]
    assert res[0][1] == res[1][1] == [0, 6, 0, 0, 6]
    assert res[0][2] == [2] and res[1][2] == [3]
