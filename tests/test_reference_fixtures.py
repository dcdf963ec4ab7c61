"""Loader for golden vectors dumped from the reference itself (rust/qq-golden -> tests/golden/ref/reference_vectors.bin).
The generator cannot run in the build container (no Rust toolchain), so these tests are skipped until a maintainer drops the
file in; they are what pins the oracle and the library to curve25519-dalek / bulletproofs / quisquislib output bytes."""
import os
import struct

import numpy as np
import pytest

import ristretto_ref as R

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref", "reference_vectors.bin")
needs_file = pytest.mark.skipif(not os.path.exists(PATH), reason="no reference-generated vectors (run rust/qq-golden with a Rust toolchain)")


def records():
    data = open(PATH, "rb").read()
    o, out = 0, []
    while o < len(data):
        tag, ln = struct.unpack_from("<IQ", data, o)
        o += 12
        out.append((tag, data[o:o + ln]))
        o += ln
    return out


def blobs(payload, count):
    o, out = 0, []
    for _ in range(count):
        (ln,) = struct.unpack_from("<Q", payload, o)
        o += 8
        out.append(payload[o:o + ln])
        o += ln
    return out


@needs_file
def test_oracle_against_reference_vectors():
    """CPU: update_account, delta / epsilon accounts and the generator chains of the oracle equal the reference's bytes."""
    for tag, p in records():
        if tag == 1:
            for i in range(len(p) // 352):
                r = p[352 * i:352 * (i + 1)]
                exp, st = R.update_account(r[:128], r[128:160], r[160:192], r[192:224])
                assert st == 0 and exp == r[224:352], i
        elif tag == 6:
            import rangeproof_ref as RP
            g = RP.BulletproofGens(64, 16)
            got = b"".join(g.G(party)[k] for party in range(2) for k in range(4)) if hasattr(g, "G") else None
            if got is not None:
                # layout: party 0 G[0..4) H[0..4), party 1 G[0..4) H[0..4)
                exp = b"".join(p[256 * party:256 * party + 128] for party in range(2))
                assert got == exp
        elif tag == 7:
            h, g = R.vector_pedersen_gens(4)
            assert h + b"".join(g[:3]) == p


@needs_file
@pytest.mark.gpu
def test_library_against_reference_vectors(engine, pkg):
    """GPU: the library on the reference's own vectors - update_account bytes, a bincode ShuffleProof verifies (through the wire
    format helpers), the DLOG sigma proof verifies, the bulletproofs crate's range proofs verify, BulletproofGens match."""
    from quisquis_rust_b200 import binding as B
    for tag, p in records():
        if tag == 1:
            n = len(p) // 352
            rec = np.frombuffer(p, np.uint8).reshape(n, 352)
            out, st = engine.update_account(rec[:, :128].copy(), rec[:, 128:160].copy(), rec[:, 160:192].copy(), rec[:, 192:224].copy())
            assert not st.any() and np.array_equal(out, rec[:, 224:352])
        elif tag == 2:
            (n,) = struct.unpack_from("<Q", p, 0)
            a = np.frombuffer(p[8:], np.uint8)
            acc, bl, r = a[:128 * n], a[128 * n:160 * n], a[160 * n:192 * n]
            delta, eps = a[192 * n:320 * n], a[320 * n:448 * n]
            d, e, st = engine.delta_epsilon(acc, bl, r, R.BASEPOINT_COMPRESSED + R.PEDERSEN_H_COMPRESSED)
            assert not st.any() and d.tobytes() == delta.tobytes() and e.tobytes() == eps.tobytes()
        elif tag == 3:
            bi, bo, bs, bp = blobs(p, 4)
            si, _ = B.accounts_from_bincode(bi)
            so, _ = B.accounts_from_bincode(bo)
            stm, _ = B.shuffle_statements_from_bincode(bs, 1)
            prf, _ = B.shuffle_proofs_from_bincode(bp, 1)
            st, sg, det = engine.verify_shuffle(si, so, stm, prf)
            assert (int(st[0]), int(sg[0])) == (0, 0)
        elif tag == 4:
            bu, bd, bsg = blobs(p, 3)
            upd, _ = B.accounts_from_bincode(bu)
            dlt, _ = B.accounts_from_bincode(bd)
            kind, vecs, x, _ = B.sigma_proof_from_bincode(bsg)
            assert kind == "dlog"
            st = engine.verify_update_account_dlog(upd, dlt, vecs[0], x, upd.shape[0], b"UpdateAccount", b"DLOGProof")
            assert int(st[0]) == 0
        elif tag == 5:
            (m,) = struct.unpack_from("<I", p, 0)
            cm, prf = p[4:4 + 32 * m], p[4 + 32 * m:]
            st = engine.verify_range_proofs(cm, prf, m, transcript_label=b"qq-golden", verifier_label=None, domain_label=None)
            assert int(st[0]) == 0
        elif tag == 6:
            g, h = engine.bulletproof_gens(64, 16)
            g, h = np.asarray(g).reshape(16, 64, 32), np.asarray(h).reshape(16, 64, 32)
            got = b"".join(g[party, :4].tobytes() + h[party, :4].tobytes() for party in range(2))
            assert got == p
