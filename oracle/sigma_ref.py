"""TEST INFRASTRUCTURE (oracle): the reference's "DLOG" sigma proof that delta accounts were updated correctly,
restated for the parity tests: prover src/accounts/prover.rs:264-343, verifier src/accounts/verifier.rs:223-292,
Account::update_delta_accounts src/accounts/accounts.rs:225-250.  Never imported by the product."""
import ristretto_ref as R
from merlin_ref import Transcript


def update_delta_accounts(updated_accounts, delta_accounts):
    out = []
    for u, d in zip(updated_accounts, delta_accounts):
        if u[:64] != d[:64]:
            raise ValueError("pks are not equal")
        comm, st = R.add_commitments(u[64:], d[64:])
        assert st == 0
        out.append(u[:64] + comm)
    return out


def _scalar_bytes(k):
    return (k % R.L).to_bytes(32, "little")


def prove_update_account_dlog(updated_input, updated_delta, delta_r, s_scalar, transcript_label=b"UpdateAccount",
                              prover_label=b"DLOGProof"):
    """-> (z_vector [ints], x int).  `delta_r` ints, `s_scalar` the blinding scalar (the reference draws it from a
    transcript RNG; any value gives a valid proof)."""
    tr = Transcript(transcript_label)
    tr.domain_sep(prover_label)                      # Prover::new
    # anonymity set: accounts whose commitment difference equals pk_delta * r  (prover.rs:272-296)
    idx = []
    for i, (inp, dl, r) in enumerate(zip(updated_input, updated_delta, delta_r)):
        diff, st = R.sub_commitments(dl[64:], inp[64:])
        pkr, st2 = R.update_public_key(dl[:64], _scalar_bytes(r))
        if st == 0 and st2 == 0 and diff == pkr:
            idx.append(i)
    tr.domain_sep(b"DLOGProof")
    for i in idx:
        tr.append_point_var(b"inputgr", updated_input[i][0:32])
        tr.append_point_var(b"inputgrsk", updated_input[i][32:64])
        tr.append_point_var(b"outputgr", updated_delta[i][0:32])
        tr.append_point_var(b"outputgrsk", updated_delta[i][32:64])
    for i in idx:
        pks, st = R.update_public_key(updated_input[i][:64], _scalar_bytes(s_scalar))
        tr.append_point_var(b"commitmentgr", pks[0:32])
        tr.append_point_var(b"commitmentgrsk", pks[32:64])
    x = tr.get_challenge(b"chal")
    z = [(s_scalar - x * delta_r[i]) % R.L for i in idx]
    return z, x


def verify_update_account_dlog(updated_input, updated_delta, z_vector, x, transcript_label=b"UpdateAccount",
                               verifier_label=b"DLOGProof"):
    """-> True / False ("DLOG Proof Verify: Failed"); raises on an undecodable commitment like the reference panics."""
    tr = Transcript(transcript_label)
    tr.domain_sep(verifier_label)                    # Verifier::new
    e11, e12 = [], []
    for inp, dl, z in zip(updated_input, updated_delta, z_vector):
        a, st = R.sub_commitments(dl[64:], inp[64:])
        if st:
            raise ValueError("called `Option::unwrap()` on a `None` value")
        for pt_key, pt_a, acc in ((inp[0:32], a[0:32], e11), (inp[32:64], a[32:64], e12)):
            out, st = R.msm([_scalar_bytes(z), _scalar_bytes(x)], [pt_key, pt_a])
            if st:
                return False
            acc.append(out)
    tr.domain_sep(b"DLOGProof")
    for inp, dl in zip(updated_input, updated_delta):
        tr.append_point_var(b"inputgr", inp[0:32])
        tr.append_point_var(b"inputgrsk", inp[32:64])
        tr.append_point_var(b"outputgr", dl[0:32])
        tr.append_point_var(b"outputgrsk", dl[32:64])
    for a, b in zip(e11, e12):
        tr.append_point_var(b"commitmentgr", a)
        tr.append_point_var(b"commitmentgrsk", b)
    return tr.get_challenge(b"chal") == x % R.L


# ---- "delta compact" DLEQ proof: prover src/accounts/prover.rs:120-254, verifier src/accounts/verifier.rs:138-209 ----
def prove_delta_compact(delta_accounts, epsilon_accounts, rscalar, value_vector, blindings, transcript_label=b"DeltaCompact",
                        prover_label=b"DLEQProof"):
    """-> (zv[], zr1[], zr2[], x) as ints.  `blindings`: per account (r1', r2', v'') (the reference draws them from a
    transcript RNG)."""
    tr = Transcript(transcript_label)
    tr.domain_sep(prover_label)
    tr.domain_sep(b"VerifyDeltaCompact")
    for d, e in zip(delta_accounts, epsilon_accounts):
        tr.append_account_var(b"delta_account", d)
        tr.append_account_var(b"epsilon_account", e)
    for d, e, (r1, r2, vd) in zip(delta_accounts, epsilon_accounts, blindings):
        gv = R.mul(vd % R.L, R.BASEPOINT)
        e_delta = R.mul(r1 % R.L, R.decompress(d[0:32]))
        f_delta = R.add(gv, R.mul(r1 % R.L, R.decompress(d[32:64])))
        e_eps = R.mul(r2 % R.L, R.decompress(e[0:32]))
        f_eps = R.add(gv, R.mul(r2 % R.L, R.decompress(e[32:64])))
        tr.append_point_var(b"e_delta", R.compress(e_delta))
        tr.append_point_var(b"f_delta", R.compress(f_delta))
        tr.append_point_var(b"e_epsilon", R.compress(e_eps))
        tr.append_point_var(b"f_epsilon", R.compress(f_eps))
    x = tr.get_challenge(b"challenge")
    zv = [(b[2] - v * x) % R.L for b, v in zip(blindings, value_vector)]
    zr1 = [(b[0] - r * x) % R.L for b, r in zip(blindings, rscalar)]
    zr2 = [(b[1] - r * x) % R.L for b, r in zip(blindings, rscalar)]
    return zv, zr1, zr2, x


def verify_delta_compact(delta_accounts, epsilon_accounts, zv, zr1, zr2, x, transcript_label=b"DeltaCompact",
                         verifier_label=b"DLEQProof"):
    tr = Transcript(transcript_label)
    tr.domain_sep(verifier_label)
    tr.domain_sep(b"VerifyDeltaCompact")
    for d, e in zip(delta_accounts, epsilon_accounts):
        tr.append_account_var(b"delta_account", d)
        tr.append_account_var(b"epsilon_account", e)
    xb = _scalar_bytes(x)
    for d, e, v, r1, r2 in zip(delta_accounts, epsilon_accounts, zv, zr1, zr2):
        pts = []
        for acc, zr in ((d, r1), (e, r2)):
            out, st = R.msm([_scalar_bytes(zr), xb], [acc[0:32], acc[64:96]])
            if st:
                return False
            pts.append(out)
            out, st = R.msm([_scalar_bytes(zr), xb, _scalar_bytes(v)], [acc[32:64], acc[96:128], R.BASEPOINT_COMPRESSED])
            if st:
                return False
            pts.append(out)
        tr.append_point_var(b"e_delta", pts[0])
        tr.append_point_var(b"f_delta", pts[1])
        tr.append_point_var(b"e_epsilon", pts[2])
        tr.append_point_var(b"f_epsilon", pts[3])
    return tr.get_challenge(b"challenge") == x % R.L
