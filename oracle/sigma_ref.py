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
