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


# =====================================================================================================================
# the remaining sigma proofs of src/accounts/{prover,verifier}.rs.  Blinding scalars are arguments (the reference draws
# them from a transcript RNG; any value gives a valid proof).
# =====================================================================================================================
def _mulc(k, point_bytes):
    """enc(k * P) for a compressed P (prover side: the reference unwrap()s the decompression)."""
    return R.compress(R.mul(k % R.L, R.decompress(point_bytes)))


def _msm_or_none(scalars, points):
    out, st = R.msm([_scalar_bytes(s) for s in scalars], points)
    return None if st else out


def create_epsilon_account(base_pk, rscalar, bl):
    """Account::create_epsilon_account, src/accounts/accounts.rs:298-311."""
    if bl < 0:
        raise ValueError("Not enough balance in the sender account")
    comm, st = R.generate_commitment(base_pk, _scalar_bytes(rscalar), _scalar_bytes(bl))
    assert st == 0
    return base_pk + comm


# ---- sender account proof: prover src/accounts/prover.rs:355-500, verifier src/accounts/verifier.rs:396-470 -------------
def prove_account(delta_accounts, bl, sk, base_pk, eps_rscalars, blindings, transcript_label=b"SenderAccountProof",
                  prover_label=b"DLOGProof", tr=None):
    """-> (epsilon_accounts, zv[], zsk[], zr[], x).  blindings: per account (r_v, r_sk, r_dash).
    tr: the Prover's running transcript (Transcript::new + Prover::new already applied), continued in place."""
    if tr is None:
        tr = Transcript(transcript_label)
        tr.domain_sep(prover_label)
    tr.domain_sep(b"VerifyAccountProof")
    eps = [create_epsilon_account(base_pk, r, v) for r, v in zip(eps_rscalars, bl)]
    for d, e in zip(delta_accounts, eps):
        tr.append_account_var(b"delta_account", d)
        tr.append_account_var(b"epsilon_account", e)
    for d, e, (rv, rsk, rd) in zip(delta_accounts, eps, blindings):
        g_rv = R.mul(rv % R.L, R.decompress(e[0:32]))
        e_delta = _mulc(rsk, d[0:32])
        f_delta = R.compress(R.add(g_rv, R.mul(rsk % R.L, R.decompress(d[64:96]))))
        e_eps = _mulc(rd, e[0:32])
        f_eps = R.compress(R.add(g_rv, R.mul(rd % R.L, R.decompress(e[32:64]))))
        tr.append_point_var(b"e_delta", e_delta)
        tr.append_point_var(b"f_delta", f_delta)
        tr.append_point_var(b"e_epsilon", e_eps)
        tr.append_point_var(b"f_epsilon", f_eps)
    x = tr.get_challenge(b"challenge")
    zv = [(b[0] - v * x) % R.L for b, v in zip(blindings, bl)]
    zsk = [(b[1] - s * x) % R.L for b, s in zip(blindings, sk)]
    zr = [(b[2] - r * x) % R.L for b, r in zip(blindings, eps_rscalars)]
    return eps, zv, zsk, zr, x


def verify_account(delta_accounts, epsilon_accounts, base_pk, zv, zsk, zr, x, transcript_label=b"SenderAccountProof",
                   verifier_label=b"DLOGProof", tr=None):
    """-> True, False ("sender account verification failed") or None ("Account Verify: Failed").
    tr: the Verifier's running transcript, continued in place."""
    if tr is None:
        tr = Transcript(transcript_label)
        tr.domain_sep(verifier_label)
    tr.domain_sep(b"VerifyAccountProof")
    for d, e in zip(delta_accounts, epsilon_accounts):
        tr.append_account_var(b"delta_account", d)
        tr.append_account_var(b"epsilon_account", e)
    G, H = base_pk[0:32], base_pk[32:64]
    for d, e, v, s, r in zip(delta_accounts, epsilon_accounts, zv, zsk, zr):
        pts = [_msm_or_none([s, x], [d[0:32], d[32:64]]),
               _msm_or_none([v, s, x], [G, d[64:96], d[96:128]]),
               _msm_or_none([x, r], [e[64:96], G]),
               _msm_or_none([v, r, x], [G, H, e[96:128]])]
        if any(q is None for q in pts):
            return None
        for label, q in zip((b"e_delta", b"f_delta", b"e_epsilon", b"f_epsilon"), pts):
            tr.append_point_var(label, q)
    return tr.get_challenge(b"challenge") == x % R.L


# ---- zero balance: prover src/accounts/prover.rs:602-702, verifier src/accounts/verifier.rs:593-680 ---------------------
def prove_zero_balance(accounts, comm_rscalars, blindings, vector_form, transcript_label=b"ZeroBalanceAccount",
                       prover_label=b"DLOGProof", domain=None):
    """-> (z[], x).  The reference prover's vector-form separator is b"ZeroBalanceAccountVectorProof" (prover.rs:613);
    `domain` overrides it (the verifier spells it differently, verifier.rs:605)."""
    tr = Transcript(transcript_label)
    tr.domain_sep(prover_label)
    if domain is None:
        domain = b"ZeroBalanceAccountVectorProof" if vector_form else b"ZeroBalanceAccountProof"
    tr.domain_sep(domain)
    for a in accounts:
        tr.append_account_var(b"anonymity_account" if vector_form else b"zero_account", a)
    for a, r in zip(accounts, blindings):
        tr.append_point_var(b"e", _mulc(r, a[0:32]))
        tr.append_point_var(b"f", _mulc(r, a[32:64]))
    x = tr.get_challenge(b"challenge")
    return [(r - x * c) % R.L for r, c in zip(blindings, comm_rscalars)], x


def verify_zero_balance(accounts, z, x, vector_form, transcript_label=b"ZeroBalanceAccount", verifier_label=b"DLOGProof"):
    tr = Transcript(transcript_label)
    tr.domain_sep(verifier_label)
    tr.domain_sep(b"ZeroBalanceAccounVectorProof" if vector_form else b"ZeroBalanceAccountProof")
    for a in accounts:
        tr.append_account_var(b"anonymity_account" if vector_form else b"zero_account", a)
    for a, zi in zip(accounts, z):
        e = _msm_or_none([zi, x], [a[0:32], a[64:96]])
        f = _msm_or_none([zi, x], [a[32:64], a[96:128]])
        if e is None or f is None:
            return None
        tr.append_point_var(b"e", e)
        tr.append_point_var(b"f", f)
    return tr.get_challenge(b"challenge") == x % R.L


# ---- destroy account: prover src/accounts/prover.rs:715-770, verifier src/accounts/verifier.rs:693-735 ------------------
def prove_destroy_account(accounts, sk, blindings, transcript_label=b"DestroyAccount", prover_label=b"DLOGProof"):
    tr = Transcript(transcript_label)
    tr.domain_sep(prover_label)
    tr.domain_sep(b"DestroyAccountProof")
    for a in accounts:
        tr.append_account_var(b"account", a)
    for a, r in zip(accounts, blindings):
        tr.append_point_var(b"e", _mulc(r, a[0:32]))
        tr.append_point_var(b"f", _mulc(r, a[64:96]))
    x = tr.get_challenge(b"challenge")
    return [(r - x * s) % R.L for r, s in zip(blindings, sk)], x


def verify_destroy_account(accounts, z, x, transcript_label=b"DestroyAccount", verifier_label=b"DLOGProof"):
    tr = Transcript(transcript_label)
    tr.domain_sep(verifier_label)
    tr.domain_sep(b"DestroyAccountProof")
    for a in accounts:
        tr.append_account_var(b"account", a)
    for a, zi in zip(accounts, z):
        e = _msm_or_none([zi, x], [a[0:32], a[32:64]])
        f = _msm_or_none([zi, x], [a[64:96], a[96:128]])
        if e is None or f is None:
            return None
        tr.append_point_var(b"e", e)
        tr.append_point_var(b"f", f)
    return tr.get_challenge(b"challenge") == x % R.L


# ---- same value (ElGamal vs Pedersen): prover src/accounts/prover.rs:784-850, verifier src/accounts/verifier.rs:747-806 --
def pedersen_commit(value, blinding):
    """PedersenGens::default().commit(value, blinding) = value * B + blinding * B_blinding."""
    h = R.decompress(R.PEDERSEN_H_COMPRESSED)
    return R.compress(R.add(R.mul(value % R.L, R.BASEPOINT), R.mul(blinding % R.L, h)))


def prove_same_value(enc_account, rscalar, value, pedersen_commitment, r1_dash, v_doubledash):
    tr = Transcript(b"SameValueProof")
    tr.domain_sep(b"DLEQProof")
    tr.append_account_var(b"encrypted_account", enc_account)
    tr.append_point_var(b"G", R.BASEPOINT_COMPRESSED)
    tr.append_point_var(b"H", R.PEDERSEN_H_COMPRESSED)
    tr.append_point_var(b"d", pedersen_commitment)
    gv = R.mul(v_doubledash % R.L, R.BASEPOINT)
    f_delta = R.add(gv, R.mul(r1_dash % R.L, R.decompress(enc_account[32:64])))
    f_eps = R.add(gv, R.mul(r1_dash % R.L, R.decompress(R.PEDERSEN_H_COMPRESSED)))
    tr.append_point_var(b"f_delta", R.compress(f_delta))
    tr.append_point_var(b"f_epsilon", R.compress(f_eps))
    x = tr.get_challenge(b"challenge")
    return (v_doubledash - x * value) % R.L, (r1_dash - rscalar * x) % R.L, x


def verify_same_value(enc_account, commitment, zv, zr, x):
    tr = Transcript(b"SameValueProof")
    tr.domain_sep(b"DLEQProof")
    tr.append_account_var(b"encrypted_account", enc_account)
    tr.append_point_var(b"G", R.BASEPOINT_COMPRESSED)
    tr.append_point_var(b"H", R.PEDERSEN_H_COMPRESSED)
    tr.append_point_var(b"d", commitment)
    f_enc = _msm_or_none([zr, x, zv], [enc_account[32:64], enc_account[96:128], R.BASEPOINT_COMPRESSED])
    f_ped = _msm_or_none([zr, x, zv], [R.PEDERSEN_H_COMPRESSED, commitment, R.BASEPOINT_COMPRESSED])
    if f_enc is None or f_ped is None:
        return None
    tr.append_point_var(b"f_delta", f_enc)
    tr.append_point_var(b"f_epsilon", f_ped)
    return tr.get_challenge(b"challenge") == x % R.L


# ---- dark-transaction output update: prover src/accounts/prover.rs:864-950, verifier src/accounts/verifier.rs:818-917 ----
def prove_update_account_dark_tx(delta_accounts, output_accounts, pk_rscalar, comm_rscalar, pk_blinding, comm_blinding,
                                 transcript_label=b"UpdateAccount", prover_label=b"DLOGProof"):
    tr = Transcript(transcript_label)
    tr.domain_sep(prover_label)
    tr.domain_sep(b"VerifyUpdateAccountDarkTx")
    for d, o in zip(delta_accounts, output_accounts):
        diff, st = R.sub_commitments(o[64:], d[64:])
        pkc, st2 = R.update_public_key(d[:64], _scalar_bytes(comm_rscalar))
        if st or st2 or diff != pkc:
            raise ValueError("Commitments are not properly updated. Every Commitment should be updated with 0 balance")
    for d, o in zip(delta_accounts, output_accounts):
        tr.append_account_var(b"account", d)
        tr.append_account_var(b"updatedaccount", o)
    for d in delta_accounts:
        pk, _ = R.update_public_key(d[:64], _scalar_bytes(pk_blinding))
        tr.append_point_var(b"commitmentgr", pk[0:32])
        tr.append_point_var(b"commitmentgrsk", pk[32:64])
    for d in delta_accounts:
        pk, _ = R.update_public_key(d[:64], _scalar_bytes(comm_blinding))
        tr.append_point_var(b"commitmentc", pk[0:32])
        tr.append_point_var(b"commitmentd", pk[32:64])
    x = tr.get_challenge(b"challenge")
    return [(pk_blinding - x * pk_rscalar) % R.L, (comm_blinding - x * comm_rscalar) % R.L], x


def verify_update_account_dark_tx(delta_accounts, output_accounts, z_vector, x, transcript_label=b"UpdateAccount",
                                  verifier_label=b"DLOGProof"):
    """-> True / False / None (Err on an undecodable key); raises where the reference panics (undecodable commitment)."""
    tr = Transcript(transcript_label)
    tr.domain_sep(verifier_label)
    e = []
    for d, o in zip(delta_accounts, output_accounts):
        a = _msm_or_none([z_vector[0], x], [d[0:32], o[0:32]])
        b = _msm_or_none([z_vector[0], x], [d[32:64], o[32:64]])
        if a is None or b is None:
            return None
        e.append((a, b))
    diffs = []
    for d, o in zip(delta_accounts, output_accounts):
        diff, st = R.sub_commitments(o[64:], d[64:])
        if st:
            raise ValueError("called `Option::unwrap()` on a `None` value")
        diffs.append(diff)
    f = []
    for d, diff in zip(delta_accounts, diffs):
        a = _msm_or_none([z_vector[1], x], [d[0:32], diff[0:32]])
        b = _msm_or_none([z_vector[1], x], [d[32:64], diff[32:64]])
        if a is None or b is None:
            return None
        f.append((a, b))
    tr.domain_sep(b"VerifyUpdateAccountDarkTx")
    for d, o in zip(delta_accounts, output_accounts):
        tr.append_account_var(b"account", d)
        tr.append_account_var(b"updatedaccount", o)
    for a, b in e:
        tr.append_point_var(b"commitmentgr", a)
        tr.append_point_var(b"commitmentgrsk", b)
    for a, b in f:
        tr.append_point_var(b"commitmentc", a)
        tr.append_point_var(b"commitmentd", b)
    return tr.get_challenge(b"challenge") == x % R.L
