"""Shared helpers for the parity tests: seeded synthetic inputs and edge cases."""
import hashlib

import numpy as np

import ristretto_ref as R

SEED = b"QUISQUIS"


class Stream:
    """Counter-mode SHAKE256 stream keyed by the ASCII seed 'QUISQUIS' (0x5155495351554953)."""

    def __init__(self, label=b""):
        self.label = label
        self.ctr = 0

    def bytes(self, n):
        out = hashlib.shake_256(SEED + self.label + self.ctr.to_bytes(8, "little")).digest(n)
        self.ctr += 1
        return out

    def scalar(self):
        return int.from_bytes(self.bytes(64), "little") % R.L

    def scalar_bytes(self):
        return self.scalar().to_bytes(32, "little")


def sb(k):
    return (k % R.L).to_bytes(32, "little")


def cat(items):
    return np.frombuffer(b"".join(items), dtype=np.uint8).copy()


def make_account(st, value=0):
    """Account as in the reference tests (src/shuffle/shuffle.rs:762-768): pk = (rho*B, sk*rho*B), comm = commit(pk, k, v)."""
    sk, rho, k = st.scalar(), st.scalar(), st.scalar()
    gr = R.mul(rho, R.BASEPOINT)
    pk = R.compress(gr) + R.compress(R.mul(sk, gr))
    comm, s = R.generate_commitment(pk, sb(k), sb(value))
    assert s == 0
    return pk + comm, sk, k


def invalid_encodings():
    """One representative per reject class of RFC 9496 4.3.1 (each verified against the oracle in test_oracle.py)."""
    out = []
    out.append(("non_canonical_p", R.P.to_bytes(32, "little")))
    out.append(("non_canonical_p_plus_2", (R.P + 2).to_bytes(32, "little")))
    out.append(("bit255_set", (2 | (1 << 255)).to_bytes(32, "little")))
    out.append(("all_ff", b"\xff" * 32))
    out.append(("negative_s", (1).to_bytes(32, "little")))
    # search small even s for the remaining classes
    found = {}
    s = 2
    while len(found) < 3 and s < 4000:
        b = s.to_bytes(32, "little")
        if R.decompress(b) is None:
            ss = s * s % R.P
            u1, u2 = (1 - ss) % R.P, (1 + ss) % R.P
            v = (-(R.D * u1 * u1) - u2 * u2) % R.P
            ok, inv = R.sqrt_ratio_i(1, v * u2 * u2 % R.P)
            if not ok:
                found.setdefault("non_square", b)
            else:
                dx = inv * u2 % R.P
                x = R._abs(2 * s * dx % R.P)
                y = u1 * (inv * dx % R.P * v % R.P) % R.P
                if y == 0:
                    found.setdefault("y_zero", b)
                elif (x * y % R.P) & 1:
                    found.setdefault("negative_t", b)
        s += 2
    out.extend(found.items())
    return out


# ---- sigma-proof scenarios of the reference's own tests (src/accounts/verifier.rs tests), shared by the oracle round-trip
# test (CPU) and the batched-verifier parity tests (GPU).  Proofs come from the oracle's prover restatements.
def random_account_with_value(st, value):
    """Account::generate_random_account_with_value (src/accounts/accounts.rs:331-347) -> (account, sk)."""
    acc, sk, _ = make_account(st, 0)
    upd, s = R.update_account(acc, sb(value), st.scalar_bytes(), st.scalar_bytes())
    assert s == 0
    return upd, sk


def scenario_dark_tx(st, n=4):
    """verifier.rs:1075-1111 -> (delta_accounts, output_accounts, z[2], x)."""
    import sigma_ref as S
    u, c = st.scalar(), st.scalar()
    accs = [random_account_with_value(st, 10)[0] for _ in range(n)]
    outs = []
    for a in accs:
        o, s = R.update_account(a, sb(0), sb(u), sb(c))
        assert s == 0
        outs.append(o)
    z, x = S.prove_update_account_dark_tx(accs, outs, u, c, st.scalar(), st.scalar())
    return accs, outs, z, x


def zero_balance_accounts(st, n):
    """verifier.rs:1386-1393 / :1408-1421: zero-balance accounts on re-keyed base keys -> (accounts, comm_rscalars)."""
    pk, s = R.update_public_key(R.BASE_PK, st.scalar_bytes())
    accs, rs = [], []
    for _ in range(n):
        k, s = R.update_public_key(pk, st.scalar_bytes())
        r = st.scalar()
        comm, s = R.generate_commitment(k, sb(r), sb(0))
        accs.append(k + comm)
        rs.append(r)
        pk, s = R.update_public_key(pk, st.scalar_bytes())
    return accs, rs


def scenario_destroy(st, n=4):
    """verifier.rs:1455-1479 -> (accounts, z[], x)."""
    import sigma_ref as S
    pairs = [random_account_with_value(st, 0) for _ in range(n)]
    accs, sks = [p[0] for p in pairs], [p[1] for p in pairs]
    z, x = S.prove_destroy_account(accs, sks, [st.scalar() for _ in range(n)])
    return accs, z, x


def scenario_same_value(st, value=10, committed=None):
    """verifier.rs:1736-1775 -> (enc_account, pedersen_commitment, zv, zr, x); committed != value is the fail test."""
    import sigma_ref as S
    sk, rho, r = st.scalar(), st.scalar(), st.scalar()
    gr = R.mul(rho, R.BASEPOINT)
    pk = R.compress(gr) + R.compress(R.mul(sk, gr))
    comm, s = R.generate_commitment(pk, sb(r), sb(value))
    acc = pk + comm
    pc = S.pedersen_commit(value if committed is None else committed, r)
    zv, zr, x = S.prove_same_value(acc, r, value, pc, st.scalar(), st.scalar())
    return acc, pc, zv, zr, x


def scenario_sender_account(st):
    """The reference's (commented-out) verify_account_verifier test, verifier.rs:1115-1216: 9 accounts of value 10,
    transfers [-5, -3, 5, 3, 0 x 5], the two senders prove their remaining balances 5 and 7.
    -> (updated_delta_sender[2], epsilon[2], base_pk, zv, zsk, zr, x)."""
    import sigma_ref as S
    vals = [R.L - 5, R.L - 3, 5, 3, 0, 0, 0, 0, 0]
    pairs = [random_account_with_value(st, 10) for _ in range(9)]
    accs, sks = [p[0] for p in pairs], [p[1] for p in pairs]
    delta = [R.delta_epsilon(a, sb(v), st.scalar_bytes())[0] for a, v in zip(accs, vals)]
    upd = S.update_delta_accounts(accs, delta)
    senders, bl = upd[0:2], [5, 7]
    eps, zv, zsk, zr, x = S.prove_account(senders, bl, sks[0:2], R.BASE_PK, [st.scalar(), st.scalar()],
                                          [(st.scalar(), st.scalar(), st.scalar()) for _ in range(2)])
    return senders, eps, R.BASE_PK, zv, zsk, zr, x


def scenario_range_batch(st):
    """verifier.rs verify_non_negative_sender_receiver_bulletproof_batch_verifier_test (:1525-1628): the sender-account proof
    of the two senders and then, on the SAME transcript (b"SenderAccountProof" / b"BulletProof"), one aggregated range proof
    over [5, 7, 5, 3] (senders' remaining balances, receivers' amounts).
    -> (senders[2], eps_sender[2], base_pk, zv, zsk, zr, x, epsilon_accounts_bp[4], proof bytes)."""
    import rangeproof_ref as RP
    import sigma_ref as S
    from merlin_ref import Transcript
    vals = [R.L - 5, R.L - 3, 5, 3, 0, 0, 0, 0, 0]
    pairs = [random_account_with_value(st, 10) for _ in range(9)]
    accs, sks = [p[0] for p in pairs], [p[1] for p in pairs]
    r_scalars = [st.scalar() for _ in range(9)]
    de = [R.delta_epsilon(a, sb(v), sb(r)) for a, v, r in zip(accs, vals, r_scalars)]
    delta, epsilon = [d[0] for d in de], [d[1] for d in de]
    upd = S.update_delta_accounts(accs, delta)
    senders, bl = upd[0:2], [5, 7]
    tr = Transcript(b"SenderAccountProof")
    tr.domain_sep(b"BulletProof")                      # Prover::new
    rs_sender = [st.scalar(), st.scalar()]
    eps, zv, zsk, zr, x = S.prove_account(senders, bl, sks[0:2], R.BASE_PK, rs_sender,
                                          [(st.scalar(), st.scalar(), st.scalar()) for _ in range(2)], tr=tr)
    proofs = RP.quisquis_range_prover(tr, [5, 7, 5, 3], [rs_sender[0], rs_sender[1], r_scalars[2], r_scalars[3]], st.scalar)
    return senders, eps, R.BASE_PK, zv, zsk, zr, x, [eps[0], eps[1], epsilon[2], epsilon[3]], proofs[0]


def scenario_range_vector(st, values=(5, 3, 0, 0, 0), label=b"Test_notPower"):
    """prover.rs verify_non_negative_sender_receiver_prover_test (:965-992), the odd-sized case: one single-value proof per
    balance on one transcript -> (epsilon accounts, [proof bytes])."""
    import rangeproof_ref as RP
    import sigma_ref as S
    from merlin_ref import Transcript
    tr = Transcript(label)
    tr.domain_sep(b"Bulletproof")
    rs = [st.scalar() for _ in values]
    eps = [S.create_epsilon_account(R.BASE_PK, r, v) for r, v in zip(rs, values)]
    return eps, RP.quisquis_range_prover(tr, list(values), rs, st.scalar)


# ---- leaf arguments of the shuffle proof: the reference's own test scenarios (src/shuffle/{ddh,singlevalueproduct,
# hadamard}.rs tests) with proofs from the oracle's prover restatements (oracle/shuffle_ref.py)
def scenario_ddh(st):
    """ddh.rs:162-194 -> (G, H, G_dash, H_dash, challenge, z); transcript b"ShuffleProof" / b"DDHTuple"."""
    import shuffle_ref as F
    pks = [make_account(st, 0)[0][:64] for _ in range(9)]
    g_i, h_i = [p[:32] for p in pks], [p[32:] for p in pks]
    x, rho = st.scalar(), st.scalar()
    exp_x = F.exp_iter(x, 9, skip=1)
    G, H = R.compress(F._msm_point(exp_x, g_i)), R.compress(F._msm_point(exp_x, h_i))
    tr = F.new_transcript(b"ShuffleProof", b"DDHTuple")
    (challenge, z), (G_dash, H_dash) = F.ddh_prove(tr, g_i, h_i, exp_x, G, H, rho, st.scalar())
    return G, H, G_dash, H_dash, challenge, z


def scenario_svp(st, pi=(7, 6, 1, 5, 3, 4, 2, 8, 9)):
    """singlevalueproduct.rs:269-327 -> (commitment_a, b, proof dict); transcript b"SingleValue" / b"Shuffle"."""
    import shuffle_ref as F
    xpc = F.XpcGens(4)
    rows = [pi[0:3], pi[3:6], pi[6:9]]
    bvec = [r[0] * r[1] * r[2] % R.L for r in rows]
    s = st.scalar()
    cb = xpc.commit(bvec, s)
    b = bvec[0] * bvec[1] * bvec[2] % R.L
    tr = F.new_transcript(b"SingleValue", b"Shuffle")
    proof = F.svp_prove(tr, xpc, s, bvec, [st.scalar() for _ in range(3)], st.scalar(), [st.scalar()], st.scalar(),
                        st.scalar())
    return cb, b, proof


def scenario_hadamard(st, random_matrices=False):
    """hadamard.rs:395-470 -> (omega, commit_a, commit_b, commit_c, proof dict); transcript b"Hadamard" / b"Shuffle"."""
    import shuffle_ref as F
    xpc = F.XpcGens(4)
    if random_matrices:
        a = [[st.scalar() for _ in range(3)] for _ in range(3)]
        b = [[st.scalar() for _ in range(3)] for _ in range(3)]
        r, s, t = ([st.scalar() for _ in range(3)] for _ in range(3))
    else:
        av, bv = (7, 6, 1, 5, 3, 4, 2, 8, 9), (3, 2, 1, 7, 3, 5, 8, 3, 6)
        a, b = [list(av[3 * i:3 * i + 3]) for i in range(3)], [list(bv[3 * i:3 * i + 3]) for i in range(3)]
        r, s, t = [6, 2, 5], [7, 1, 3], [5, 2, 1]
    c = [[x * y % R.L for x, y in zip(ra, rb)] for ra, rb in zip(a, b)]
    ca = [xpc.commit(a[i], r[i]) for i in range(3)]
    cb = [xpc.commit(b[i], s[i]) for i in range(3)]
    cc = [xpc.commit(c[i], t[i]) for i in range(3)]
    rnd = {"a_0": [st.scalar() for _ in range(3)], "b_0": [st.scalar() for _ in range(3)], "r_0": st.scalar(),
           "s_0": st.scalar(), "t_0": st.scalar(), "omega": [st.scalar() for _ in range(3)],
           "rho": [st.scalar() for _ in range(4)]}
    tr = F.new_transcript(b"Hadamard", b"Shuffle")
    proof, omega = F.hadamard_prove(tr, xpc, a, b, c, ca, cb, cc, r, s, t, rnd)
    return omega, ca, cb, cc, proof


def scenario_product(st, pi=(7, 6, 1, 5, 3, 4, 2, 8, 9)):
    """product.rs product_proof_test (:595-640) -> (c_prod_A[3], proof, statement); transcript b"ShuffleProof" / b"Shuffle"."""
    import shuffle_ref as F
    xpc = F.XpcGens(4)
    rows = [list(pi[0:3]), list(pi[3:6]), list(pi[6:9])]
    r = [st.scalar() for _ in range(3)]
    cols = F.columns(rows)
    c_prod_A = [xpc.commit(cols[i], r[i]) for i in range(3)]
    rnd = {"s": st.scalar(),
           "mh": {"s_mid": st.scalar(),
                  "zero": {"a_0": [st.scalar() for _ in range(3)], "b_m": [st.scalar() for _ in range(3)], "r_0": st.scalar(),
                           "s_m": st.scalar(), "t": [st.scalar() for _ in range(7)]}},
           "svp": ([st.scalar() for _ in range(3)], st.scalar(), [st.scalar()], st.scalar(), st.scalar())}
    tr = F.new_transcript(b"ShuffleProof", b"Shuffle")
    proof, statement = F.product_prove(tr, xpc, rows, r, rnd)
    return c_prod_A, proof, statement


def scenario_shuffle(st, perm=None):
    """shuffle.rs shuffle_proof_test (:759-795): 9 zero-balance accounts, Shuffle::input_shuffle, create_shuffle_proof.
    -> (shuffle_input[9], shuffle_output[9], proof, statement); transcript b"ShuffleProof" / b"Shuffle"."""
    import shuffle_ref as F
    xpc = F.XpcGens(4)
    accounts = [make_account(st, 0)[0] for _ in range(9)]
    if perm is None:
        perm = list(range(1, 10))
        for i in range(8, 0, -1):
            j = st.scalar() % (i + 1)
            perm[i], perm[j] = perm[j], perm[i]
    sh = F.input_shuffle(accounts, perm, [st.scalar() for _ in range(9)], st.scalar())
    v3 = lambda: [st.scalar() for _ in range(3)]  # noqa: E731
    mexp = lambda: {"a_0": v3(), "r_0": st.scalar(), "b_vec": [st.scalar() for _ in range(6)],  # noqa: E731
                    "s_vec": [st.scalar() for _ in range(6)], "tau_vec": [st.scalar() for _ in range(6)]}
    rnd = {"r": v3(), "r_dash": v3(), "s": v3(), "s_dash": v3(),
           "hadamard": {"a_0": v3(), "b_0": v3(), "r_0": st.scalar(), "s_0": st.scalar(), "t_0": st.scalar(), "omega": v3(),
                        "rho": [st.scalar() for _ in range(4)]},
           "product": {"s": st.scalar(),
                       "mh": {"s_mid": st.scalar(),
                              "zero": {"a_0": v3(), "b_m": v3(), "r_0": st.scalar(), "s_m": st.scalar(),
                                       "t": [st.scalar() for _ in range(7)]}},
                       "svp": (v3(), st.scalar(), [st.scalar()], st.scalar(), st.scalar())},
           "ddh_r": st.scalar(), "mexp_pk": mexp(), "mexp_comm": mexp()}
    tr = F.new_transcript(b"ShuffleProof", b"Shuffle")
    proof, statement = F.shuffle_prove(tr, sh, xpc, rnd)
    return sh["inputs"], sh["outputs"], proof, statement
