"""TEST INFRASTRUCTURE (oracle): Bulletproofs aggregated range proof, prover and verifier, restated for the parity tests.

The reference calls the `bulletproofs` crate (Cargo.toml:50-53, git branch `develop`, NOT vendored under /root/reference):
  prover    RangeProof::prove_multiple / prove_single     reference src/accounts/prover.rs:544-590
  verifier  RangeProof::verify_multiple / verify_single   reference src/accounts/verifier.rs:504-555
on the running transcript after `domain_sep(b"AggregateBulletProof")`, with PedersenGens::default() (B, B_blinding =
hash-to-group of SHA3-512(enc(B)) = the reference's BASE_PK_BTC_COMPRESSED[1]) and BulletproofGens::new(64, 16 | 1).

What is restated here is the crate's published algorithm (Bulletproofs paper section 4.3 + the crate's transcript protocol:
"rangeproof v1" / n / m, V, A, S -> y, z, T_1, T_2 -> x, t_x, t_x_blinding, e_blinding -> w, "ipp v1" / n, (L, R -> u)*;
proof bytes = A | S | T_1 | T_2 | t_x | t_x_blinding | e_blinding | L_0 | R_0 | ... | a | b).
PARITY UNPINNED for the transcript labels and the wire layout: the crate's own golden proofs are not available here; what
is pinned is the algebra (prover and verifier are independent derivations that must agree) and Merlin (conformance vector).

Scalars are Python ints mod l; points are 32-byte encodings.  The aggregated proof is computed directly (one prover holding
all m values), which gives the same proof as the crate's dealer / party protocol with the same blinding scalars.
Never imported by the product."""
import numpy as np

import c_oracle as C
import ristretto_ref as R
from merlin_ref import Transcript

L = R.L


def sb(k):
    return (k % L).to_bytes(32, "little")


def inv(x):
    return pow(x % L, L - 2, L)


def _msm(scalars, points):
    """sum s_i * P_i over encodings -> (32-byte encoding, status), through the C restatement (fast)."""
    if not scalars:
        return bytes(32), 0
    out, st = C.msm(np.frombuffer(b"".join(sb(s) for s in scalars), np.uint8), np.frombuffer(b"".join(points), np.uint8))
    return out.tobytes(), st


class PedersenGens:
    """PedersenGens::default(): B = basepoint, B_blinding = hash_from_bytes::<Sha3_512>(enc(B))."""

    def __init__(self):
        self.B = R.BASEPOINT_COMPRESSED
        self.B_blinding = R.compress(R.hash_to_point_sha3_512(R.BASEPOINT_COMPRESSED))

    def commit(self, value, blinding):
        return _msm([value, blinding], [self.B, self.B_blinding])[0]


class BulletproofGens:
    """BulletproofGens::new(gens_capacity, party_capacity); G(n, m) / H(n, m) = the first n generators of the first m parties."""

    _cache = {}

    def __init__(self, gens_capacity, party_capacity):
        key = (gens_capacity, party_capacity)
        if key not in BulletproofGens._cache:
            BulletproofGens._cache[key] = R.bulletproof_gens(gens_capacity, party_capacity)
        self.g_rows, self.h_rows = BulletproofGens._cache[key]
        self.gens_capacity, self.party_capacity = gens_capacity, party_capacity

    def G(self, n, m):
        return [p for row in self.g_rows[:m] for p in row[:n]]

    def H(self, n, m):
        return [p for row in self.h_rows[:m] for p in row[:n]]


# ---- the crate's TranscriptProtocol (transcript.rs [upstream]) ---------------------------------------------------------
def rangeproof_domain_sep(tr, n, m):
    tr.append_message(b"dom-sep", b"rangeproof v1")
    tr.append_message(b"n", n.to_bytes(8, "little"))
    tr.append_message(b"m", m.to_bytes(8, "little"))


def innerproduct_domain_sep(tr, n):
    tr.append_message(b"dom-sep", b"ipp v1")
    tr.append_message(b"n", n.to_bytes(8, "little"))


def challenge_scalar(tr, label):
    return int.from_bytes(tr.challenge_bytes(label, 64), "little") % L


def validate_and_append_point(tr, label, point):
    if point == bytes(32):          # the identity's encoding
        return False
    tr.append_message(label, point)
    return True


def sum_of_powers(x, n):
    return sum(pow(x, i, L) for i in range(n)) % L


def delta(n, m, y, z):
    return ((z - z * z) * sum_of_powers(y, n * m) - z * z * z * sum_of_powers(2, n) * sum_of_powers(z, m)) % L


# ---- prover ------------------------------------------------------------------------------------------------------------
def prove_multiple(tr, values, blindings, n, rnd, bp_gens=None, pc_gens=None):
    """RangeProof::prove_multiple(bp_gens, pc_gens, transcript, values, blindings, n) -> (proof bytes, [V_j]).
    rnd() supplies the blinding scalars the crate draws from its transcript RNG."""
    m = len(values)
    assert m == len(blindings) and m & (m - 1) == 0 and n in (8, 16, 32, 64)
    pc = pc_gens or PedersenGens()
    bp = bp_gens or BulletproofGens(64, max(m, 1))
    N = n * m
    G, H = bp.G(n, m), bp.H(n, m)
    V = [pc.commit(v, g) for v, g in zip(values, blindings)]
    rangeproof_domain_sep(tr, n, m)
    for v in V:
        tr.append_message(b"V", v)
    aL = [(values[j] >> i) & 1 for j in range(m) for i in range(n)]
    aR = [(b - 1) % L for b in aL]
    alpha, rho = rnd(), rnd()
    sL, sR = [rnd() for _ in range(N)], [rnd() for _ in range(N)]
    A = _msm([alpha] + aL + aR, [pc.B_blinding] + G + H)[0]
    S = _msm([rho] + sL + sR, [pc.B_blinding] + G + H)[0]
    tr.append_message(b"A", A)
    tr.append_message(b"S", S)
    y, z = challenge_scalar(tr, b"y"), challenge_scalar(tr, b"z")
    yp = [pow(y, i, L) for i in range(N)]
    zz = z * z % L
    # l(X) = (aL - z) + sL X ;  r(X) = y^i (aR + z + sR X) + z^2 z^j 2^(i mod n)
    l0 = [(a - z) % L for a in aL]
    l1 = sL
    r0 = [(yp[i] * (aR[i] + z) + zz * pow(z, i // n, L) * pow(2, i % n, L)) % L for i in range(N)]
    r1 = [yp[i] * sR[i] % L for i in range(N)]
    t1 = sum(a * d + b * c for a, b, c, d in zip(l0, l1, r0, r1)) % L
    t2 = sum(b * d for b, d in zip(l1, r1)) % L
    tau1, tau2 = rnd(), rnd()
    T1, T2 = pc.commit(t1, tau1), pc.commit(t2, tau2)
    tr.append_message(b"T_1", T1)
    tr.append_message(b"T_2", T2)
    x = challenge_scalar(tr, b"x")
    lv = [(a + b * x) % L for a, b in zip(l0, l1)]
    rv = [(c + d * x) % L for c, d in zip(r0, r1)]
    t_x = sum(a * b for a, b in zip(lv, rv)) % L
    t_x_blinding = (tau1 * x + tau2 * x * x + sum(zz * pow(z, j, L) * blindings[j] for j in range(m))) % L
    e_blinding = (alpha + rho * x) % L
    tr.append_message(b"t_x", sb(t_x))
    tr.append_message(b"t_x_blinding", sb(t_x_blinding))
    tr.append_message(b"e_blinding", sb(e_blinding))
    w = challenge_scalar(tr, b"w")
    Q = _msm([w], [pc.B])[0]
    yinv = inv(y)
    ipp = inner_product_create(tr, Q, [1] * N, [pow(yinv, i, L) for i in range(N)], G, H, lv, rv)
    return A + S + T1 + T2 + sb(t_x) + sb(t_x_blinding) + sb(e_blinding) + ipp, V


def inner_product_create(tr, Q, g_factors, h_factors, G, H, a, b):
    """InnerProductProof::create: the folded generators are kept as coefficient vectors over the original G, H."""
    N = len(a)
    innerproduct_domain_sep(tr, N)
    gs, hs = list(g_factors), list(h_factors)
    a, b = list(a), list(b)
    out_lr = b""
    cur = N
    while cur != 1:
        n = cur // 2
        aL, aR, bL, bR = a[:n], a[n:], b[:n], b[n:]
        cL = sum(p * q for p, q in zip(aL, bR)) % L
        cR = sum(p * q for p, q in zip(aR, bL)) % L
        Ls, Lp, Rs, Rp = [cL], [Q], [cR], [Q]
        for i in range(N):
            s = i % cur
            if s >= n:      # G_R, H_R halves
                Ls.append(aL[s - n] * gs[i])
                Lp.append(G[i])
                Rs.append(bL[s - n] * hs[i])
                Rp.append(H[i])
            else:           # G_L, H_L halves
                Rs.append(aR[s] * gs[i])
                Rp.append(G[i])
                Ls.append(bR[s] * hs[i])
                Lp.append(H[i])
        Lpt, Rpt = _msm(Ls, Lp)[0], _msm(Rs, Rp)[0]
        out_lr += Lpt + Rpt
        tr.append_message(b"L", Lpt)
        tr.append_message(b"R", Rpt)
        u = challenge_scalar(tr, b"u")
        ui = inv(u)
        a = [(aL[k] * u + ui * aR[k]) % L for k in range(n)]
        b = [(bL[k] * ui + u * bR[k]) % L for k in range(n)]
        for i in range(N):
            if i % cur < n:
                gs[i] = gs[i] * ui % L
                hs[i] = hs[i] * u % L
            else:
                gs[i] = gs[i] * u % L
                hs[i] = hs[i] * ui % L
        cur = n
    return out_lr + sb(a[0]) + sb(b[0])


# ---- verifier ----------------------------------------------------------------------------------------------------------
def parse_proof(proof):
    """RangeProof::from_bytes: None on a format error (length, non-canonical scalar)."""
    if len(proof) % 32 or len(proof) < 7 * 32:
        return None
    w = [proof[32 * i:32 * i + 32] for i in range(len(proof) // 32)]
    rest = w[7:]
    if len(rest) < 2 or (len(rest) - 2) % 2 or (len(rest) - 2) // 2 >= 32:
        return None
    for s in (w[4], w[5], w[6], rest[-2], rest[-1]):
        if not R.scalar_is_canonical(s):
            return None
    lg = (len(rest) - 2) // 2
    sc = lambda b: int.from_bytes(b, "little")  # noqa: E731
    return dict(A=w[0], S=w[1], T1=w[2], T2=w[3], t_x=sc(w[4]), t_x_blinding=sc(w[5]), e_blinding=sc(w[6]),
                Lv=[rest[2 * k] for k in range(lg)], Rv=[rest[2 * k + 1] for k in range(lg)], a=sc(rest[-2]), b=sc(rest[-1]))


def verification_terms(tr, proof, commitments, n, c, bp_gens=None, pc_gens=None):
    """The scalars and points of verify_multiple's mega-check for the random weight c, or None when verification stops
    before it (format, sizes, identity points)."""
    m = len(commitments)
    pr = parse_proof(proof)
    if pr is None or n not in (8, 16, 32, 64):
        return None
    bp = bp_gens or BulletproofGens(64, max(m, 1))
    pc = pc_gens or PedersenGens()
    if bp.gens_capacity < n or bp.party_capacity < m:
        return None
    rangeproof_domain_sep(tr, n, m)
    for v in commitments:
        tr.append_message(b"V", v)
    if not validate_and_append_point(tr, b"A", pr["A"]) or not validate_and_append_point(tr, b"S", pr["S"]):
        return None
    y, z = challenge_scalar(tr, b"y"), challenge_scalar(tr, b"z")
    if not validate_and_append_point(tr, b"T_1", pr["T1"]) or not validate_and_append_point(tr, b"T_2", pr["T2"]):
        return None
    x = challenge_scalar(tr, b"x")
    tr.append_message(b"t_x", sb(pr["t_x"]))
    tr.append_message(b"t_x_blinding", sb(pr["t_x_blinding"]))
    tr.append_message(b"e_blinding", sb(pr["e_blinding"]))
    w = challenge_scalar(tr, b"w")
    # InnerProductProof::verification_scalars(n * m, transcript)
    N = n * m
    lg = len(pr["Lv"])
    if lg >= 32 or N != (1 << lg):
        return None
    innerproduct_domain_sep(tr, N)
    us = []
    for Lp, Rp in zip(pr["Lv"], pr["Rv"]):
        if not validate_and_append_point(tr, b"L", Lp) or not validate_and_append_point(tr, b"R", Rp):
            return None
        us.append(challenge_scalar(tr, b"u"))
    if any(u == 0 for u in us) or y == 0:
        return None         # (probability 2^-252; Scalar::invert of zero)
    allinv = 1
    for u in us:
        allinv = allinv * inv(u) % L
    u_sq = [u * u % L for u in us]
    u_inv_sq = [inv(u) ** 2 % L for u in us]
    s = [allinv]
    for i in range(1, N):
        lg_i = i.bit_length() - 1
        s.append(s[i - (1 << lg_i)] * u_sq[(lg - 1) - lg_i] % L)
    a, b = pr["a"], pr["b"]
    zz = z * z % L
    yinv = inv(y)
    g = [(-z - a * s[i]) % L for i in range(N)]
    h = [(z + pow(yinv, i, L) * (zz * pow(z, i // n, L) * pow(2, i % n, L) - b * s[N - 1 - i])) % L for i in range(N)]
    vcs = [c * zz * pow(z, j, L) % L for j in range(m)]
    basepoint_scalar = (w * (pr["t_x"] - a * b) + c * (delta(n, m, y, z) - pr["t_x"])) % L
    scalars = [1, x, c * x % L, c * x * x % L] + u_sq + u_inv_sq + [(-pr["e_blinding"] - c * pr["t_x_blinding"]) % L,
                                                                   basepoint_scalar] + g + h + vcs
    points = [pr["A"], pr["S"], pr["T1"], pr["T2"]] + pr["Lv"] + pr["Rv"] + [pc.B_blinding, pc.B] + bp.G(n, m) + bp.H(n, m) + list(commitments)
    return scalars, points


def verify_multiple(tr, proof, commitments, n, c=0x1234567, bp_gens=None, pc_gens=None):
    """RangeProof::verify_multiple -> True (Ok) / False (Err).  c stands for the crate's random Scalar."""
    terms = verification_terms(tr, proof, commitments, n, c, bp_gens, pc_gens)
    if terms is None:
        return False
    out, st = _msm(*terms)
    return st == 0 and out == bytes(32)


def verify_single(tr, proof, commitment, n, c=0x1234567, bp_gens=None, pc_gens=None):
    return verify_multiple(tr, proof, [commitment], n, c, bp_gens, pc_gens)


# ---- the reference's call sites ----------------------------------------------------------------------------------------
def quisquis_range_prover(tr, balances, rscalars, rnd):
    """Prover::verify_non_negative_sender_receiver_prover (reference src/accounts/prover.rs:544-590) -> [proof bytes]."""
    tr.domain_sep(b"AggregateBulletProof")
    size = len(balances)
    if size & (size - 1) == 0:
        return [prove_multiple(tr, balances, rscalars, 64, rnd, BulletproofGens(64, 16))[0]]
    return [prove_multiple(tr, [v], [r], 64, rnd, BulletproofGens(64, 1))[0] for v, r in zip(balances, rscalars)]


def quisquis_range_batch_verifier(tr, epsilon_accounts, proof, c=0x1234567):
    """Verifier::verify_non_negative_sender_receiver_bulletproof_batch_verifier (reference src/accounts/verifier.rs:504-523)."""
    tr.domain_sep(b"AggregateBulletProof")
    return verify_multiple(tr, proof, [acc[96:128] for acc in epsilon_accounts], 64, c, BulletproofGens(64, 16))


def quisquis_range_vector_verifier(tr, epsilon_accounts, proofs, c=0x1234567):
    """Verifier::verify_non_negative_sender_receiver_bulletproof_vector_verifier (reference src/accounts/verifier.rs:534-555)."""
    tr.domain_sep(b"AggregateBulletProof")
    for proof, acc in zip(proofs, epsilon_accounts):
        if not verify_single(tr, proof, acc[96:128], 64, c, BulletproofGens(64, 1)):
            return False
    return True
