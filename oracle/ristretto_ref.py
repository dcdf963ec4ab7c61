"""CPU ORACLE (test infrastructure, NOT a product path) -- big-int restatement of Ristretto255.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

The reference (twilight-project/quisquis-rust) performs every group operation through the third-party crate
curve25519-dalek = "3" (reference Cargo.toml:42, docs pin 3.2.1 -- RELEASE_NOTES.md:24); that crate is NOT
vendored under /root/reference, so this file restates the *published* algorithm it implements: RFC 9496
(ristretto255) over Edwards25519, a = -1, d = -121665/121666.  Canonical 32-byte encodings make any correct
implementation byte-identical to dalek.  Parity is pinned against: (1) RFC 9496 Appendix A vectors
(tests/golden/rfc9496.json), (2) libsodium 1.0.20 (independent implementation, via ctypes) and (3) the
reference's only fixed point bytes, BASE_PK_BTC_COMPRESSED (reference src/ristretto/constants.rs:12-21).

Reference call sites restated at the bottom of this file:
  RistrettoPublicKey::update_public_key     src/ristretto/keys.rs:146-148, Mul :266-282
  RistrettoPublicKey::verify_keypair        src/ristretto/keys.rs:187-195
  verify_public_key_update                  src/ristretto/keys.rs:161-169
  ElGamalCommitment::generate_commitment    src/elgamal/elgamal.rs:41-53
  ElGamalCommitment::add_commitments        src/elgamal/elgamal.rs:65-69
  ElGamalCommitment::verify_commitment      src/elgamal/elgamal.rs:81-95
  Account::update_account                   src/accounts/accounts.rs:143-154
  Account::verify_account                   src/accounts/accounts.rs:81-84
  Account::create_delta_and_epsilon_accounts src/accounts/accounts.rs:198-220
  Verifier::multiscalar_multiplication      src/accounts/verifier.rs:91-99
  Verifier::verify_delta_identity_check     src/accounts/verifier.rs:566-581
"""
import hashlib

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493
D = (-121665 * pow(121666, P - 2, P)) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)


def _is_neg(x):
    return (x % P) & 1


def _abs(x):
    x %= P
    return P - x if x & 1 else x


def sqrt_ratio_i(u, v):
    """RFC 9496 4.2 SQRT_RATIO_M1 (dalek field.rs sqrt_ratio_i)."""
    u %= P
    v %= P
    v3 = v * v % P * v % P
    v7 = v3 * v3 % P * v % P
    r = u * v3 % P * pow(u * v7 % P, (P - 5) // 8, P) % P
    check = v * r % P * r % P
    correct = check == u
    flipped = check == (-u) % P
    flipped_i = check == (-u * SQRT_M1) % P
    if flipped or flipped_i:
        r = r * SQRT_M1 % P
    r = _abs(r)
    return (correct or flipped), r


# derived constants (RFC 9496 4.1); the specific roots are fixed by checking against the spec values
INVSQRT_A_MINUS_D = sqrt_ratio_i(1, (-1 - D) % P)[1]
SQRT_AD_MINUS_ONE = 25063068953384623474111414158702152701244531502492656460079210482610430750235
assert SQRT_AD_MINUS_ONE * SQRT_AD_MINUS_ONE % P == (-D - 1) % P
assert INVSQRT_A_MINUS_D == 54469307008909316920995813868745141605393597292927456921205312896311721017578
ONE_MINUS_D_SQ = (1 - D * D) % P
D_MINUS_ONE_SQ = (D - 1) * (D - 1) % P

IDENTITY = (0, 1, 1, 0)


def decompress(b):
    """RFC 9496 4.3.1 Decode. Returns extended point (X,Y,Z,T) or None."""
    if len(b) != 32:
        return None
    s = int.from_bytes(b, "little")
    if s >= P or (s & 1):  # non-canonical (incl. bit 255 set) or negative
        return None
    ss = s * s % P
    u1 = (1 - ss) % P
    u2 = (1 + ss) % P
    u2s = u2 * u2 % P
    v = (-(D * u1 % P * u1) - u2s) % P
    ok, inv = sqrt_ratio_i(1, v * u2s % P)
    dx = inv * u2 % P
    dy = inv * dx % P * v % P
    x = _abs(2 * s * dx % P)
    y = u1 * dy % P
    t = x * y % P
    if (not ok) or _is_neg(t) or y == 0:
        return None
    return (x, y, 1, t)


def compress(pt):
    """RFC 9496 4.3.2 Encode."""
    X, Y, Z, T = pt
    u1 = (Z + Y) * (Z - Y) % P
    u2 = X * Y % P
    _, inv = sqrt_ratio_i(1, u1 * u2 % P * u2 % P)
    i1 = inv * u1 % P
    i2 = inv * u2 % P
    z_inv = i1 * i2 % P * T % P
    den_inv = i2
    if _is_neg(T * z_inv % P):
        X, Y = Y * SQRT_M1 % P, X * SQRT_M1 % P
        den_inv = i1 * INVSQRT_A_MINUS_D % P
    if _is_neg(X * z_inv % P):
        Y = (-Y) % P
    s = _abs(den_inv * ((Z - Y) % P) % P)
    return s.to_bytes(32, "little")


def add(p, q):
    X1, Y1, Z1, T1 = p
    X2, Y2, Z2, T2 = q
    A = (Y1 - X1) * (Y2 - X2) % P
    B = (Y1 + X1) * (Y2 + X2) % P
    C = 2 * D * T1 % P * T2 % P
    Dd = 2 * Z1 * Z2 % P
    E, F, G, H = (B - A) % P, (Dd - C) % P, (Dd + C) % P, (B + A) % P
    return (E * F % P, G * H % P, F * G % P, E * H % P)


def neg(p):
    X, Y, Z, T = p
    return ((-X) % P, Y, Z, (-T) % P)


def sub(p, q):
    return add(p, neg(q))


def double(p):
    return add(p, p)


def mul(k, p):
    k %= L
    r = IDENTITY
    for i in reversed(range(k.bit_length())):
        r = double(r)
        if (k >> i) & 1:
            r = add(r, p)
    return r


def eq(p, q):
    X1, Y1, _, _ = p
    X2, Y2, _, _ = q
    return (X1 * Y2 - Y1 * X2) % P == 0 or (X1 * X2 - Y1 * Y2) % P == 0


def is_identity(p):
    return eq(p, IDENTITY)


# Ed25519 basepoint
_By = 4 * pow(5, P - 2, P) % P
_Bx2 = (_By * _By - 1) * pow(D * _By * _By + 1, P - 2, P) % P
_Bx = pow(_Bx2, (P + 3) // 8, P)
if (_Bx * _Bx - _Bx2) % P != 0:
    _Bx = _Bx * SQRT_M1 % P
if _Bx & 1:
    _Bx = P - _Bx
BASEPOINT = (_Bx, _By, 1, _Bx * _By % P)
BASEPOINT_COMPRESSED = compress(BASEPOINT)


def elligator(r0):
    """RFC 9496 4.3.4 MAP (dalek elligator_ristretto_flavor)."""
    r = SQRT_M1 * r0 % P * r0 % P
    ns = (r + 1) * ONE_MINUS_D_SQ % P
    c = P - 1
    Dd = (c - D * r) % P * ((r + D) % P) % P
    sq, s = sqrt_ratio_i(ns, Dd)
    s_prime = (-_abs(s * r0 % P)) % P
    if not sq:
        s = s_prime
        c = r
    nt = (c * ((r - 1) % P) % P * D_MINUS_ONE_SQ - Dd) % P
    w0 = 2 * s * Dd % P
    w1 = nt * SQRT_AD_MINUS_ONE % P
    w2 = (1 - s * s) % P
    w3 = (1 + s * s) % P
    return (w0 * w3 % P, w2 * w1 % P, w1 * w3 % P, w0 * w2 % P)


def from_uniform_bytes(b):
    assert len(b) == 64
    r0 = int.from_bytes(b[:32], "little") & (2**255 - 1)
    r1 = int.from_bytes(b[32:], "little") & (2**255 - 1)
    return add(elligator(r0 % P), elligator(r1 % P))


# Pedersen H = bulletproofs PedersenGens::default().B_blinding = from_uniform_bytes(SHA3-512(B_compressed));
# equals BASE_PK_BTC_COMPRESSED[1] (reference src/ristretto/constants.rs:17-20, src/pedersen/vectorpedersen.rs:49-51)
PEDERSEN_H = from_uniform_bytes(hashlib.sha3_512(BASEPOINT_COMPRESSED).digest())
PEDERSEN_H_COMPRESSED = compress(PEDERSEN_H)
BASE_PK = BASEPOINT_COMPRESSED + PEDERSEN_H_COMPRESSED


def scalar_from_bytes(b):
    return int.from_bytes(b, "little")


def scalar_to_bytes(k):
    return (k % L).to_bytes(32, "little")


def scalar_is_canonical(b):
    return int.from_bytes(b, "little") < L


# ----------------------------------------------------------------------------------------------------------------
# Reference call sites, byte-level (inputs/outputs are the reference's serialized forms).
# Status codes mirror include/qq_b200.h: 0 ok, 1 invalid point encoding, 2 non-canonical scalar,
# 3 keypair mismatch, 4 commitment mismatch.
# ----------------------------------------------------------------------------------------------------------------
ST_OK, ST_BAD_POINT, ST_BAD_SCALAR, ST_KEYPAIR, ST_COMMIT = 0, 1, 2, 3, 4
ZERO32 = bytes(32)


def _scal(b):
    if not scalar_is_canonical(b):
        return None
    return int.from_bytes(b, "little")


def update_public_key(pk, r):
    """src/ristretto/keys.rs:146-148 -> (r*gr, r*grsk). Returns (64 bytes, status)."""
    k = _scal(r)
    if k is None:
        return bytes(64), ST_BAD_SCALAR
    gr, grsk = decompress(pk[:32]), decompress(pk[32:64])
    if gr is None or grsk is None:
        return bytes(64), ST_BAD_POINT
    return compress(mul(k, gr)) + compress(mul(k, grsk)), ST_OK


def generate_commitment(pk, r, v):
    """src/elgamal/elgamal.rs:41-53 -> c = r*gr, d = v*B + r*grsk."""
    kr, kv = _scal(r), _scal(v)
    if kr is None or kv is None:
        return bytes(64), ST_BAD_SCALAR
    gr, grsk = decompress(pk[:32]), decompress(pk[32:64])
    if gr is None or grsk is None:
        return bytes(64), ST_BAD_POINT
    c = mul(kr, gr)
    d = add(mul(kv, BASEPOINT), mul(kr, grsk))
    return compress(c) + compress(d), ST_OK


def add_commitments(a, b):
    """src/elgamal/elgamal.rs:65-69."""
    pts = [decompress(x[i:i + 32]) for x in (a, b) for i in (0, 32)]
    if any(p is None for p in pts):
        return bytes(64), ST_BAD_POINT
    return compress(add(pts[0], pts[2])) + compress(add(pts[1], pts[3])), ST_OK


def sub_commitments(a, b):
    """src/elgamal/elgamal.rs:201-218 (impl Sub)."""
    pts = [decompress(x[i:i + 32]) for x in (a, b) for i in (0, 32)]
    if any(p is None for p in pts):
        return bytes(64), ST_BAD_POINT
    return compress(sub(pts[0], pts[2])) + compress(sub(pts[1], pts[3])), ST_OK


def mul_commitment(a, s):
    """src/elgamal/elgamal.rs:220-236 (impl Mul<&Scalar>)."""
    return update_public_key(a, s)  # same arithmetic: scalar times both points


def update_account(acc, bl, u, c):
    """src/accounts/accounts.rs:143-154: pk' = u*pk ; comm' = generate_commitment(OLD pk, c, bl) + comm."""
    ks = [_scal(x) for x in (bl, u, c)]
    if any(k is None for k in ks):
        return bytes(128), ST_BAD_SCALAR
    pts = [decompress(acc[i:i + 32]) for i in (0, 32, 64, 96)]
    if any(p is None for p in pts):
        return bytes(128), ST_BAD_POINT
    kbl, ku, kc = ks
    gr, grsk, cc, dd = pts
    out = compress(mul(ku, gr)) + compress(mul(ku, grsk))
    out += compress(add(mul(kc, gr), cc))
    out += compress(add(add(mul(kbl, BASEPOINT), mul(kc, grsk)), dd))
    return out, ST_OK


def verify_account(acc, sk, bl):
    """src/accounts/accounts.rs:81-84 -> status (keypair check first, then commitment)."""
    ksk, kbl = _scal(sk), _scal(bl)
    if ksk is None or kbl is None:
        return ST_BAD_SCALAR
    gr = decompress(acc[0:32])
    if gr is None:
        return ST_BAD_POINT
    if compress(mul(ksk, gr)) != acc[32:64]:
        return ST_KEYPAIR
    c = decompress(acc[64:96])
    if c is None:
        return ST_BAD_POINT
    if compress(add(mul(kbl, BASEPOINT), mul(ksk, c))) != acc[96:128]:
        return ST_COMMIT
    return ST_OK


def verify_public_key_update(upd, pk, r):
    """src/ristretto/keys.rs:161-169 -> status 0 (true) / 3 (false) / 1 (reference panics)."""
    k = _scal(r)
    if k is None:
        return ST_BAD_SCALAR
    pts = [decompress(x[i:i + 32]) for x in (pk, upd) for i in (0, 32)]
    if any(p is None for p in pts):
        return ST_BAD_POINT
    ok = eq(mul(k, pts[0]), pts[2]) and eq(mul(k, pts[1]), pts[3])
    return ST_OK if ok else ST_KEYPAIR


def delta_epsilon(acc, bl, r, base_pk=None):
    """src/accounts/accounts.rs:198-220, per account, with the random scalar r supplied by the caller.
    delta = (pk, generate_commitment(pk, r, bl)); epsilon = (base_pk, generate_commitment(base_pk, r, bl))."""
    base_pk = BASE_PK if base_pk is None else base_pk
    cd, st = generate_commitment(acc[:64], r, bl)
    if st:
        return bytes(128), bytes(128), st
    ce, st = generate_commitment(base_pk, r, bl)
    if st:
        return bytes(128), bytes(128), st
    return acc[:64] + cd, base_pk + ce, ST_OK


def fixed_base(which, s):
    """s*B (which=0) or s*H (which=1), compressed."""
    k = _scal(s)
    if k is None:
        return ZERO32, ST_BAD_SCALAR
    return compress(mul(k, BASEPOINT if which == 0 else PEDERSEN_H)), ST_OK


def msm(scalars, points):
    """src/accounts/verifier.rs:91-99 optional_multiscalar_mul: (32 bytes, status)."""
    acc = IDENTITY
    for s, p in zip(scalars, points):
        k = _scal(s)
        if k is None:
            return ZERO32, ST_BAD_SCALAR
        q = decompress(p)
        if q is None:
            return ZERO32, ST_BAD_POINT
        acc = add(acc, mul(k, q))
    return compress(acc), ST_OK


def delta_identity_check(accounts):
    """src/accounts/verifier.rs:566-581 -> 0 ok / 4 failed / 1 bad point."""
    sc, sd = IDENTITY, IDENTITY
    for a in accounts:
        c, d = decompress(a[64:96]), decompress(a[96:128])
        if c is None or d is None:
            return ST_BAD_POINT
        sc, sd = add(sc, c), add(sd, d)
    return ST_OK if is_identity(sc) and is_identity(sd) else ST_COMMIT


# ---- generator derivation (reference src/pedersen/vectorpedersen.rs:45-75; bulletproofs generators.rs [upstream]) ----
def hash_to_point_sha3_512(data):
    """RistrettoPoint::hash_from_bytes::<Sha3_512>."""
    return from_uniform_bytes(hashlib.sha3_512(data).digest())


def vector_pedersen_gens(capacity):
    """VectorPedersenGens::new(capacity) -> (H compressed, [G_vec compressed]); len(G_vec) = capacity - 1."""
    h = hash_to_point_sha3_512(BASEPOINT_COMPRESSED)
    others = [h]
    for i in range(capacity - 2):
        others.append(hash_to_point_sha3_512(compress(others[i])))
    return compress(h), [BASEPOINT_COMPRESSED] + [compress(p) for p in others[1:]]


def bulletproof_gens(gens_capacity, party_capacity):
    """BulletproofGens::new: per party i the chains Shake256("GeneratorsChain" || b"G"/b"H" || LE32(i)), 64 bytes per point."""
    out = {}
    for tag in (b"G", b"H"):
        rows = []
        for i in range(party_capacity):
            stream = hashlib.shake_256(b"GeneratorsChain" + tag + i.to_bytes(4, "little")).digest(64 * gens_capacity)
            rows.append([compress(from_uniform_bytes(stream[64 * j:64 * j + 64])) for j in range(gens_capacity)])
        out[tag] = rows
    return out[b"G"], out[b"H"]


# ---- decommit (reference src/elgamal/elgamal.rs:106-108) ----
def decommit(comm, sk):
    """enc(d - sk*c); (bytes, status) with status 1 on a bad point, 2 on a non-canonical scalar."""
    if not scalar_is_canonical(sk):
        return bytes(32), 2
    c, d = decompress(comm[:32]), decompress(comm[32:64])
    if c is None or d is None:
        return bytes(32), 1
    return compress(sub(d, mul(_scal(sk), c))), 0
