"""TEST INFRASTRUCTURE (oracle): the leaf arguments of the reference's Bayer-Groth shuffle proof, restated for the parity
tests - prover and verifier of each, following the reference line by line:
  DDH tuple proof            src/shuffle/ddh.rs:50-142
  single value product (SVP) src/shuffle/singlevalueproduct.rs:61-257
  Hadamard product           src/shuffle/hadamard.rs:96-389, polynomials src/shuffle/polynomial.rs:350-515
Scalars are Python ints mod l, points are 32-byte encodings.  Blinding scalars are arguments (the reference draws them
from a transcript RNG; any value gives a valid proof).  Never imported by the product."""
import ristretto_ref as R
from merlin_ref import Transcript

L = R.L
ROWS = COLUMNS = 3


def sb(k):
    return (k % L).to_bytes(32, "little")


def exp_iter(x, n, skip=0):
    """vectorutil::exp_iter(x).skip(skip).take(n): 1, x, x^2, ..."""
    out, cur = [], 1
    for i in range(skip + n):
        if i >= skip:
            out.append(cur)
        cur = cur * x % L
    return out


class XpcGens:
    """VectorPedersenGens::new(capacity): commit(values, blinding) = blinding * H + sum values_i * G_i."""

    def __init__(self, capacity):
        self.h, self.g = R.vector_pedersen_gens(capacity)

    def commit_point(self, values, blinding):
        acc = R.mul(blinding % L, R.decompress(self.h))
        for v, g in zip(values, self.g):
            acc = R.add(acc, R.mul(v % L, R.decompress(g)))
        return acc

    def commit(self, values, blinding):
        return R.compress(self.commit_point(values, blinding))


def _msm_point(scalars, points):
    """sum s_i * P_i as a point, or None when a point does not decode."""
    acc = R.IDENTITY if hasattr(R, "IDENTITY") else R.mul(0, R.BASEPOINT)
    for s, p in zip(scalars, points):
        q = R.decompress(p)
        if q is None:
            return None
        acc = R.add(acc, R.mul(s % L, q))
    return acc


def new_transcript(transcript_label, side_label):
    tr = Transcript(transcript_label)
    tr.domain_sep(side_label)          # Prover::new / Verifier::new
    return tr


# ---- DDH tuple proof ---------------------------------------------------------------------------------------------------
def ddh_prove(tr, g_i, h_i, exp_x, G, H, rho, r_scalar):
    """-> ((challenge, z), (G_dash, H_dash)).  g_i, h_i, G, H: compressed points."""
    tr.domain_sep(b"DDHTupleProof")
    exp_x_rho = [x * rho % L for x in exp_x]
    G_dash = R.compress(_msm_point(exp_x_rho, g_i))
    H_dash = R.compress(_msm_point(exp_x_rho, h_i))
    g_r = R.compress(R.mul(r_scalar % L, R.decompress(G)))
    h_r = R.compress(R.mul(r_scalar % L, R.decompress(H)))
    for label, p in ((b"g", G), (b"g_dash", G_dash), (b"h", H), (b"h_dash", H_dash), (b"gr", g_r), (b"hr", h_r)):
        tr.append_point_var(label, p)
    challenge = tr.get_challenge(b"Challenge")
    return (challenge, (r_scalar - challenge * rho) % L), (G_dash, H_dash)


def ddh_verify(tr, proof, statement, G, H):
    """-> True, False or None (= Err on an undecodable point; the reference gives the same message for both)."""
    challenge, z = proof
    G_dash, H_dash = statement
    tr.domain_sep(b"DDHTupleProof")
    for label, p in ((b"g", G), (b"g_dash", G_dash), (b"h", H), (b"h_dash", H_dash)):
        tr.append_point_var(label, p)
    g_r = _msm_point([z, challenge], [G, G_dash])
    if g_r is None:
        return None
    h_r = _msm_point([z, challenge], [H, H_dash])
    if h_r is None:
        return None
    tr.append_point_var(b"gr", R.compress(g_r))
    tr.append_point_var(b"hr", R.compress(h_r))
    return tr.get_challenge(b"Challenge") == challenge % L


# ---- single value product argument ---------------------------------------------------------------------------------------
def svp_prove(tr, xpc, r, a_vec, d_vec, rd, delta_mid, s_1, s_x):
    """-> proof dict.  a_vec: COLUMNS scalars committed in c_a = xpc.commit(a_vec, r); d_vec, rd, delta_mid (the random
    middle entries of delta_vec, COLUMNS - 2 of them), s_1, s_x: the prover's randomness."""
    tr.domain_sep(b"SingleValueProductProof")
    bvec, prod = [], 1
    for a in a_vec:
        prod = prod * a % L
        bvec.append(prod)
    commit_d = xpc.commit(d_vec, rd)
    delta_vec = [d_vec[0]] + list(delta_mid) + [0]
    assert len(delta_vec) == COLUMNS
    delta_lower = [(-delta_vec[i]) * d_vec[i + 1] % L for i in range(COLUMNS - 1)]
    delta_upper = [(delta_vec[i + 1] - a_vec[i + 1] * delta_vec[i] - bvec[i] * d_vec[i + 1]) % L for i in range(COLUMNS - 1)]
    trun = XpcGens(len(delta_lower) + 1)
    c_small, c_cap = trun.commit(delta_lower, s_1), trun.commit(delta_upper, s_x)
    tr.append_point_var(b"DeltaSmall", c_small)
    tr.append_point_var(b"DeltaCapital", c_cap)
    tr.append_point_var(b"d", commit_d)
    x = tr.get_challenge(b"challenge")
    return {"commitment_d": commit_d, "commitment_delta_small": c_small, "commitment_delta_capital": c_cap,
            "a_twildle": [(a * x + d) % L for a, d in zip(a_vec, d_vec)],
            "b_twildle": [(b * x + d) % L for b, d in zip(bvec, delta_vec)],
            "r_twildle": (r * x + rd) % L, "s_twildle": (s_x * x + s_1) % L}


def svp_verify(tr, proof, commitment_a, b, xpc):
    """-> True, False ("SingleValue Product Proof Verify: Failed"), None (Decompression Failed), "size" (Size check failed)."""
    at, bt = proof["a_twildle"], proof["b_twildle"]
    if len(at) != COLUMNS or len(bt) != COLUMNS:
        return "size"
    if at[0] % L != bt[0] % L:
        return False
    tr.domain_sep(b"SingleValueProductProof")
    tr.append_point_var(b"DeltaSmall", proof["commitment_delta_small"])
    tr.append_point_var(b"DeltaCapital", proof["commitment_delta_capital"])
    tr.append_point_var(b"d", proof["commitment_d"])
    x = tr.get_challenge(b"challenge")
    if b * x % L != bt[COLUMNS - 1] % L:
        return False
    comit_a_bar = xpc.commit_point(at, proof["r_twildle"])
    lhs = _msm_point([x, 1], [commitment_a, proof["commitment_d"]])
    if lhs is None:
        return None
    if not R.eq(lhs, comit_a_bar):
        return False
    lhs2 = _msm_point([x, 1], [proof["commitment_delta_capital"], proof["commitment_delta_small"]])
    if lhs2 is None:
        return None
    comvec = [(bt[i + 1] * x - bt[i] * at[i + 1]) % L for i in range(COLUMNS - 1)]
    trun = XpcGens(len(comvec) + 1)
    return bool(R.eq(lhs2, trun.commit_point(comvec, proof["s_twildle"])))


# ---- polynomials over Z/l (coefficient lists, lowest degree first) ----------------------------------------------------------
def poly_mul(a, b):
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] = (out[i + j] + x * y) % L
    return out


def poly_add(a, b):
    n = max(len(a), len(b))
    return [((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % L for i in range(n)]


def poly_sub(a, b):
    return poly_add(a, [(-x) % L for x in b])


def poly_scale(a, s):
    return [x * s % L for x in a]


def poly_eval(a, x):
    acc = 0
    for c in reversed(a):
        acc = (acc * x + c) % L
    return acc


def poly_divexact(num, den):
    num = list(num)
    out = [0] * (len(num) - len(den) + 1)
    inv_lead = pow(den[-1], -1, L)
    for k in range(len(out) - 1, -1, -1):
        q = num[k + len(den) - 1] * inv_lead % L
        out[k] = q
        for j, d in enumerate(den):
            num[k + j] = (num[k + j] - q * d) % L
    assert all(v == 0 for v in num), "polynomial division left a remainder"
    return out


def l_polys(w):
    """polynomial::create_l_i_x_polynomial: [l(X) = prod (X - w_j), l_1(X), l_2(X), l_3(X)] (Lagrange basis on w)."""
    def lx(ws):
        p = [1]
        for v in ws:
            p = poly_mul(p, [(-v) % L, 1])
        return p
    out = [lx(w)]
    for i in range(3):
        others = [w[j] for j in range(3) if j != i]
        den = 1
        for o in others:
            den = den * (w[i] - o) % L
        out.append(poly_scale(lx(others), pow(den, -1, L)))
    return out


def l_evals(w, x):
    """[l(x), l_1(x), l_2(x), l_3(x)] - what the verifier needs."""
    return [poly_eval(p, x) for p in l_polys(w)]


# ---- Hadamard product argument ------------------------------------------------------------------------------------------
def hadamard_prove(tr, xpc, a, b, c, commit_a, commit_b, commit_c, wr, ws, wt, rnd):
    """a, b, c: ROWS x COLUMNS matrices as lists of rows with c = a o b; commit_x[i] = xpc.commit(x[i], w_x[i]).
    rnd: dict with a_0, b_0 (COLUMNS each), r_0, s_0, t_0, omega (3 distinct), rho (ROWS + 1).
    -> (proof dict, omega)."""
    tr.domain_sep(b"HadamardProductProof")
    for pa, pb, pc in zip(commit_a, commit_b, commit_c):
        tr.append_point_var(b"c_a", pa)
        tr.append_point_var(b"c_b", pb)
        tr.append_point_var(b"c_c", pc)
    a_0, b_0 = rnd["a_0"], rnd["b_0"]
    c_0 = [x * y % L for x, y in zip(a_0, b_0)]
    c_a_0, c_b_0, c_c_0 = xpc.commit(a_0, rnd["r_0"]), xpc.commit(b_0, rnd["s_0"]), xpc.commit(c_0, rnd["t_0"])
    omega = rnd["omega"]
    lp = l_polys(omega)

    def expression(m, m0):
        # compute_polynomial_expression: column j -> m0[j] l(X) + sum_i m[i][j] l_{i+1}(X)
        out = []
        for j in range(COLUMNS):
            p = poly_scale(lp[0], m0[j])
            for i in range(ROWS):
                p = poly_add(p, poly_scale(lp[i + 1], m[i][j]))
            out.append(p)
        return out
    ae, be, ce = expression(a, a_0), expression(b, b_0), expression(c, c_0)
    div = [poly_divexact(poly_sub(poly_mul(x, y), z), lp[0]) for x, y, z in zip(ae, be, ce)]
    delta_vec = [[(div[j][i] if i < len(div[j]) else 0) for j in range(3)] for i in range(4)]
    rho = rnd["rho"]
    c_delta = [xpc.commit(row, rho[i]) for i, row in enumerate(delta_vec)]
    tr.append_point_var(b"c_a_0", c_a_0)
    tr.append_point_var(b"c_b_0", c_b_0)
    tr.append_point_var(b"c_c_0", c_c_0)
    for cd in c_delta:
        tr.append_point_var(b"c_delta", cd)
    x = tr.get_challenge(b"challenge")
    ev = [poly_eval(p, x) for p in lp]
    r_bar = (rnd["r_0"] * ev[0] + sum(wr[i] * ev[i + 1] for i in range(3))) % L
    s_bar = (rnd["s_0"] * ev[0] + sum(ws[i] * ev[i + 1] for i in range(3))) % L
    t_bar = (rnd["t_0"] * ev[0] + sum(wt[i] * ev[i + 1] for i in range(3))) % L
    xs = exp_iter(x, 4)
    rho_bar = ev[0] * (sum(p * r for p, r in zip(xs, rho)) % L) % L
    proof = {"commitment_a_0": c_a_0, "commitment_b_0": c_b_0, "commitment_c_0": c_c_0, "commitment_delta": c_delta,
             "a_bar": [poly_eval(p, x) for p in ae], "b_bar": [poly_eval(p, x) for p in be],
             "c_bar": [poly_eval(p, x) for p in ce], "r_bar": r_bar, "s_bar": s_bar, "t_bar": t_bar, "rho_bar": rho_bar}
    return proof, list(omega)


def hadamard_verify(tr, proof, omega, commit_a, commit_b, commit_c, xpc):
    """-> True, None (Decompression Failed), or the reference's message tail: "omega", "abc", "delta"."""
    if len({w % L for w in omega}) != 3:
        return "omega"
    tr.domain_sep(b"HadamardProductProof")
    for pa, pb, pc in zip(commit_a, commit_b, commit_c):
        tr.append_point_var(b"c_a", pa)
        tr.append_point_var(b"c_b", pb)
        tr.append_point_var(b"c_c", pc)
    tr.append_point_var(b"c_a_0", proof["commitment_a_0"])
    tr.append_point_var(b"c_b_0", proof["commitment_b_0"])
    tr.append_point_var(b"c_c_0", proof["commitment_c_0"])
    for cd in proof["commitment_delta"]:
        tr.append_point_var(b"c_delta", cd)
    x = tr.get_challenge(b"challenge")
    ev = l_evals(omega, x)
    bars = [xpc.commit_point(proof["a_bar"], proof["r_bar"]), xpc.commit_point(proof["b_bar"], proof["s_bar"]),
            xpc.commit_point(proof["c_bar"], proof["t_bar"])]
    zeros = [R.decompress(proof[k]) for k in ("commitment_a_0", "commitment_b_0", "commitment_c_0")]
    if any(z is None for z in zeros):
        return None
    sums = [R.mul(ev[0], z) for z in zeros]
    for i in range(3):
        for k, cm in enumerate((commit_a, commit_b, commit_c)):
            p = R.decompress(cm[i])
            if p is None:
                return None
            sums[k] = R.add(sums[k], R.mul(ev[i + 1], p))
    if not all(R.eq(s, bar) for s, bar in zip(sums, bars)):
        return "abc"
    xs = exp_iter(x, 4)
    acc = R.decompress(proof["commitment_delta"][0])
    if acc is None:
        return None
    for i in range(1, 4):
        p = R.decompress(proof["commitment_delta"][i])
        if p is None:
            return None
        acc = R.add(acc, R.mul(xs[i], p))
    lhs = R.mul(ev[0], acc)
    diff = [(a * b - c) % L for a, b, c in zip(proof["a_bar"], proof["b_bar"], proof["c_bar"])]
    return True if R.eq(lhs, xpc.commit_point(diff, proof["rho_bar"])) else "delta"


# ---- product argument = multi-Hadamard + zero argument + SVP (reference src/shuffle/product.rs) --------------------------
def columns(m):
    return [list(c) for c in zip(*m)]


def pc_commit_point(value, blinding):
    """bulletproofs::PedersenGens::default().commit(value, blinding) = value * B + blinding * B_blinding."""
    return R.add(R.mul(value % L, R.BASEPOINT), R.mul(blinding % L, R.decompress(R.PEDERSEN_H_COMPRESSED)))


def single_bilinearmap(ai, bj, yi):
    """product.rs single_bilinearmap: sum_t a_t b_t y_t."""
    assert len(ai) == len(bj) == len(yi)
    return sum(a * b * y for a, b, y in zip(ai, bj, yi)) % L


def bilinearmap(a_cols, b_cols, y):
    """product.rs bilinearmap over the m + 1 = 4 columns of a and b: d_k = sum_{j = m - k + i} a_i * b_j (k = 0..2m)."""
    y_i = exp_iter(y, ROWS, skip=1)
    m = ROWS
    out = []
    for k in range(2 * ROWS + 1):
        s = 0
        for i in range(m + 1):
            j = m - k + i
            if 0 <= j <= m:
                s = (s + single_bilinearmap(a_cols[i], b_cols[j], y_i)) % L
        out.append(s)
    return out


def zero_prove(tr, xpc, a_cols3, b_cols3, r_vec, s_vec, y, rnd):
    """ZeroProof::create_zero_argument_proof.  a_cols3 / b_cols3: the three columns of a_2d / b_2d; r_vec: the blindings of
    the statement commitments (entries 1, 2 are used); s_vec: three blindings, a fourth (s_m) is appended as in the reference.
    rnd: a_0, b_m (COLUMNS each), r_0, s_m, t (2 ROWS + 1; entry ROWS + 1 is forced to 0)."""
    tr.domain_sep(b"ZeroArgumentProof")
    a_0, b_m, r_0, s_m = rnd["a_0"], rnd["b_m"], rnd["r_0"], rnd["s_m"]
    c_a_0, c_b_m = xpc.commit(a_0, r_0), xpc.commit(b_m, s_m)
    a_columns = [list(a_0)] + [list(c) for c in a_cols3]
    b_columns = [list(c) for c in b_cols3] + [list(b_m)]
    dv = bilinearmap(a_columns, b_columns, y)
    t = list(rnd["t"])
    t[ROWS + 1] = 0
    c_D = [R.compress(pc_commit_point(d, tt)) for d, tt in zip(dv, t)]
    tr.append_point_var(b"A0Commitment", c_a_0)
    tr.append_point_var(b"BmCommitment", c_b_m)
    for cd in c_D:
        tr.append_point_var(b"DCommitment", cd)
    x = tr.get_challenge(b"challenge")
    x_exp = exp_iter(x, 2 * ROWS + 1)
    x_exp_m = x_exp[0:ROWS + 1]
    x_m_j = x_exp_m[::-1]
    a_bar = [sum(a_columns[i][c] * x_exp_m[i] for i in range(ROWS + 1)) % L for c in range(ROWS)]
    b_bar = [sum(b_columns[i][c] * x_m_j[i] for i in range(ROWS + 1)) % L for c in range(ROWS)]
    r_ext = [r_0] + [r_vec[i] for i in range(1, ROWS)] + [0]
    s_full = list(s_vec) + [s_m]
    return {"c_A_0": c_a_0, "c_B_m": c_b_m, "c_D": c_D, "a_vec": a_bar, "b_vec": b_bar,
            "r": sum(p * q for p, q in zip(r_ext, x_exp_m)) % L, "s": sum(p * q for p, q in zip(s_full, x_m_j)) % L,
            "t": sum(p * q for p, q in zip(t, x_exp)) % L}


ZERO_ERRORS = {"size": "Zero Argument Proof Verify: Size check failed",
               "d": "Zero Argument Proof Verify: c_d_(m+1) == com(0,0) Failed",
               "a": "Zero Argument Proof Verify: com(a_bar, r) verification check Failed",
               "b": "Zero Argument Proof Verify: com(b_bar, s) verification check Failed",
               "ab": "Zero Argument Proof Verify: com(a_bar * b_bar, t) verification check Failed"}


def zero_verify(tr, proof, c_A, xpc, c_B_points, chal_y):
    """ZeroProof::verify.  c_A: 3 compressed points; c_B_points: 3 points (already decompressed by the caller, as in the
    reference).  -> True, None (an Err that comes from a failed decompression), or a key of ZERO_ERRORS."""
    if len(proof["c_D"]) != 2 * ROWS + 1 or len(proof["a_vec"]) != COLUMNS or len(proof["b_vec"]) != COLUMNS:
        return "size"
    d_m_1 = R.decompress(proof["c_D"][ROWS + 1])
    if d_m_1 is None:
        return None
    if not R.is_identity(d_m_1):
        return "d"
    tr.domain_sep(b"ZeroArgumentProof")
    tr.append_point_var(b"A0Commitment", proof["c_A_0"])
    tr.append_point_var(b"BmCommitment", proof["c_B_m"])
    for cd in proof["c_D"]:
        tr.append_point_var(b"DCommitment", cd)
    x = tr.get_challenge(b"challenge")
    x_exp = exp_iter(x, 2 * ROWS + 1)
    temp_a = _msm_point(x_exp[1:ROWS + 1], c_A)
    if temp_a is None:
        return None
    a0 = R.decompress(proof["c_A_0"])
    if a0 is None:
        return None
    if not R.eq(xpc.commit_point(proof["a_vec"], proof["r"]), R.add(a0, temp_a)):
        return "a"
    full = R.mul(0, R.BASEPOINT)
    for s, p in zip(x_exp[1:ROWS + 1][::-1], c_B_points):
        full = R.add(full, R.mul(s, p))
    bm = R.decompress(proof["c_B_m"])
    if bm is None:
        return None
    if not R.eq(xpc.commit_point(proof["b_vec"], proof["s"]), R.add(full, bm)):
        return "b"
    y_i = exp_iter(chal_y, ROWS, skip=1)
    abb = single_bilinearmap(proof["a_vec"], proof["b_vec"], y_i)
    dxk = _msm_point(x_exp, proof["c_D"])
    if dxk is None:
        return None
    return True if R.eq(pc_commit_point(abb, proof["t"]), dxk) else "ab"


def multihadamard_prove(tr, xpc, pi, bvec, comit_a, cb, r, s_3, rnd):
    """MultiHadamardProof::create_multi_hadamard_product_arg.  pi: 3 x 3 rows; comit_a[i] = xpc.commit(column i, r[i]);
    cb = xpc.commit(bvec, s_3).  rnd: s_mid (one scalar) and "zero" (the zero argument's randomness).
    -> (proof {c_B, zero_proof}, statement {c_b, zero_c_A})."""
    tr.domain_sep(b"MultiHadamardProductProof")
    cols = columns(pi)
    b2 = [p * q % L for p, q in zip(cols[0], cols[1])]
    s_vec_product = [r[0], rnd["s_mid"], s_3]
    c_B = [comit_a[0], xpc.commit(b2, s_vec_product[1]), cb]
    for c in c_B:
        tr.append_point_var(b"BVectorCommitment", c)
    x = tr.get_challenge(b"XChallenge")
    y = tr.get_challenge(b"YChallenge")
    xe = exp_iter(x, ROWS, skip=1)
    neg_one = [L - 1] * 3
    d_1 = [v * xe[0] % L for v in cols[0]]
    d_2 = [v * xe[1] % L for v in b2]
    d = [(p * xe[0] + q * xe[1]) % L for p, q in zip(b2, bvec)]
    s = [s_vec_product[0] * xe[0] % L, s_vec_product[1] * xe[1] % L, (xe[0] * s_vec_product[1] + xe[1] * s_vec_product[2]) % L]
    zero = zero_prove(tr, xpc, [cols[1], cols[2], neg_one], [d_1, d_2, d], r, s, y, rnd["zero"])
    c_minus_one = xpc.commit(neg_one, 0)
    return {"c_B": c_B, "zero_proof": zero}, {"c_b": cb, "zero_c_A": [comit_a[1], comit_a[2], c_minus_one]}


def multihadamard_verify(tr, proof, statement, c_A, xpc):
    """MultiHadamardProof::verify.  c_A: 3 compressed points (the reference holds them decompressed and compares their
    encodings).  -> True, None, "c_B_1", "c_B_m" or a zero_verify result."""
    cB, zA = proof["c_B"], statement["zero_c_A"]
    if not (c_A[0] == cB[0] and c_A[1] == zA[0] and c_A[2] == zA[1]):
        return "c_B_1"
    if statement["c_b"] != cB[ROWS - 1]:
        return "c_B_m"
    tr.domain_sep(b"MultiHadamardProductProof")
    for c in cB:
        tr.append_point_var(b"BVectorCommitment", c)
    x = tr.get_challenge(b"XChallenge")
    y = tr.get_challenge(b"YChallenge")
    xe = exp_iter(x, ROWS, skip=1)
    pts = [R.decompress(c) for c in cB]
    if any(p is None for p in pts):
        return None
    d_vec = [R.mul(xe[0], pts[0]), R.mul(xe[1], pts[1]), R.add(R.mul(xe[0], pts[1]), R.mul(xe[1], pts[2]))]
    c_zero_A = [zA[0], zA[1], xpc.commit([L - 1] * 3, 0)]
    return zero_verify(tr, proof["zero_proof"], c_zero_A, xpc, d_vec, y)


def product_prove(tr, xpc, pi, witness_r, rnd):
    """ProductProof::create_product_argument_proof.  rnd: s, "mh" (multihadamard_prove's rnd), "svp" = (d_vec, rd,
    delta_mid, s_1, s_x).  -> (proof, statement)."""
    cols = columns(pi)
    c_prod_A = [xpc.commit(cols[i], witness_r[i]) for i in range(ROWS)]
    bvec = [row[0] * row[1] * row[2] % L for row in pi]
    s = rnd["s"]
    cb = xpc.commit(bvec, s)
    b = bvec[0] * bvec[1] * bvec[2] % L
    mh_proof, mh_state = multihadamard_prove(tr, xpc, pi, bvec, c_prod_A, cb, witness_r, s, rnd["mh"])
    svp = svp_prove(tr, xpc, s, bvec, *rnd["svp"])
    return {"mh": mh_proof, "svp": svp}, {"mh": mh_state, "svp": (cb, b)}


def product_verify(tr, proof, statement, c_prod_A, xpc):
    res = multihadamard_verify(tr, proof["mh"], statement["mh"], c_prod_A, xpc)
    if res is not True:
        return res
    res = svp_verify(tr, proof["svp"], statement["svp"][0], statement["svp"][1], xpc)
    return "svp" if res is False else res


# ---- multi-exponentiation arguments (reference src/shuffle/multiexponential.rs) ------------------------------------------
def _ek_common(C_rows, A_rows):
    """create_ek_common: C_rows 3 x 3 compressed points, A_rows 4 x 3 scalars -> the 2m = 6 diagonal sums E_0..E_5."""
    def ms(a_idx, c_idx):
        sc, pt = [], []
        for i in a_idx:
            sc += A_rows[i]
        for i in c_idx:
            pt += C_rows[i]
        return _msm_point(sc, pt)
    return [ms([0], [2]), ms([0, 1], [1, 2]), ms([0, 1, 2], [0, 1, 2]), ms([1, 2, 3], [0, 1, 2]), ms([2, 3], [0, 1]), ms([3], [0])]


def _multiexpo_common(xpc, rnd):
    a_0, r_0 = rnd["a_0"], rnd["r_0"]
    b_vec, s_vec = list(rnd["b_vec"]), list(rnd["s_vec"])
    b_vec[ROWS] = 0
    s_vec[ROWS] = 0
    c_A_0 = xpc.commit(a_0, r_0)
    cb_k = [R.compress(pc_commit_point(b, s)) for b, s in zip(b_vec, s_vec)]
    return a_0, b_vec, s_vec, c_A_0, cb_k, r_0


def _multiexpo_response(a_witness, x_exp, a_0, s_dash, b_vec, s_vec, r_0):
    cols = columns(a_witness)
    a_vec = [(sum(cols[i][j] * x_exp[1 + j] for j in range(ROWS)) + a_0[i]) % L for i in range(ROWS)]
    r = (r_0 + sum(s_dash[j] * x_exp[1 + j] for j in range(ROWS))) % L
    return a_vec, r, sum(b * x for b, x in zip(b_vec, x_exp)) % L, sum(s * x for s, x in zip(s_vec, x_exp)) % L


def _rows3(flat):
    return [list(flat[0:3]), list(flat[3:6]), list(flat[6:9])]


def multiexpo_pubkey_prove(tr, xpc, pks, a_witness, s_dash, base_pk, rnd):
    """create_multiexponential_pubkey_proof.  pks: 9 x 64 B (updated keys), a_witness: 3 x 3 rows.
    rnd: a_0 (3), r_0, b_vec (6), s_vec (6)."""
    tr.domain_sep(b"MultiExponentialPubKeyProof")
    a_0, b_vec, s_vec, c_A_0, cb_k, r_0 = _multiexpo_common(xpc, rnd)
    A = [list(a_0)] + [list(r) for r in a_witness]
    e_g = _ek_common(_rows3([p[0:32] for p in pks]), A)
    e_h = _ek_common(_rows3([p[32:64] for p in pks]), A)
    Gb, Hb = R.decompress(base_pk[0:32]), R.decompress(base_pk[32:64])
    ek_g = [R.compress(R.add(R.mul(b % L, Gb), e)) for b, e in zip(b_vec, e_g)]
    ek_h = [R.compress(R.add(R.mul(b % L, Hb), e)) for b, e in zip(b_vec, e_h)]
    tr.append_point_var(b"A0Commitment", c_A_0)
    for cbk, g, h in zip(cb_k, ek_g, ek_h):
        tr.append_point_var(b"BKCommitment", cbk)
        tr.append_point_var(b"EK0Commitment", g)
        tr.append_point_var(b"EK1Commitment", h)
    x = tr.get_challenge(b"xchallenege")
    x_exp = exp_iter(x, 2 * ROWS)
    a_vec, r, bx, sx = _multiexpo_response(a_witness, x_exp, a_0, s_dash, b_vec, s_vec, r_0)
    return {"c_A_0": c_A_0, "c_B_k": cb_k, "E_k_0": ek_g, "E_k_1": ek_h, "a_vec": a_vec, "r": r, "b": bx, "s": sx, "t": 0}


def multiexpo_commit_prove(tr, xpc, comms, a_witness, s_dash, base_pk, rho, rnd):
    """create_multiexponential_elgamal_commit_proof.  comms: 9 x 64 B (updated commitments); rnd additionally tau_vec (6)."""
    tr.domain_sep(b"MultiExponentialElgamalCommmitmentProof")
    a_0, b_vec, s_vec, c_A_0, cb_k, r_0 = _multiexpo_common(xpc, rnd)
    tau = list(rnd["tau_vec"])
    tau[ROWS] = rho % L
    A = [list(a_0)] + [list(r) for r in a_witness]
    e_c = _ek_common(_rows3([c[0:32] for c in comms]), A)
    e_d = _ek_common(_rows3([c[32:64] for c in comms]), A)
    Gb, Hb = R.decompress(base_pk[0:32]), R.decompress(base_pk[32:64])
    E_c = [R.compress(R.add(R.mul(t % L, Gb), e)) for t, e in zip(tau, e_c)]
    E_d = [R.compress(R.add(R.add(R.mul(b % L, R.BASEPOINT), R.mul(t % L, Hb)), e)) for b, t, e in zip(b_vec, tau, e_d)]
    tr.append_point_var(b"A0Commitment", c_A_0)
    for cbk, c, d in zip(cb_k, E_c, E_d):
        tr.append_point_var(b"BKCommitment", cbk)
        tr.append_point_var(b"EK0Commitment", c)
        tr.append_point_var(b"EK1Commitment", d)
    x = tr.get_challenge(b"xchallenege")
    x_exp = exp_iter(x, 2 * ROWS)
    a_vec, r, bx, sx = _multiexpo_response(a_witness, x_exp, a_0, s_dash, b_vec, s_vec, r_0)
    return {"c_A_0": c_A_0, "c_B_k": cb_k, "E_k_0": E_c, "E_k_1": E_d, "a_vec": a_vec, "r": r, "b": bx, "s": sx,
            "t": sum(t * x for t, x in zip(tau, x_exp)) % L}


def _multiexpo_scalars(proof, c_A, x_exp, xpc):
    """verify_multiexpo_scalars -> True, None, "a", "b"."""
    lhs = _msm_point(x_exp[1:ROWS + 1], c_A)
    if lhs is None:
        return None
    a0 = R.decompress(proof["c_A_0"])
    if a0 is None:
        return None
    if not R.eq(R.add(lhs, a0), xpc.commit_point(proof["a_vec"], proof["r"])):
        return "a"
    bk = _msm_point(x_exp, proof["c_B_k"])
    if bk is None:
        return None
    return True if R.eq(pc_commit_point(proof["b"], proof["s"]), bk) else "b"


def _multiexpo_ek(proof, x_exp, c, d):
    """verify_multiexpo_ek: c, d = 9 decompressed points each -> (E_k_0 sum or None, E_k_1 sum or None, c_c_x, c_d_x)."""
    sc = ([a * x_exp[2] % L for a in proof["a_vec"]] + [a * x_exp[1] % L for a in proof["a_vec"]] + list(proof["a_vec"]))

    def ms(points):
        acc = R.mul(0, R.BASEPOINT)
        for s, p in zip(sc, points):
            acc = R.add(acc, R.mul(s % L, p))
        return acc
    return _msm_point(x_exp, proof["E_k_0"]), _msm_point(x_exp, proof["E_k_1"]), ms(c), ms(d)


def _multiexpo_transcript(tr, proof, domain):
    tr.domain_sep(domain)
    tr.append_point_var(b"A0Commitment", proof["c_A_0"])
    for cbk, e0, e1 in zip(proof["c_B_k"], proof["E_k_0"], proof["E_k_1"]):
        tr.append_point_var(b"BKCommitment", cbk)
        tr.append_point_var(b"EK0Commitment", e0)
        tr.append_point_var(b"EK1Commitment", e1)
    return exp_iter(tr.get_challenge(b"xchallenege"), 2 * ROWS)


def multiexpo_pubkey_verify(tr, proof, c_A, updated_accounts, base_pk, pk_GH, xpc):
    """verify_multiexponential_pubkey_proof -> True, None, "c_B_m", "Em", "a", "b", "E_K"."""
    if len(proof["a_vec"]) != COLUMNS or proof["c_B_k"][ROWS] != bytes(32):
        return "c_B_m"
    if pk_GH != proof["E_k_0"][ROWS] + proof["E_k_1"][ROWS]:
        return "Em"
    x_exp = _multiexpo_transcript(tr, proof, b"MultiExponentialPubKeyProof")
    res = _multiexpo_scalars(proof, c_A, x_exp, xpc)
    if res is not True:
        return res
    g = [R.decompress(a[0:32]) for a in updated_accounts]
    h = [R.decompress(a[32:64]) for a in updated_accounts]
    Gb, Hb = R.decompress(base_pk[0:32]), R.decompress(base_pk[32:64])
    if any(p is None for p in g + h + [Gb, Hb]):
        return None
    eg, eh, cg, ch = _multiexpo_ek(proof, x_exp, g, h)
    if eg is None:
        return None
    if not R.eq(eg, R.add(cg, R.mul(proof["b"] % L, Gb))):
        return "E_K"
    if eh is None:
        return None
    return True if R.eq(eh, R.add(ch, R.mul(proof["b"] % L, Hb))) else "E_K"


def multiexpo_commit_verify(tr, proof, c_A, updated_accounts, accounts, base_pk, exp_x, xpc):
    """verify_multiexponential_elgamal_commit_proof -> True, None, "c_B_m", "Em", "a", "b", "E_K"."""
    if len(proof["a_vec"]) != COLUMNS or proof["c_B_k"][ROWS] != bytes(32):
        return "c_B_m"
    C_c = _msm_point(exp_x, [a[64:96] for a in accounts])
    if C_c is None:
        return None
    C_d = _msm_point(exp_x, [a[96:128] for a in accounts])
    if C_d is None:
        return None
    if R.compress(C_c) != proof["E_k_0"][ROWS] or R.compress(C_d) != proof["E_k_1"][ROWS]:
        return "Em"
    x_exp = _multiexpo_transcript(tr, proof, b"MultiExponentialElgamalCommmitmentProof")
    res = _multiexpo_scalars(proof, c_A, x_exp, xpc)
    if res is not True:
        return res
    c = [R.decompress(a[64:96]) for a in updated_accounts]
    d = [R.decompress(a[96:128]) for a in updated_accounts]
    if any(p is None for p in c + d):
        return None
    Gb, Hb = R.decompress(base_pk[0:32]), R.decompress(base_pk[32:64])
    bb_c = R.mul(proof["t"] % L, Gb)
    bb_d = R.add(R.mul(proof["b"] % L, R.BASEPOINT), R.mul(proof["t"] % L, Hb))
    ec, ed, cc, cd = _multiexpo_ek(proof, x_exp, c, d)
    if ec is None:
        return None
    if not R.eq(ec, R.add(cc, bb_c)):
        return "E_K"
    if ed is None:
        return None
    return True if R.eq(ed, R.add(cd, bb_d)) else "E_K"


# ---- the shuffle proof (reference src/shuffle/shuffle.rs) -------------------------------------------------------------------
def input_shuffle(inputs, perm, tau, rho):
    """Shuffle::input_shuffle with its randomness as arguments: perm = a permutation of 1..9 (row major), tau (9), rho.
    -> dict(inputs, outputs, tau, rho, pi) as the reference's Shuffle struct holds them (pi already inverted)."""
    shuffled = [inputs[perm[i] - 1] for i in range(9)]
    inv = [0] * 9
    for i in range(9):
        inv[perm[i] - 1] = i + 1
    outs = []
    for acc, t in zip(inputs, tau):
        o, st = R.update_account(acc, sb(0), sb(t), sb(rho))
        assert st == 0
        outs.append(o)
    return {"inputs": shuffled, "outputs": outs, "tau": list(tau), "rho": rho, "pi": inv}


def shuffle_prove(tr, shuffle, xpc, rnd):
    """ShuffleProof::create_shuffle_proof.  rnd: r, r_dash, s, s_dash (3 each), "hadamard", "product", "ddh_r", "mexp_pk",
    "mexp_comm" (the sub-provers' randomness).  -> (proof, statement)."""
    pi, tau, rho = shuffle["pi"], shuffle["tau"], shuffle["rho"]
    witness = _rows3(pi)
    r, r_dash, s, s_dash = rnd["r"], rnd["r_dash"], rnd["s"], rnd["s_dash"]
    c_A = [xpc.commit(witness[i], r[i]) for i in range(ROWS)]
    tau_rows = _rows3(tau)
    c_tau = [xpc.commit(tau_rows[i], r_dash[i]) for i in range(COLUMNS)]
    for a, t in zip(c_A, c_tau):
        tr.append_point_var(b"ACommitment", a)
        tr.append_point_var(b"tauCommitment", t)
    x = tr.get_challenge(b"xChallenge")
    exp_x = exp_iter(x, 9, skip=1)
    x_psi = [exp_x[pi[i] - 1] for i in range(9)]
    b_dash = [x_psi[i] * pow(tau[i], -1, L) % L for i in range(9)]
    b_rows, b_dash_rows = _rows3(x_psi), _rows3(b_dash)
    c_B = [xpc.commit(b_rows[i], s[i]) for i in range(ROWS)]
    c_B_dash = [xpc.commit(b_dash_rows[i], s_dash[i]) for i in range(ROWS)]
    for cb, cbd in zip(c_B, c_B_dash):
        tr.append_point_var(b"BCommitment", cb)
        tr.append_point_var(b"BDashCommitment", cbd)
    had_proof, omega = hadamard_prove(tr, xpc, b_dash_rows, tau_rows, b_rows, c_B_dash, c_tau, c_B, s_dash, r_dash, s,
                                      rnd["hadamard"])
    y = tr.get_challenge(b"yChallenge")
    z = tr.get_challenge(b"zChallenge")
    e = [(a * y + b - z) % L for a, b in zip(pi, x_psi)]
    t = [(ri * y + si) % L for ri, si in zip(r, s)]
    e_2d_rows = [[e[c * ROWS + rr] for c in range(COLUMNS)] for rr in range(ROWS)]      # Array2D::from_column_major
    prod_proof, prod_state = product_prove(tr, xpc, e_2d_rows, t, rnd["product"])
    g_i = [a[0:32] for a in shuffle["inputs"]]
    h_i = [a[32:64] for a in shuffle["inputs"]]
    G, H = R.compress(_msm_point(exp_x, g_i)), R.compress(_msm_point(exp_x, h_i))
    ddh_proof, ddh_state = ddh_prove(tr, g_i, h_i, exp_x, G, H, rho, rnd["ddh_r"])
    upk = [a[0:64] for a in shuffle["outputs"]]
    ucomm = [a[64:128] for a in shuffle["outputs"]]
    mexp_pk = multiexpo_pubkey_prove(tr, xpc, upk, b_dash_rows, s_dash, R.BASE_PK, rnd["mexp_pk"])
    mexp_comm = multiexpo_commit_prove(tr, xpc, ucomm, b_rows, s, G + H, (-rho) % L, rnd["mexp_comm"])
    proof = {"c_A": c_A, "c_tau": c_tau, "c_B": c_B, "c_B_dash": c_B_dash, "hadamard": had_proof, "product": prod_proof,
             "mexp_pk": mexp_pk, "mexp_comm": mexp_comm, "ddh": ddh_proof}
    return proof, {"omega": omega, "product": prod_state, "ddh": ddh_state}


def shuffle_verify(tr, proof, statement, shuffle_input, shuffle_output, xpc):
    """ShuffleProof::verify -> (True, None) or (False, (stage, reason)): stage in "length", "hadamard", "product_b", "c_F",
    "product", "pk", "ddh", "mexp_pk", "mexp_comm"; reason = the sub-verifier's result (None = a decompression Err)."""
    if not all(len(proof[k]) == ROWS for k in ("c_A", "c_B", "c_B_dash", "c_tau")):
        return False, ("length", None)
    for ca, ctau in zip(proof["c_A"], proof["c_tau"]):
        tr.append_point_var(b"ACommitment", ca)
        tr.append_point_var(b"tauCommitment", ctau)
    x = tr.get_challenge(b"xChallenge")
    exp_x = exp_iter(x, 9, skip=1)
    for b, bd in zip(proof["c_B"], proof["c_B_dash"]):
        tr.append_point_var(b"BCommitment", b)
        tr.append_point_var(b"BDashCommitment", bd)
    res = hadamard_verify(tr, proof["hadamard"], statement["omega"], proof["c_B_dash"], proof["c_tau"], proof["c_B"], xpc)
    if res is not True:
        return False, ("hadamard", res)
    y = tr.get_challenge(b"yChallenge")
    z = tr.get_challenge(b"zChallenge")
    product = 1
    for i, xi in enumerate(exp_x):
        product = product * (y * (i + 1) + xi - z) % L
    if product != statement["product"]["svp"][1] % L:
        return False, ("product_b", False)
    z_neg = xpc.commit_point([(-z) % L] * 3, 0)
    c_E = []
    for ca, cb in zip(proof["c_A"], proof["c_B"]):
        pa, pb = R.decompress(ca), R.decompress(cb)
        if pa is None or pb is None:
            return False, ("c_F", None)
        c_E.append(R.compress(R.add(R.add(R.mul(y, pa), pb), z_neg)))
    res = product_verify(tr, proof["product"], statement["product"], c_E, xpc)
    if res is not True:
        return False, ("product", res)
    g_i = [a[0:32] for a in shuffle_input]
    h_i = [a[32:64] for a in shuffle_input]
    Gp, Hp = _msm_point(exp_x, g_i), _msm_point(exp_x, h_i)
    if Gp is None or Hp is None:
        return False, ("pk", None)
    G, H = R.compress(Gp), R.compress(Hp)
    res = ddh_verify(tr, proof["ddh"], statement["ddh"], G, H)
    if res is not True:
        return False, ("ddh", res)
    res = multiexpo_pubkey_verify(tr, proof["mexp_pk"], proof["c_B_dash"], shuffle_output, R.BASE_PK, G + H, xpc)
    if res is not True:
        return False, ("mexp_pk", res)
    res = multiexpo_commit_verify(tr, proof["mexp_comm"], proof["c_B"], shuffle_output, shuffle_input, G + H, exp_x, xpc)
    if res is not True:
        return False, ("mexp_comm", res)
    return True, None
