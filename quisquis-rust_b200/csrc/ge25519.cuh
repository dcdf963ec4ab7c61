// Edwards25519 (a = -1) group operations in extended coordinates, on top of fe25519.cuh.
//
// Replaces (for the hot path) curve25519-dalek 3.x `backend/serial/curve_models/mod.rs`
// (ProjectivePoint::double, EdwardsPoint + ProjectiveNielsPoint / AffineNielsPoint) which every reference call
// site reaches through `&Scalar * &RistrettoPoint`, point `+`/`-` and `multiscalar_mul`
// (reference src/ristretto/keys.rs:277-281, src/elgamal/elgamal.rs:47-52,66-68).  Formulas: HWCD-2008 (RFC 8032 5.1.4).
//
// Field elements are saturated 8 x 32-bit values in [0, 2^256) (fe25519.cuh): every fe_* accepts the full range, so no
// magnitude bookkeeping is needed here.
#pragma once
#include "fe25519.cuh"

namespace qq {

struct ge_p3 {      // extended (X:Y:Z:T), x = X/Z, y = Y/Z, T = XY/Z
    fe X, Y, Z, T;
};
struct ge_cached {  // projective Niels with the factor 2 of the addition formula folded in: (Y+X, Y-X, 2Z, 2dT)
    fe YpX, YmX, Z2, T2d;
};
struct ge_niels {   // affine Niels (Z = 1): (y+x, y-x, 2dxy)
    fe ypx, ymx, xy2d;
};

QQ_HD void ge_identity(ge_p3& p) {
    fe_0(p.X);
    fe_1(p.Y);
    fe_1(p.Z);
    fe_0(p.T);
}

QQ_HD void ge_to_cached(ge_cached& c, const ge_p3& p) {
    fe_add(c.YpX, p.Y, p.X);
    fe_sub(c.YmX, p.Y, p.X);
    fe_add(c.Z2, p.Z, p.Z);
    fe_mul(c.T2d, p.T, fe_2d());
}
// affine Niels from a point with Z == 1 (e.g. straight out of decompress)
QQ_HD void ge_to_niels_z1(ge_niels& n, const ge_p3& p) {
    fe_add(n.ypx, p.Y, p.X);
    fe_sub(n.ymx, p.Y, p.X);
    fe_mul(n.xy2d, p.T, fe_2d());
}

QQ_HD void ge_neg(ge_p3& r, const ge_p3& p) {
    fe_neg(r.X, p.X);
    r.Y = p.Y;
    r.Z = p.Z;
    fe_neg(r.T, p.T);
}
// c = b ? -c : c   for cached points: swap (Y+X, Y-X), negate 2dT
QQ_HD void ge_cached_cneg(ge_cached& c, u32 b) {
    u32 m = 0u - b;
#pragma unroll
    for (int i = 0; i < QQ_FE_LIMBS; i++) {
        u32 x = m & (c.YpX.v[i] ^ c.YmX.v[i]);
        c.YpX.v[i] ^= x;
        c.YmX.v[i] ^= x;
    }
    fe n;
    fe_neg(n, c.T2d);
    fe_cmov(c.T2d, n, b);
}
QQ_HD void ge_niels_cneg(ge_niels& c, u32 b) {
    u32 m = 0u - b;
#pragma unroll
    for (int i = 0; i < QQ_FE_LIMBS; i++) {
        u32 x = m & (c.ypx.v[i] ^ c.ymx.v[i]);
        c.ypx.v[i] ^= x;
        c.ymx.v[i] ^= x;
    }
    fe n;
    fe_neg(n, c.xy2d);
    fe_cmov(c.xy2d, n, b);
}

// Field-multiplication flavour used inside a group operation: INL = true inlines the products (straight-line code
// whose independent multiplications ptxas interleaves), INL = false calls the out-of-line fe_mul / fe_sq.
template <bool INL>
QQ_HD void fe_mul_t(fe& h, const fe& f, const fe& g) {
    if (INL) fe_mul_inl(h, f, g);
    else fe_mul(h, f, g);
}
template <bool INL>
QQ_HD void fe_sq_t(fe& h, const fe& f) {
    if (INL) fe_sq_inl(h, f);
    else fe_sq(h, f);
}

// r = p + q  (8 M)
template <bool INL>
QQ_HD void ge_add_t(ge_p3& r, const ge_p3& p, const ge_cached& q) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sub(t, p.Y, p.X);
    fe_mul_t<INL>(a, t, q.YmX);
    fe_add(t, p.Y, p.X);
    fe_mul_t<INL>(b, t, q.YpX);
    fe_mul_t<INL>(c, p.T, q.T2d);
    fe_mul_t<INL>(d, p.Z, q.Z2);           // 2 Z1 Z2
    fe_sub(e, b, a);
    fe_sub(f, d, c);
    fe_add(g, d, c);
    fe_add(h, b, a);
    fe_mul_t<INL>(r.X, f, e);
    fe_mul_t<INL>(r.Y, g, h);
    fe_mul_t<INL>(r.Z, f, g);
    fe_mul_t<INL>(r.T, e, h);
}
#if defined(__CUDA_ARCH__) && !defined(QQ_INLINE_FIELD_OPS) && !defined(QQ_FE_SINGLE)
// device build: the two rounds of four independent products go through fe_mul4_ool (fe25519.cuh)
QQ_HD void ge_add(ge_p3& r, const ge_p3& p, const ge_cached& q) {
    fe e, f, g, h, t, u;
    fe_sub(t, p.Y, p.X);
    fe_add(u, p.Y, p.X);
    fe4 m = fe_mul4_ool(t, q.YmX, u, q.YpX, p.T, q.T2d, p.Z, q.Z2);
    fe_sub(e, m.b, m.a);
    fe_sub(f, m.d, m.c);
    fe_add(g, m.d, m.c);
    fe_add(h, m.b, m.a);
    fe4 o = fe_mul4_ool(f, e, g, h, f, g, e, h);
    r.X = o.a; r.Y = o.b; r.Z = o.c; r.T = o.d;
}
#else
QQ_HD void ge_add(ge_p3& r, const ge_p3& p, const ge_cached& q) { ge_add_t<false>(r, p, q); }
#endif
// r = p + q, q affine Niels (7 M)
template <bool INL>
QQ_HD void ge_madd_t(ge_p3& r, const ge_p3& p, const ge_niels& q) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sub(t, p.Y, p.X);
    fe_mul_t<INL>(a, t, q.ymx);
    fe_add(t, p.Y, p.X);
    fe_mul_t<INL>(b, t, q.ypx);
    fe_mul_t<INL>(c, p.T, q.xy2d);
    fe_add(d, p.Z, p.Z);
    fe_sub(e, b, a);
    fe_sub(f, d, c);
    fe_add(g, d, c);
    fe_add(h, b, a);
    fe_mul_t<INL>(r.X, f, e);
    fe_mul_t<INL>(r.Y, g, h);
    fe_mul_t<INL>(r.Z, f, g);
    fe_mul_t<INL>(r.T, e, h);
}
#if defined(__CUDA_ARCH__) && !defined(QQ_INLINE_FIELD_OPS) && !defined(QQ_FE_SINGLE)
QQ_HD void ge_madd(ge_p3& r, const ge_p3& p, const ge_niels& q) {
    fe d, e, f, g, h, t, u;
    fe_sub(t, p.Y, p.X);
    fe_add(u, p.Y, p.X);
    fe3 m = fe_mul3_ool(t, q.ymx, u, q.ypx, p.T, q.xy2d);
    fe_add(d, p.Z, p.Z);
    fe_sub(e, m.b, m.a);
    fe_sub(f, d, m.c);
    fe_add(g, d, m.c);
    fe_add(h, m.b, m.a);
    fe4 o = fe_mul4_ool(f, e, g, h, f, g, e, h);
    r.X = o.a; r.Y = o.b; r.Z = o.c; r.T = o.d;
}
#else
QQ_HD void ge_madd(ge_p3& r, const ge_p3& p, const ge_niels& q) { ge_madd_t<false>(r, p, q); }
#endif

// r = 2p.  WITH_T = false skips T3 (4S + 3M) when the next operation is another doubling.  p.T is not read.
template <bool WITH_T, bool INL>
QQ_HD void ge_dbl_t(ge_p3& r, const ge_p3& p) {
    fe xx, yy, zz, s, cx, cy, cz, ct, t;
    fe_sq_t<INL>(xx, p.X);
    fe_sq_t<INL>(yy, p.Y);
    fe_sq_t<INL>(zz, p.Z);
    fe_add(t, p.X, p.Y);
    fe_sq_t<INL>(s, t);
    fe_add(cy, yy, xx);
    fe_sub(cz, yy, xx);
    fe_sub(cx, s, cy);
    fe_add(t, zz, zz);
    fe_sub(ct, t, cz);              // 2ZZ - (YY - XX)
    fe_mul_t<INL>(r.X, cx, ct);
    fe_mul_t<INL>(r.Y, cy, cz);
    fe_mul_t<INL>(r.Z, cz, ct);
    if (WITH_T) fe_mul_t<INL>(r.T, cx, cy);
}
#if defined(__CUDA_ARCH__) && !defined(QQ_INLINE_FIELD_OPS) && !defined(QQ_FE_SINGLE)
template <bool WITH_T>
QQ_HD void ge_dbl(ge_p3& r, const ge_p3& p) {
    fe cx, cy, cz, ct, t;
    fe_add(t, p.X, p.Y);
    fe4 q = fe_sq4_ool(p.X, p.Y, p.Z, t);
    fe_add(cy, q.b, q.a);
    fe_sub(cz, q.b, q.a);
    fe_sub(cx, q.d, cy);
    fe_add(t, q.c, q.c);
    fe_sub(ct, t, cz);              // 2ZZ - (YY - XX)
    if (WITH_T) {
        fe4 o = fe_mul4_ool(cx, ct, cy, cz, cz, ct, cx, cy);
        r.X = o.a; r.Y = o.b; r.Z = o.c; r.T = o.d;
    } else {
        fe3 o = fe_mul3_ool(cx, ct, cy, cz, cz, ct);
        r.X = o.a; r.Y = o.b; r.Z = o.c;
    }
}
#else
template <bool WITH_T>
QQ_HD void ge_dbl(ge_p3& r, const ge_p3& p) { ge_dbl_t<WITH_T, false>(r, p); }
#endif

// Ristretto equality (dalek RistrettoPoint::ct_eq): X1*Y2 == Y1*X2  or  X1*X2 == Y1*Y2
QQ_HD u32 ge_ristretto_eq(const ge_p3& p, const ge_p3& q) {
    fe a, b;
    fe_mul(a, p.X, q.Y);
    fe_mul(b, p.Y, q.X);
    u32 e1 = fe_eq(a, b);
    fe_mul(a, p.X, q.X);
    fe_mul(b, p.Y, q.Y);
    u32 e2 = fe_eq(a, b);
    return e1 | e2;
}
// identity coset test: X == 0 or Y == 0
QQ_HD u32 ge_ristretto_is_identity(const ge_p3& p) { return fe_iszero(p.X) | fe_iszero(p.Y); }

}  // namespace qq
