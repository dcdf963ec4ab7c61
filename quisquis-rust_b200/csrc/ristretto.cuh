// Ristretto255 decode / encode (RFC 9496 4.3.1 / 4.3.2) on top of fe25519.cuh / ge25519.cuh.
//
// Replaces curve25519-dalek 3.x `ristretto.rs` CompressedRistretto::decompress and RistrettoPoint::compress, which
// the reference calls on every API entry/exit (e.g. src/ristretto/keys.rs:278-281, src/elgamal/elgamal.rs:47-52,
// src/accounts/verifier.rs:95-98).  Branch-free; the reject rules are part of parity.
#pragma once
#include "ge25519.cuh"

namespace qq {

// invsqrt(v) = sqrt_ratio_i(1, v) with the multiplications by u = 1 removed.
QQ_HD u32 fe_invsqrt(fe& r, const fe& v) {
    fe v3, v7, t, check, one, mone, monei;
    fe_sq(v3, v);
    fe_mul(v3, v3, v);
    fe_sq(v7, v3);
    fe_mul(v7, v7, v);
    fe_pow22523(t, v7);
    fe_mul(r, v3, t);
    fe_sq(check, r);
    fe_mul(check, v, check);
    fe_1(one);
    fe_neg(mone, one);
    fe_neg(monei, fe_sqrt_m1());
    u32 correct = fe_eq(check, one);
    u32 flipped = fe_eq(check, mone);
    u32 flipped_i = fe_eq(check, monei);
    fe ri;
    fe_mul(ri, r, fe_sqrt_m1());
    fe_cmov(r, ri, flipped | flipped_i);
    fe_abs(r);
    return correct | flipped;
}

// 32 little-endian bytes (as 8 words) -> extended point with Z = 1.  Returns 1 if the encoding is valid.
QQ_HD u32 ristretto_decompress(ge_p3& p, const u32 w[8]) {
    fe s, ss, u1, u2, u1s, u2s, v, I, dx, dy, t;
    fe_fromwords(s, w);
    u32 cw[8];
    fe_towords(cw, s);
    u32 diff = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) diff |= cw[i] ^ w[i];
    u32 canonical = diff == 0 ? 1u : 0u;       // rejects s >= p and bit 255 set
    u32 s_neg = w[0] & 1u;
    fe_sq(ss, s);
    fe one;
    fe_1(one);
    fe_sub(u1, one, ss);
    fe_add(u2, one, ss);
    fe_sq(u2s, u2);
    fe_sq(u1s, u1);
    fe_mul(t, u1s, fe_d());
    fe_neg(v, t);
    fe_sub(v, v, u2s);
    fe_mul(t, v, u2s);
    u32 ok = fe_invsqrt(I, t);
    fe_mul(dx, I, u2);
    fe_mul(t, I, dx);
    fe_mul(dy, v, t);
    fe_add(t, s, s);
    fe_mul(p.X, t, dx);
    fe_abs(p.X);
    fe_mul(p.Y, u1, dy);
    fe_1(p.Z);
    fe_mul(p.T, p.X, p.Y);
    u32 t_neg = fe_isnegative(p.T);
    u32 y_zero = fe_iszero(p.Y);
    return canonical & (s_neg ^ 1u) & ok & (t_neg ^ 1u) & (y_zero ^ 1u);
}

// extended point -> canonical 32-byte encoding (8 words)
QQ_HD void ristretto_compress(u32 w[8], const ge_p3& p) {
    fe u1, u2, t, inv, i1, i2, zinv, deninv, ix, iy, ed, x, y, zy;
    fe_add(u1, p.Z, p.Y);
    fe_sub(t, p.Z, p.Y);
    fe_mul(u1, u1, t);
    fe_mul(u2, p.X, p.Y);
    fe_sq(t, u2);
    fe_mul(t, u1, t);
    fe_invsqrt(inv, t);                          // always square for valid points
    fe_mul(i1, inv, u1);
    fe_mul(i2, inv, u2);
    fe_mul(t, i1, i2);
    fe_mul(zinv, t, p.T);
    fe_mul(ix, p.X, fe_sqrt_m1());
    fe_mul(iy, p.Y, fe_sqrt_m1());
    fe_mul(ed, i1, fe_invsqrt_a_minus_d());
    fe_mul(t, p.T, zinv);
    u32 rotate = fe_isnegative(t);
    x = p.X;
    y = p.Y;
    deninv = i2;
    fe_cmov(x, iy, rotate);
    fe_cmov(y, ix, rotate);
    fe_cmov(deninv, ed, rotate);
    fe_mul(t, x, zinv);
    fe yc = y;
    fe_cneg(yc, fe_isnegative(t));
    fe_sub(zy, p.Z, yc);
    fe_mul(t, zy, deninv);
    fe_abs(t);
    fe_towords(w, t);
}

// Elligator 2 map of RFC 9496 4.3.4 (dalek ristretto.rs elligator_ristretto_flavor): field element -> point.
// Reached in the reference through RistrettoPoint::hash_from_bytes::<Sha3_512> (src/pedersen/vectorpedersen.rs:49-51,
// 66-70) and through the Bulletproofs generator chains (src/accounts/verifier.rs:510,540).
QQ_HD void ristretto_elligator(ge_p3& p, const fe& r0) {
    fe r, ns, c, dd, t, u, s, sp, nt, w0, w1, w2, w3, one, ss;
    fe_1(one);
    fe_sq(t, r0);
    fe_mul(r, t, fe_sqrt_m1());                 // r = i r0^2
    fe_add(t, r, one);
    fe_mul(ns, t, fe_one_minus_d_sq());         // (r + 1)(1 - d^2)
    fe_neg(c, one);                             // c = -1
    fe_mul(t, fe_d(), r);
    fe_sub(t, c, t);                            // c - d r
    fe_add(u, r, fe_d());
    fe_mul(dd, t, u);                           // (c - d r)(r + d)
    u32 was_square = fe_sqrt_ratio_i(s, ns, dd);
    fe_mul(sp, s, r0);
    fe_abs(sp);
    fe_neg(sp, sp);                             // s' = -|s r0|
    fe_cmov(s, sp, was_square ^ 1u);
    fe_cmov(c, r, was_square ^ 1u);
    fe_sub(t, r, one);
    fe_mul(t, c, t);
    fe_mul(t, t, fe_d_minus_one_sq());
    fe_sub(nt, t, dd);                          // c (r - 1)(d - 1)^2 - D
    fe_sq(ss, s);
    fe_add(t, s, s);
    fe_mul(w0, t, dd);                          // 2 s D
    fe_mul(w1, nt, fe_sqrt_ad_minus_one());
    fe_sub(w2, one, ss);
    fe_add(w3, one, ss);
    fe_mul(p.X, w0, w3);
    fe_mul(p.Y, w2, w1);
    fe_mul(p.Z, w1, w3);
    fe_mul(p.T, w0, w2);
}
// RistrettoPoint::from_uniform_bytes: 64 bytes (16 words) -> elligator(lo) + elligator(hi); bit 255 of each half ignored
QQ_HD void ristretto_from_uniform(ge_p3& p, const u32 w[16]) {
    fe r0, r1;
    fe_fromwords(r0, w);
    fe_fromwords(r1, w + 8);
    ge_p3 a, b;
    ristretto_elligator(a, r0);
    ristretto_elligator(b, r1);
    ge_cached cb;
    ge_to_cached(cb, b);
    ge_add(p, a, cb);
}

}  // namespace qq
