/* CPU ORACLE (test infrastructure, NOT a product path) -- plain-C restatement of the reference's hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * Nothing under quisquis-rust_b200/ links or calls it.
 *
 * What it restates.  quisquis-rust performs all group arithmetic through the third-party crate
 * curve25519-dalek = "3" (reference Cargo.toml:42; 3.2.1 per RELEASE_NOTES.md:24), which is NOT vendored under
 * /root/reference and cannot be built here (no cargo/rustc).  This file therefore restates the published
 * algorithms that crate uses, so that it can double as the CPU baseline:
 *   - field GF(2^255-19), radix 2^51, 5 x u64 with 128-bit products      [dalek backend/serial/u64/field.rs]
 *   - sqrt_ratio_i / invsqrt, Ristretto decode / encode (RFC 9496 4.3)    [dalek field.rs, ristretto.rs]
 *   - extended twisted-Edwards add / double, projective- and affine-Niels [dalek backend/serial/curve_models]
 *   - variable-base multiplication: signed radix-16, 8-entry table        [dalek scalar_mul/variable_base.rs]
 *   - fixed-base multiplication from a precomputed radix-16 table         [dalek edwards.rs EdwardsBasepointTable]
 *   - Straus (n < 190) and Pippenger (n >= 190; w = 6/7/8) MSM            [dalek scalar_mul/{straus,pippenger}.rs]
 * and, on top, the reference call sites (file:line under /root/reference):
 *   update_public_key   src/ristretto/keys.rs:146-148     generate_commitment  src/elgamal/elgamal.rs:41-53
 *   add_commitments     src/elgamal/elgamal.rs:65-69      update_account       src/accounts/accounts.rs:143-154
 *   verify_account      src/accounts/accounts.rs:81-84    delta/epsilon        src/accounts/accounts.rs:198-220
 *   multiscalar_multiplication  src/accounts/verifier.rs:91-99
 * Pinned by tests/test_oracle.py against RFC 9496 vectors, oracle/ristretto_ref.py (big-int) and libsodium 1.0.20.
 * Batch entry points are parallelised with OpenMP over all host cores (the rayon-style baseline of BASELINE.md).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef unsigned __int128 u128;
typedef struct { u64 v[5]; } fe;
typedef struct { fe X, Y, Z, T; } ge;
typedef struct { fe YpX, YmX, Z, T2d; } ge_cached;
typedef struct { fe ypx, ymx, xy2d; } ge_niels;

#define M51 0x7ffffffffffffULL

/* ------------------------------------------------------------------ field ---------------------------------- */
static void fe_0(fe* h) { memset(h, 0, sizeof *h); }
static void fe_1(fe* h) { fe_0(h); h->v[0] = 1; }
static void fe_add(fe* h, const fe* f, const fe* g) { for (int i = 0; i < 5; i++) h->v[i] = f->v[i] + g->v[i]; }
static void fe_carry(fe* h) {
    u64 c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
    c = h->v[1] >> 51; h->v[1] &= M51; h->v[2] += c;
    c = h->v[2] >> 51; h->v[2] &= M51; h->v[3] += c;
    c = h->v[3] >> 51; h->v[3] &= M51; h->v[4] += c;
    c = h->v[4] >> 51; h->v[4] &= M51; h->v[0] += 19 * c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
}
/* h = f - g ; adds 4p first so g may have limbs up to 2^53 */
static void fe_sub(fe* h, const fe* f, const fe* g) {
    h->v[0] = f->v[0] + 0x1fffffffffffb4ULL - g->v[0];
    for (int i = 1; i < 5; i++) h->v[i] = f->v[i] + 0x1ffffffffffffcULL - g->v[i];
    fe_carry(h);
}
static void fe_neg(fe* h, const fe* f) { fe z; fe_0(&z); fe_sub(h, &z, f); }
static void fe_mul(fe* h, const fe* f, const fe* g) {
    u128 f0 = f->v[0], f1 = f->v[1], f2 = f->v[2], f3 = f->v[3], f4 = f->v[4];
    u64 g0 = g->v[0], g1 = g->v[1], g2 = g->v[2], g3 = g->v[3], g4 = g->v[4];
    u64 g1_19 = 19 * g1, g2_19 = 19 * g2, g3_19 = 19 * g3, g4_19 = 19 * g4;
    u128 r0 = f0 * g0 + f1 * g4_19 + f2 * g3_19 + f3 * g2_19 + f4 * g1_19;
    u128 r1 = f0 * g1 + f1 * g0 + f2 * g4_19 + f3 * g3_19 + f4 * g2_19;
    u128 r2 = f0 * g2 + f1 * g1 + f2 * g0 + f3 * g4_19 + f4 * g3_19;
    u128 r3 = f0 * g3 + f1 * g2 + f2 * g1 + f3 * g0 + f4 * g4_19;
    u128 r4 = f0 * g4 + f1 * g3 + f2 * g2 + f3 * g1 + f4 * g0;
    u64 c;
    r1 += (u64)(r0 >> 51); h->v[0] = (u64)r0 & M51;
    r2 += (u64)(r1 >> 51); h->v[1] = (u64)r1 & M51;
    r3 += (u64)(r2 >> 51); h->v[2] = (u64)r2 & M51;
    r4 += (u64)(r3 >> 51); h->v[3] = (u64)r3 & M51;
    c = (u64)(r4 >> 51);   h->v[4] = (u64)r4 & M51;
    h->v[0] += 19 * c;
    c = h->v[0] >> 51; h->v[0] &= M51; h->v[1] += c;
}
static void fe_sq(fe* h, const fe* f) { fe_mul(h, f, f); }
static void fe_sqn(fe* h, const fe* f, int n) { fe_sq(h, f); for (int i = 1; i < n; i++) fe_sq(h, h); }
static void fe_pow22523(fe* out, const fe* z) {
    fe t0, t1, t2;
    fe_sq(&t0, z); fe_sqn(&t1, &t0, 2); fe_mul(&t1, z, &t1); fe_mul(&t0, &t0, &t1);
    fe_sq(&t0, &t0); fe_mul(&t0, &t1, &t0);
    fe_sqn(&t1, &t0, 5); fe_mul(&t0, &t1, &t0);
    fe_sqn(&t1, &t0, 10); fe_mul(&t1, &t1, &t0);
    fe_sqn(&t2, &t1, 20); fe_mul(&t1, &t2, &t1);
    fe_sqn(&t1, &t1, 10); fe_mul(&t0, &t1, &t0);
    fe_sqn(&t1, &t0, 50); fe_mul(&t1, &t1, &t0);
    fe_sqn(&t2, &t1, 100); fe_mul(&t1, &t2, &t1);
    fe_sqn(&t1, &t1, 50); fe_mul(&t0, &t1, &t0);
    fe_sqn(&t0, &t0, 2); fe_mul(out, &t0, z);
}
static void fe_invert(fe* out, const fe* z) {
    fe t, z2, z3;
    fe_pow22523(&t, z); fe_sqn(&t, &t, 3); fe_sq(&z2, z); fe_mul(&z3, &z2, z); fe_mul(out, &t, &z3);
}
static void fe_tobytes(uint8_t* s, const fe* f) {
    fe t = *f;
    fe_carry(&t); fe_carry(&t);
    u64 q = (t.v[0] + 19) >> 51;
    q = (t.v[1] + q) >> 51; q = (t.v[2] + q) >> 51; q = (t.v[3] + q) >> 51; q = (t.v[4] + q) >> 51;
    t.v[0] += 19 * q;
    u64 c;
    c = t.v[0] >> 51; t.v[0] &= M51; t.v[1] += c;
    c = t.v[1] >> 51; t.v[1] &= M51; t.v[2] += c;
    c = t.v[2] >> 51; t.v[2] &= M51; t.v[3] += c;
    c = t.v[3] >> 51; t.v[3] &= M51; t.v[4] += c;
    t.v[4] &= M51;
    u64 w0 = t.v[0] | (t.v[1] << 51);
    u64 w1 = (t.v[1] >> 13) | (t.v[2] << 38);
    u64 w2 = (t.v[2] >> 26) | (t.v[3] << 25);
    u64 w3 = (t.v[3] >> 39) | (t.v[4] << 12);
    memcpy(s, &w0, 8); memcpy(s + 8, &w1, 8); memcpy(s + 16, &w2, 8); memcpy(s + 24, &w3, 8);
}
static void fe_frombytes(fe* h, const uint8_t* s) {
    u64 w0, w1, w2, w3;
    memcpy(&w0, s, 8); memcpy(&w1, s + 8, 8); memcpy(&w2, s + 16, 8); memcpy(&w3, s + 24, 8);
    h->v[0] = w0 & M51;
    h->v[1] = ((w0 >> 51) | (w1 << 13)) & M51;
    h->v[2] = ((w1 >> 38) | (w2 << 26)) & M51;
    h->v[3] = ((w2 >> 25) | (w3 << 39)) & M51;
    h->v[4] = (w3 >> 12) & M51;
}
static int fe_isneg(const fe* f) { uint8_t s[32]; fe_tobytes(s, f); return s[0] & 1; }
static int fe_iszero(const fe* f) {
    uint8_t s[32]; fe_tobytes(s, f);
    uint8_t r = 0; for (int i = 0; i < 32; i++) r |= s[i];
    return r == 0;
}
static int fe_eq(const fe* f, const fe* g) { uint8_t a[32], b[32]; fe_tobytes(a, f); fe_tobytes(b, g); return memcmp(a, b, 32) == 0; }
static void fe_cneg(fe* h, int b) { if (b) { fe n; fe_neg(&n, h); *h = n; } }

static fe C_D, C_2D, C_SQRTM1, C_INVSQRT_A_MINUS_D;
static void fe_from_u64(fe* h, u64 x) { fe_0(h); h->v[0] = x & M51; h->v[1] = x >> 51; }

/* sqrt_ratio_i: RFC 9496 4.2 */
static int fe_sqrt_ratio_i(fe* r, const fe* u, const fe* v) {
    fe v3, v7, t, check, nu, nui;
    fe_sq(&v3, v); fe_mul(&v3, &v3, v);
    fe_sq(&v7, &v3); fe_mul(&v7, &v7, v);
    fe_mul(&t, u, &v7); fe_pow22523(&t, &t);
    fe_mul(r, u, &v3); fe_mul(r, r, &t);
    fe_sq(&check, r); fe_mul(&check, v, &check);
    fe_neg(&nu, u); fe_mul(&nui, &nu, &C_SQRTM1);
    int correct = fe_eq(&check, u), flipped = fe_eq(&check, &nu), flipped_i = fe_eq(&check, &nui);
    if (flipped | flipped_i) fe_mul(r, r, &C_SQRTM1);
    fe_cneg(r, fe_isneg(r));
    return correct | flipped;
}

/* ------------------------------------------------------------------ group ---------------------------------- */
static void ge_identity(ge* p) { fe_0(&p->X); fe_1(&p->Y); fe_1(&p->Z); fe_0(&p->T); }
static void ge_to_cached(ge_cached* c, const ge* p) {
    fe_add(&c->YpX, &p->Y, &p->X); fe_sub(&c->YmX, &p->Y, &p->X); c->Z = p->Z; fe_mul(&c->T2d, &p->T, &C_2D);
}
static void ge_add(ge* r, const ge* p, const ge_cached* q) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sub(&t, &p->Y, &p->X); fe_mul(&a, &t, &q->YmX);
    fe_add(&t, &p->Y, &p->X); fe_mul(&b, &t, &q->YpX);
    fe_mul(&c, &p->T, &q->T2d);
    fe_mul(&d, &p->Z, &q->Z); fe_add(&d, &d, &d);
    fe_sub(&e, &b, &a); fe_sub(&f, &d, &c); fe_add(&g, &d, &c); fe_add(&h, &b, &a);
    fe_mul(&r->X, &e, &f); fe_mul(&r->Y, &g, &h); fe_mul(&r->Z, &f, &g); fe_mul(&r->T, &e, &h);
}
static void ge_sub(ge* r, const ge* p, const ge_cached* q) {
    ge_cached n; n.YpX = q->YmX; n.YmX = q->YpX; n.Z = q->Z; fe_neg(&n.T2d, &q->T2d);
    ge_add(r, p, &n);
}
static void ge_madd(ge* r, const ge* p, const ge_niels* q, int negate) {
    fe a, b, c, d, e, f, g, h, t;
    fe_sub(&t, &p->Y, &p->X); fe_mul(&a, &t, negate ? &q->ypx : &q->ymx);
    fe_add(&t, &p->Y, &p->X); fe_mul(&b, &t, negate ? &q->ymx : &q->ypx);
    fe_mul(&c, &p->T, &q->xy2d);
    if (negate) fe_neg(&c, &c);
    fe_add(&d, &p->Z, &p->Z);
    fe_sub(&e, &b, &a); fe_sub(&f, &d, &c); fe_add(&g, &d, &c); fe_add(&h, &b, &a);
    fe_mul(&r->X, &e, &f); fe_mul(&r->Y, &g, &h); fe_mul(&r->Z, &f, &g); fe_mul(&r->T, &e, &h);
}
static void ge_dbl(ge* r, const ge* p, int with_t) {
    fe xx, yy, zz2, s, cx, cy, cz, ct, t;
    fe_sq(&xx, &p->X); fe_sq(&yy, &p->Y); fe_sq(&zz2, &p->Z); fe_add(&zz2, &zz2, &zz2);
    fe_add(&t, &p->X, &p->Y); fe_sq(&s, &t);
    fe_add(&cy, &yy, &xx); fe_sub(&cz, &yy, &xx); fe_sub(&cx, &s, &cy); fe_sub(&ct, &zz2, &cz);
    fe_mul(&r->X, &cx, &ct); fe_mul(&r->Y, &cy, &cz); fe_mul(&r->Z, &cz, &ct);
    if (with_t) fe_mul(&r->T, &cx, &cy);
}
static int ge_eq(const ge* p, const ge* q) {
    fe a, b;
    fe_mul(&a, &p->X, &q->Y); fe_mul(&b, &p->Y, &q->X);
    int e1 = fe_eq(&a, &b);
    fe_mul(&a, &p->X, &q->X); fe_mul(&b, &p->Y, &q->Y);
    return e1 | fe_eq(&a, &b);
}
static int ge_is_identity(const ge* p) { return fe_iszero(&p->X) | fe_iszero(&p->Y); }

/* RFC 9496 4.3.1 */
static int ristretto_decode(ge* p, const uint8_t* b) {
    fe s, ss, u1, u2, u1s, u2s, v, I, dx, dy, t, one;
    uint8_t chk[32];
    fe_frombytes(&s, b); fe_tobytes(chk, &s);
    if (memcmp(chk, b, 32) != 0 || (b[0] & 1)) return 0;
    fe_1(&one);
    fe_sq(&ss, &s); fe_sub(&u1, &one, &ss); fe_add(&u2, &one, &ss); fe_sq(&u2s, &u2);
    fe_sq(&u1s, &u1); fe_mul(&t, &u1s, &C_D); fe_neg(&v, &t); fe_sub(&v, &v, &u2s);
    fe_mul(&t, &v, &u2s);
    int ok = fe_sqrt_ratio_i(&I, &one, &t);
    fe_mul(&dx, &I, &u2); fe_mul(&t, &I, &dx); fe_mul(&dy, &t, &v);
    fe_add(&t, &s, &s); fe_mul(&p->X, &t, &dx); fe_cneg(&p->X, fe_isneg(&p->X));
    fe_mul(&p->Y, &u1, &dy); fe_1(&p->Z); fe_mul(&p->T, &p->X, &p->Y);
    if (!ok || fe_isneg(&p->T) || fe_iszero(&p->Y)) return 0;
    return 1;
}
/* RFC 9496 4.3.2 */
static void ristretto_encode(uint8_t* out, const ge* p) {
    fe u1, u2, t, inv, i1, i2, zinv, deninv, ix, iy, ed, x, y, one;
    fe_1(&one);
    fe_add(&u1, &p->Z, &p->Y); fe_sub(&t, &p->Z, &p->Y); fe_mul(&u1, &u1, &t);
    fe_mul(&u2, &p->X, &p->Y);
    fe_sq(&t, &u2); fe_mul(&t, &u1, &t);
    fe_sqrt_ratio_i(&inv, &one, &t);
    fe_mul(&i1, &inv, &u1); fe_mul(&i2, &inv, &u2);
    fe_mul(&t, &i1, &i2); fe_mul(&zinv, &t, &p->T);
    fe_mul(&ix, &p->X, &C_SQRTM1); fe_mul(&iy, &p->Y, &C_SQRTM1); fe_mul(&ed, &i1, &C_INVSQRT_A_MINUS_D);
    fe_mul(&t, &p->T, &zinv);
    x = p->X; y = p->Y; deninv = i2;
    if (fe_isneg(&t)) { x = iy; y = ix; deninv = ed; }
    fe_mul(&t, &x, &zinv);
    fe_cneg(&y, fe_isneg(&t));
    fe_sub(&t, &p->Z, &y); fe_mul(&t, &t, &deninv);
    fe_cneg(&t, fe_isneg(&t));
    fe_tobytes(out, &t);
}

/* ------------------------------------------------------------------ scalars -------------------------------- */
static const uint8_t L_BYTES[32] = {0xed, 0xd3, 0xf5, 0x5c, 0x1a, 0x63, 0x12, 0x58, 0xd6, 0x9c, 0xf7, 0xa2, 0xde, 0xf9,
                                    0xde, 0x14, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x10};
static int sc_is_canonical(const uint8_t* s) {
    for (int i = 31; i >= 0; i--) {
        if (s[i] < L_BYTES[i]) return 1;
        if (s[i] > L_BYTES[i]) return 0;
    }
    return 0;
}
/* signed radix-16 digits in [-8, 8) (last may be 8)   [dalek Scalar::to_radix_16] */
static void sc_radix16(int8_t e[64], const uint8_t* s) {
    for (int i = 0; i < 32; i++) { e[2 * i] = s[i] & 15; e[2 * i + 1] = (s[i] >> 4) & 15; }
    int8_t carry = 0;
    for (int i = 0; i < 63; i++) {
        e[i] += carry;
        carry = (e[i] + 8) >> 4;
        e[i] -= carry << 4;
    }
    e[63] += carry;
}
/* signed radix-2^w digits [dalek Scalar::to_radix_2w]; returns digit count */
static int sc_radix2w(int16_t* d, const uint8_t* s, int w) {
    u64 x[5] = {0, 0, 0, 0, 0};
    memcpy(x, s, 32);
    int n = (256 + w - 1) / w;
    int carry = 0;
    for (int i = 0; i < n; i++) {
        int bit = i * w, wi = bit >> 6, sh = bit & 63;
        u64 raw = x[wi] >> sh;
        if (sh + w > 64 && wi < 4) raw |= x[wi + 1] << (64 - sh);
        int coef = (int)(raw & ((1u << w) - 1)) + carry;
        carry = (coef + (1 << (w - 1))) >> w;
        d[i] = (int16_t)(coef - (carry << w));
    }
    d[n] = (int16_t)carry;
    return n + 1;
}

/* variable base: 8-entry table, 64 x (4 dbl + 1 add)          [dalek variable_base::mul] */
static void ge_scalarmult(ge* r, const uint8_t* s, const ge* p) {
    ge_cached tbl[8];
    ge q = *p;
    ge_to_cached(&tbl[0], p);
    for (int i = 1; i < 8; i++) { ge_add(&q, &q, &tbl[0]); ge_to_cached(&tbl[i], &q); }
    int8_t e[64];
    sc_radix16(e, s);
    ge_identity(r);
    for (int i = 63; i >= 0; i--) {
        if (i != 63) { ge_dbl(r, r, 0); ge_dbl(r, r, 0); ge_dbl(r, r, 0); ge_dbl(r, r, 1); }
        if (e[i] > 0) ge_add(r, r, &tbl[e[i] - 1]);
        else if (e[i] < 0) ge_sub(r, r, &tbl[-e[i] - 1]);
    }
}
/* fixed base: table[i][j] = (j+1) * 16^i * Base, affine Niels; 64 mixed additions */
typedef struct { ge_niels t[64][8]; } fb_table;
static fb_table TBL_B, TBL_H;
static ge PT_B, PT_H;
static void to_niels(ge_niels* n, const ge* p) {
    fe zi, x, y, t;
    fe_invert(&zi, &p->Z); fe_mul(&x, &p->X, &zi); fe_mul(&y, &p->Y, &zi);
    fe_add(&n->ypx, &y, &x); fe_carry(&n->ypx); fe_sub(&n->ymx, &y, &x);
    fe_mul(&t, &x, &y); fe_mul(&n->xy2d, &t, &C_2D);
}
static void fb_build(fb_table* tb, const ge* base) {
    ge b = *base;
    for (int i = 0; i < 64; i++) {
        ge_cached c; ge_to_cached(&c, &b);
        ge q = b;
        for (int j = 0; j < 8; j++) { to_niels(&tb->t[i][j], &q); ge_add(&q, &q, &c); }
        for (int k = 0; k < 4; k++) ge_dbl(&b, &b, 1);
    }
}
static void ge_fixedmult(ge* r, const uint8_t* s, const fb_table* tb) {
    int8_t e[64];
    sc_radix16(e, s);
    ge_identity(r);
    for (int i = 0; i < 64; i++) {
        if (e[i] > 0) ge_madd(r, r, &tb->t[i][e[i] - 1], 0);
        else if (e[i] < 0) ge_madd(r, r, &tb->t[i][-e[i] - 1], 1);
    }
}

/* Straus, signed radix-16, shared doublings      [dalek straus.rs multiscalar_mul] */
static void msm_straus(ge* r, const uint8_t* scalars, const ge* pts, size_t n) {
    ge_cached* tbl = (ge_cached*)malloc(n * 8 * sizeof(ge_cached));
    int8_t* dig = (int8_t*)malloc(n * 64);
    for (size_t k = 0; k < n; k++) {
        ge q = pts[k];
        ge_to_cached(&tbl[8 * k], &pts[k]);
        for (int i = 1; i < 8; i++) { ge_add(&q, &q, &tbl[8 * k]); ge_to_cached(&tbl[8 * k + i], &q); }
        sc_radix16(dig + 64 * k, scalars + 32 * k);
    }
    ge_identity(r);
    for (int i = 63; i >= 0; i--) {
        if (i != 63) { ge_dbl(r, r, 0); ge_dbl(r, r, 0); ge_dbl(r, r, 0); ge_dbl(r, r, 1); }
        for (size_t k = 0; k < n; k++) {
            int8_t e = dig[64 * k + i];
            if (e > 0) ge_add(r, r, &tbl[8 * k + e - 1]);
            else if (e < 0) ge_sub(r, r, &tbl[8 * k - e - 1]);
        }
    }
    free(tbl); free(dig);
}
/* Pippenger, signed radix-2^w, 2^(w-1) buckets     [dalek pippenger.rs] */
static void msm_pippenger(ge* r, const uint8_t* scalars, const ge* pts, size_t n) {
    int w = n < 500 ? 6 : (n < 800 ? 7 : 8);
    int nd = (256 + w - 1) / w + 1;
    size_t nb = (size_t)1 << (w - 1);
    int16_t* dig = (int16_t*)malloc(n * nd * sizeof(int16_t));
    ge_cached* cp = (ge_cached*)malloc(n * sizeof(ge_cached));
    ge* buckets = (ge*)malloc(nb * sizeof(ge));
    for (size_t k = 0; k < n; k++) { sc_radix2w(dig + k * nd, scalars + 32 * k, w); ge_to_cached(&cp[k], &pts[k]); }
    ge_identity(r);
    for (int col = nd - 1; col >= 0; col--) {
        for (size_t b = 0; b < nb; b++) ge_identity(&buckets[b]);
        for (size_t k = 0; k < n; k++) {
            int d = dig[k * nd + col];
            if (d > 0) ge_add(&buckets[d - 1], &buckets[d - 1], &cp[k]);
            else if (d < 0) ge_sub(&buckets[-d - 1], &buckets[-d - 1], &cp[k]);
        }
        ge run = buckets[nb - 1], sum = buckets[nb - 1];
        for (size_t b = nb - 1; b-- > 0;) {
            ge_cached c;
            ge_to_cached(&c, &buckets[b]); ge_add(&run, &run, &c);
            ge_to_cached(&c, &run); ge_add(&sum, &sum, &c);
        }
        if (col != nd - 1) for (int i = 0; i < w; i++) ge_dbl(r, r, 1);
        ge_cached c; ge_to_cached(&c, &sum); ge_add(r, r, &c);
    }
    free(dig); free(cp); free(buckets);
}
static void msm_any(ge* r, const uint8_t* scalars, const ge* pts, size_t n) {
    if (n == 0) { ge_identity(r); return; }
    if (n < 190) msm_straus(r, scalars, pts, n); else msm_pippenger(r, scalars, pts, n);
}

/* ------------------------------------------------------------------ init ----------------------------------- */
static int g_init = 0;
static const uint8_t BASE_PK[64] = {
    0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
    0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76,
    0x8c, 0x92, 0x40, 0xb4, 0x56, 0xa9, 0xe6, 0xdc, 0x65, 0xc3, 0x77, 0xa1, 0x04, 0x8d, 0x74, 0x5f,
    0x94, 0xa0, 0x8c, 0xdb, 0x7f, 0x44, 0xcb, 0xcd, 0x7b, 0x46, 0xf3, 0x40, 0x48, 0x87, 0x11, 0x34};
int oq_init(void) {
    if (g_init) return 0;
    /* d = -121665/121666 ; sqrt(-1) = 2^((p-1)/4) ; 1/sqrt(a-d) -- derived, not transcribed */
    fe a, b, t;
    fe_from_u64(&a, 121665); fe_from_u64(&b, 121666);
    fe_invert(&t, &b); fe_mul(&t, &a, &t); fe_neg(&C_D, &t);
    fe_add(&C_2D, &C_D, &C_D); fe_carry(&C_2D);
    /* 2^((p-1)/4): (p-1)/4 = 2^253 - 5 ; compute 2^(2^253-5) = (2^(2^252-3))^2 * 2 */
    fe two; fe_from_u64(&two, 2);
    fe_pow22523(&t, &two); fe_sq(&t, &t); fe_mul(&C_SQRTM1, &t, &two);
    /* invsqrt(a - d) with a = -1 */
    fe one, amd; fe_1(&one); fe_neg(&amd, &one); fe_sub(&amd, &amd, &C_D);
    fe_sqrt_ratio_i(&C_INVSQRT_A_MINUS_D, &one, &amd);
    if (!ristretto_decode(&PT_B, BASE_PK) || !ristretto_decode(&PT_H, BASE_PK + 32)) return -1;
    fb_build(&TBL_B, &PT_B);
    fb_build(&TBL_H, &PT_H);
    g_init = 1;
    return 0;
}
int oq_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oq_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

/* ------------------------------------------------------------------ primitives for tests ------------------- */
int oq_decompress_check(const uint8_t* in) { ge p; return ristretto_decode(&p, in); }
int oq_roundtrip(uint8_t* out, const uint8_t* in) { ge p; int ok = ristretto_decode(&p, in); ristretto_encode(out, &p); return ok; }
int oq_scalarmult(uint8_t* out, const uint8_t* s, const uint8_t* pt) {
    ge p, r;
    if (!ristretto_decode(&p, pt)) return 1;
    ge_scalarmult(&r, s, &p);
    ristretto_encode(out, &r);
    return 0;
}

/* ------------------------------------------------------------------ batch entry points --------------------- */
enum { ST_OK = 0, ST_BAD_POINT = 1, ST_BAD_SCALAR = 2, ST_KEYPAIR = 3, ST_COMMIT = 4 };

int oq_fixed_base_batch(int which, const uint8_t* s, uint8_t* out, uint8_t* status, size_t n) {
    if (oq_init()) return -1;
#pragma omp parallel for schedule(dynamic, 64)
    for (size_t i = 0; i < n; i++) {
        memset(out + 32 * i, 0, 32);
        if (!sc_is_canonical(s + 32 * i)) { status[i] = ST_BAD_SCALAR; continue; }
        ge r;
        ge_fixedmult(&r, s + 32 * i, which ? &TBL_H : &TBL_B);
        ristretto_encode(out + 32 * i, &r);
        status[i] = ST_OK;
    }
    return 0;
}
/* (s*P0, s*P1) for 64-byte pairs: update_public_key and Mul for ElGamalCommitment */
int oq_update_public_key_batch(const uint8_t* pk, const uint8_t* s, uint8_t* out, uint8_t* status, size_t n) {
    if (oq_init()) return -1;
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < n; i++) {
        memset(out + 64 * i, 0, 64);
        if (!sc_is_canonical(s + 32 * i)) { status[i] = ST_BAD_SCALAR; continue; }
        ge p0, p1, r;
        if (!ristretto_decode(&p0, pk + 64 * i) || !ristretto_decode(&p1, pk + 64 * i + 32)) { status[i] = ST_BAD_POINT; continue; }
        ge_scalarmult(&r, s + 32 * i, &p0); ristretto_encode(out + 64 * i, &r);
        ge_scalarmult(&r, s + 32 * i, &p1); ristretto_encode(out + 64 * i + 32, &r);
        status[i] = ST_OK;
    }
    return 0;
}
static int commit_one(uint8_t* out, const uint8_t* pk, const uint8_t* r, const uint8_t* v) {
    if (!sc_is_canonical(r) || !sc_is_canonical(v)) return ST_BAD_SCALAR;
    ge gr, grsk, c, kh, gv;
    if (!ristretto_decode(&gr, pk) || !ristretto_decode(&grsk, pk + 32)) return ST_BAD_POINT;
    ge_scalarmult(&c, r, &gr);                 /* elgamal.rs:47 */
    ge_fixedmult(&gv, v, &TBL_B);              /* elgamal.rs:49 */
    ge_scalarmult(&kh, r, &grsk);              /* elgamal.rs:50 */
    ge_cached cc; ge_to_cached(&cc, &kh); ge_add(&gv, &gv, &cc);
    ristretto_encode(out, &c); ristretto_encode(out + 32, &gv);
    return ST_OK;
}
int oq_generate_commitment_batch(const uint8_t* pk, const uint8_t* r, const uint8_t* v, uint8_t* out, uint8_t* status, size_t n) {
    if (oq_init()) return -1;
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < n; i++) {
        memset(out + 64 * i, 0, 64);
        status[i] = (uint8_t)commit_one(out + 64 * i, pk + 64 * i, r + 32 * i, v + 32 * i);
        if (status[i]) memset(out + 64 * i, 0, 64);
    }
    return 0;
}
static int addc_one(uint8_t* out, const uint8_t* a, const uint8_t* b, int negate) {
    ge p[4];
    if (!ristretto_decode(&p[0], a) || !ristretto_decode(&p[1], a + 32) || !ristretto_decode(&p[2], b) || !ristretto_decode(&p[3], b + 32))
        return ST_BAD_POINT;
    for (int j = 0; j < 2; j++) {
        ge_cached c; ge r;
        ge_to_cached(&c, &p[2 + j]);
        if (negate) ge_sub(&r, &p[j], &c); else ge_add(&r, &p[j], &c);
        ristretto_encode(out + 32 * j, &r);
    }
    return ST_OK;
}
int oq_add_commitments_batch(const uint8_t* a, const uint8_t* b, int negate, uint8_t* out, uint8_t* status, size_t n) {
    if (oq_init()) return -1;
#pragma omp parallel for schedule(dynamic, 64)
    for (size_t i = 0; i < n; i++) {
        memset(out + 64 * i, 0, 64);
        status[i] = (uint8_t)addc_one(out + 64 * i, a + 64 * i, b + 64 * i, negate);
        if (status[i]) memset(out + 64 * i, 0, 64);
    }
    return 0;
}
/* Account::update_account, following the reference step by step (accounts.rs:149-153), including the
 * compress -> decompress round trip of new_comm inside add_commitments */
static int update_one(uint8_t* out, const uint8_t* acc, const uint8_t* bl, const uint8_t* u, const uint8_t* c) {
    if (!sc_is_canonical(bl) || !sc_is_canonical(u) || !sc_is_canonical(c)) return ST_BAD_SCALAR;
    /* 8 decodes per account, as the reference: 2 in update_public_key, 2 in generate_commitment, 4 in add_commitments
     * (an undecodable c / d is reported by addc_one) */
    ge gr, grsk, r;
    if (!ristretto_decode(&gr, acc) || !ristretto_decode(&grsk, acc + 32)) return ST_BAD_POINT;
    ge_scalarmult(&r, u, &gr); ristretto_encode(out, &r);
    ge_scalarmult(&r, u, &grsk); ristretto_encode(out + 32, &r);
    uint8_t newc[64];
    int st = commit_one(newc, acc, c, bl);
    if (st) return st;
    return addc_one(out + 64, newc, acc + 64, 0);
}
int oq_update_account_batch(const uint8_t* acc, const uint8_t* bl, const uint8_t* u, const uint8_t* c, uint8_t* out, uint8_t* status, size_t n) {
    if (oq_init()) return -1;
#pragma omp parallel for schedule(dynamic, 8)
    for (size_t i = 0; i < n; i++) {
        status[i] = (uint8_t)update_one(out + 128 * i, acc + 128 * i, bl + 32 * i, u + 32 * i, c + 32 * i);
        if (status[i]) memset(out + 128 * i, 0, 128);
    }
    return 0;
}
static int verify_one(const uint8_t* acc, const uint8_t* sk, const uint8_t* bl) {
    if (!sc_is_canonical(sk) || !sc_is_canonical(bl)) return ST_BAD_SCALAR;
    ge gr, c, r, gv;
    uint8_t enc[32];
    if (!ristretto_decode(&gr, acc)) return ST_BAD_POINT;
    ge_scalarmult(&r, sk, &gr); ristretto_encode(enc, &r);
    if (memcmp(enc, acc + 32, 32) != 0) return ST_KEYPAIR;
    if (!ristretto_decode(&c, acc + 64)) return ST_BAD_POINT;
    ge_fixedmult(&gv, bl, &TBL_B); ge_scalarmult(&r, sk, &c);
    ge_cached cc; ge_to_cached(&cc, &r); ge_add(&gv, &gv, &cc);
    ristretto_encode(enc, &gv);
    return memcmp(enc, acc + 96, 32) == 0 ? ST_OK : ST_COMMIT;
}
int oq_verify_account_batch(const uint8_t* acc, const uint8_t* sk, const uint8_t* bl, uint8_t* status, size_t n) {
    if (oq_init()) return -1;
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < n; i++) status[i] = (uint8_t)verify_one(acc + 128 * i, sk + 32 * i, bl + 32 * i);
    return 0;
}
int oq_verify_public_key_update_batch(const uint8_t* upd, const uint8_t* pk, const uint8_t* r, uint8_t* status, size_t n) {
    if (oq_init()) return -1;
#pragma omp parallel for schedule(dynamic, 16)
    for (size_t i = 0; i < n; i++) {
        if (!sc_is_canonical(r + 32 * i)) { status[i] = ST_BAD_SCALAR; continue; }
        ge p0, p1, u0, u1, a, b;
        if (!ristretto_decode(&p0, pk + 64 * i) || !ristretto_decode(&p1, pk + 64 * i + 32) ||
            !ristretto_decode(&u0, upd + 64 * i) || !ristretto_decode(&u1, upd + 64 * i + 32)) { status[i] = ST_BAD_POINT; continue; }
        ge_scalarmult(&a, r + 32 * i, &p0); ge_scalarmult(&b, r + 32 * i, &p1);
        status[i] = (ge_eq(&a, &u0) && ge_eq(&b, &u1)) ? ST_OK : ST_KEYPAIR;
    }
    return 0;
}
/* create_delta_and_epsilon_accounts with caller-supplied r; the epsilon half follows the reference literally:
 * generate_commitment(base_pk, r, bl) with variable-base mults on the decompressed constants (accounts.rs:214) */
int oq_delta_epsilon_batch(const uint8_t* acc, const uint8_t* bl, const uint8_t* r, const uint8_t* base_pk, uint8_t* delta, uint8_t* eps, uint8_t* status, size_t n) {
    if (oq_init()) return -1;
#pragma omp parallel for schedule(dynamic, 8)
    for (size_t i = 0; i < n; i++) {
        uint8_t cd[64], ce[64];
        int st = commit_one(cd, acc + 128 * i, r + 32 * i, bl + 32 * i);
        if (!st) st = commit_one(ce, base_pk, r + 32 * i, bl + 32 * i);
        status[i] = (uint8_t)st;
        if (st) { memset(delta + 128 * i, 0, 128); memset(eps + 128 * i, 0, 128); continue; }
        memcpy(delta + 128 * i, acc + 128 * i, 64); memcpy(delta + 128 * i + 64, cd, 64);
        memcpy(eps + 128 * i, base_pk, 64); memcpy(eps + 128 * i + 64, ce, 64);
    }
    return 0;
}
/* optional_multiscalar_mul over compressed points; threads split the terms and add their partial sums */
int oq_msm(const uint8_t* scalars, const uint8_t* points, size_t n, uint8_t* out, uint8_t* status) {
    if (oq_init()) return -1;
    memset(out, 0, 32);
    ge* pts = (ge*)malloc((n ? n : 1) * sizeof(ge));
    size_t first_bad = (size_t)-1;
    int bad_code = 0;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        int code = 0;
        if (!sc_is_canonical(scalars + 32 * i)) code = ST_BAD_SCALAR;
        else if (!ristretto_decode(&pts[i], points + 32 * i)) code = ST_BAD_POINT;
        if (code) {
#pragma omp critical
            { if (i < first_bad) { first_bad = i; bad_code = code; } }
        }
    }
    if (bad_code) { *status = (uint8_t)bad_code; free(pts); return 0; }
    int nt = oq_threads();
    if ((size_t)nt > n / 256 + 1) nt = (int)(n / 256 + 1);
    ge* part = (ge*)malloc(nt * sizeof(ge));
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int t = 0; t < nt; t++) {
        size_t lo = n * t / nt, hi = n * (t + 1) / nt;
        msm_any(&part[t], scalars + 32 * lo, pts + lo, hi - lo);
    }
    ge acc = part[0];
    for (int t = 1; t < nt; t++) { ge_cached c; ge_to_cached(&c, &part[t]); ge_add(&acc, &acc, &c); }
    ristretto_encode(out, &acc);
    *status = ST_OK;
    free(part); free(pts);
    return 0;
}
int oq_msm_segmented(const uint8_t* scalars, const uint8_t* points, const uint32_t* offsets, size_t m, uint8_t* out, uint8_t* status) {
    if (oq_init()) return -1;
#pragma omp parallel for schedule(dynamic, 8)
    for (size_t j = 0; j < m; j++) {
        size_t lo = offsets[j], k = offsets[j + 1] - lo;
        memset(out + 32 * j, 0, 32);
        ge* pts = (ge*)malloc((k ? k : 1) * sizeof(ge));
        int code = 0;
        for (size_t i = 0; i < k && !code; i++) {
            if (!sc_is_canonical(scalars + 32 * (lo + i))) code = ST_BAD_SCALAR;
            else if (!ristretto_decode(&pts[i], points + 32 * (lo + i))) code = ST_BAD_POINT;
        }
        if (!code) { ge r; msm_any(&r, scalars + 32 * lo, pts, k); ristretto_encode(out + 32 * j, &r); }
        status[j] = (uint8_t)code;
        free(pts);
    }
    return 0;
}
int oq_delta_identity_check(const uint8_t* acc, size_t n, uint8_t* verdict) {
    if (oq_init()) return -1;
    ge sc, sd;
    ge_identity(&sc); ge_identity(&sd);
    for (size_t i = 0; i < n; i++) {
        ge c, d; ge_cached t;
        if (!ristretto_decode(&c, acc + 128 * i + 64) || !ristretto_decode(&d, acc + 128 * i + 96)) { *verdict = ST_BAD_POINT; return 0; }
        ge_to_cached(&t, &c); ge_add(&sc, &sc, &t);
        ge_to_cached(&t, &d); ge_add(&sd, &sd, &t);
    }
    *verdict = (ge_is_identity(&sc) && ge_is_identity(&sd)) ? ST_OK : ST_COMMIT;
    return 0;
}
