// GF(2^255-19) arithmetic for sm_100a -- saturated representation.
//
// Representation: 8 unsigned 32-bit limbs, little-endian, value in [0, 2^256) (NOT necessarily reduced mod p;
// 2^256 = 38 mod p).  Every function accepts and returns the full range, so there are no magnitude budgets to
// track at the call sites.
//
// Why saturated limbs: the multiply pipe is the bound of every kernel here.  On sm_100a a
// 32x32->64 product is one IMAD.WIDE.U32 (measured 0.7-0.8e13 thread-ops/s per GPU, ~2.6x slower than a plain
// IMAD), and nothing else on the SM multiplies integers faster.  The radix-2^25.5 form needs 100 products per
// multiplication; this form needs 64 plus 8 products by the constant 38 for the reduction: 72 IMAD.WIDE per
// multiplication, 44 per squaring (one level of subtractive Karatsuba, 48 + 8, is kept as fe_mul_inl but measures
// slower in the kernels because of its carry-chain glue).  Accumulation is free:
// each product is a (mad.lo.cc, madc.hi.cc) PTX pair that ptxas fuses into one IMAD.WIDE.U32.X with the carry in a
// predicate register (verified in SASS), see tools/gen_field_ops.py for the even/odd accumulator scheme.
// Additions, subtractions and the Karatsuba glue are carry chains on the ALU pipe (IADD3.X), which has 5x the
// throughput and runs beside the multiply pipe.
//
// Replaces (for the hot path) curve25519-dalek 3.x `backend/serial/u64/field.rs` + `field.rs`
// (FieldElement51::{mul,square,pow_p58,sqrt_ratio_i,to_bytes,from_bytes}); dalek is a dependency of the
// reference (Cargo.toml:42) and not vendored, so this restates the published arithmetic (RFC 9496 / RFC 7748).
//
// The file compiles for the host too (carry flag emulated in a thread-local) so tests/ can unit-test the exact
// device algorithms on CPU against the oracle; the host build is test infrastructure, never linked into the library.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QQ_HD __host__ __device__ __forceinline__
#define QQ_D __device__ __forceinline__
#else
#define QQ_HD inline
#define QQ_D inline
#endif

namespace qq {

typedef uint32_t u32;
typedef uint64_t u64;

#define QQ_FE_LIMBS 8

struct fe {
    u32 v[QQ_FE_LIMBS];
};

// ---------------------------------------------------------------------------------------------------------
// Carry-flag primitives.  Device: PTX extended-precision instructions (the flag lives in CC.CF / a predicate).
// Host: the same semantics with the flag in a thread-local, so the host unit tests run the identical algorithm.
// ---------------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define QQ_ASM2(name, ins)                                                  \
    QQ_HD u32 name(u32 a, u32 b) {                                          \
        u32 r;                                                              \
        asm volatile(ins " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));        \
        return r;                                                           \
    }
#define QQ_ASM3(name, ins)                                                          \
    QQ_HD u32 name(u32 a, u32 b, u32 c) {                                           \
        u32 r;                                                                      \
        asm volatile(ins " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));    \
        return r;                                                                   \
    }
QQ_ASM2(add_cc, "add.cc.u32")
QQ_ASM2(addc_cc, "addc.cc.u32")
QQ_ASM2(addc, "addc.u32")
QQ_ASM2(sub_cc, "sub.cc.u32")
QQ_ASM2(subc_cc, "subc.cc.u32")
QQ_ASM2(subc, "subc.u32")
QQ_ASM2(mul_lo, "mul.lo.u32")
QQ_ASM2(mul_hi, "mul.hi.u32")
QQ_ASM3(mad_lo_cc, "mad.lo.cc.u32")
QQ_ASM3(madc_lo_cc, "madc.lo.cc.u32")
QQ_ASM3(madc_hi_cc, "madc.hi.cc.u32")
QQ_ASM3(madc_hi, "madc.hi.u32")
#undef QQ_ASM2
#undef QQ_ASM3
#else
static thread_local u32 qq_cf = 0;
inline u32 add_cc(u32 a, u32 b) { u64 s = (u64)a + b; qq_cf = (u32)(s >> 32); return (u32)s; }
inline u32 addc_cc(u32 a, u32 b) { u64 s = (u64)a + b + qq_cf; qq_cf = (u32)(s >> 32); return (u32)s; }
inline u32 addc(u32 a, u32 b) { return a + b + qq_cf; }
inline u32 sub_cc(u32 a, u32 b) { u64 s = (u64)a - b; qq_cf = (u32)(s >> 63); return (u32)s; }
inline u32 subc_cc(u32 a, u32 b) { u64 s = (u64)a - b - qq_cf; qq_cf = (u32)(s >> 63); return (u32)s; }
inline u32 subc(u32 a, u32 b) { return a - b - qq_cf; }
inline u32 mul_lo(u32 a, u32 b) { return (u32)((u64)a * b); }
inline u32 mul_hi(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
inline u32 mad_lo_cc(u32 a, u32 b, u32 c) { return add_cc(mul_lo(a, b), c); }
inline u32 madc_lo_cc(u32 a, u32 b, u32 c) { return addc_cc(mul_lo(a, b), c); }
inline u32 madc_hi_cc(u32 a, u32 b, u32 c) { return addc_cc(mul_hi(a, b), c); }
inline u32 madc_hi(u32 a, u32 b, u32 c) { return addc(mul_hi(a, b), c); }
#endif

#include "fe25519_mp.inc"

QQ_HD void fe_0(fe& h) {
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] = 0;
}
QQ_HD void fe_1(fe& h) {
    h.v[0] = 1;
#pragma unroll
    for (int i = 1; i < 8; i++) h.v[i] = 0;
}

// h = f + g.  Two folds of the carry (2^256 = 38): after the first the value can wrap once more only if it is
// within 38 of 2^256, in which case the wrapped value is < 38 and the last correction cannot carry.
QQ_HD void fe_add(fe& h, const fe& f, const fe& g) {
    u32 r[8];
    r[0] = add_cc(f.v[0], g.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r[i] = addc_cc(f.v[i], g.v[i]);
    u32 k = addc(0u, 0u);
    r[0] = add_cc(r[0], (0u - k) & 38u);
#pragma unroll
    for (int i = 1; i < 8; i++) r[i] = addc_cc(r[i], 0u);
    k = addc(0u, 0u);
    r[0] += (0u - k) & 38u;
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] = r[i];
}
// h = f - g (mod p)
QQ_HD void fe_sub(fe& h, const fe& f, const fe& g) {
    u32 r[8];
    r[0] = sub_cc(f.v[0], g.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r[i] = subc_cc(f.v[i], g.v[i]);
    u32 m = subc(0u, 0u);  // all-ones when the subtraction borrowed
    r[0] = sub_cc(r[0], m & 38u);
#pragma unroll
    for (int i = 1; i < 8; i++) r[i] = subc_cc(r[i], 0u);
    m = subc(0u, 0u);
    r[0] -= m & 38u;
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] = r[i];
}
QQ_HD void fe_sub4(fe& h, const fe& f, const fe& g) { fe_sub(h, f, g); }
QQ_HD void fe_neg(fe& h, const fe& f) {
    // 2p - f = (2^256 - 38) - f  >= 0 whenever f <= 2^256 - 38; the borrow path covers the last 37 values
    fe z;
    fe_0(z);
    fe_sub(h, z, f);
}
// identity in this representation (kept so that the group-law code reads the same as before)
QQ_HD void fe_carry(fe& h, const fe& f) { h = f; }

// 16 limbs -> 8 limbs: h = t[0..7] + 38 * t[8..15] (mod p), folded into [0, 2^256).
QQ_HD void fe_reduce512(fe& h, const u32* t) {
    u32 E[8], O[8], r[8];
    // even positions accumulate on top of the low half, odd positions are fresh
    E[0] = mad_lo_cc(t[8], 38u, t[0]);
    E[1] = madc_hi_cc(t[8], 38u, t[1]);
    E[2] = madc_lo_cc(t[10], 38u, t[2]);
    E[3] = madc_hi_cc(t[10], 38u, t[3]);
    E[4] = madc_lo_cc(t[12], 38u, t[4]);
    E[5] = madc_hi_cc(t[12], 38u, t[5]);
    E[6] = madc_lo_cc(t[14], 38u, t[6]);
    E[7] = madc_hi_cc(t[14], 38u, t[7]);
    u32 ce = addc(0u, 0u);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        O[2 * k] = mul_lo(t[9 + 2 * k], 38u);
        O[2 * k + 1] = mul_hi(t[9 + 2 * k], 38u);
    }
    r[0] = E[0];
    r[1] = add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 8; k++) r[k] = addc_cc(E[k], O[k - 1]);
    u32 top = addc(ce, O[7]);  // total < 39 * 2^256  =>  top <= 38
    r[0] = add_cc(r[0], top * 38u);
#pragma unroll
    for (int k = 1; k < 8; k++) r[k] = addc_cc(r[k], 0u);
    u32 k2 = addc(0u, 0u);
    r[0] += (0u - k2) & 38u;  // wrapped value is < 38*39, cannot carry
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] = r[i];
}

// |x - y| on 4 limbs; returns the all-ones mask when x < y.
QQ_HD u32 mp_absdiff4(u32* d, const u32* x, const u32* y) {
    u32 t[4];
    t[0] = sub_cc(x[0], y[0]);
    t[1] = subc_cc(x[1], y[1]);
    t[2] = subc_cc(x[2], y[2]);
    t[3] = subc_cc(x[3], y[3]);
    u32 m = subc(0u, 0u);
    d[0] = sub_cc(t[0] ^ m, m);
    d[1] = subc_cc(t[1] ^ m, m);
    d[2] = subc_cc(t[2] ^ m, m);
    d[3] = subc(t[3] ^ m, m);
    return m;
}

// h = f * g: one level of subtractive Karatsuba over 128-bit halves (3 x 16 products) + reduction (8 products).
//   f*g = z0 + (z0 + z2 + (f0 - f1)(g1 - g0)) 2^128 + z2 2^256
QQ_HD void fe_mul_inl(fe& h, const fe& f, const fe& g) {
    u32 t[16], zm[8], df[4], dg[4], z1[9];
    mp_mul4(t, f.v, g.v);
    mp_mul4(t + 8, f.v + 4, g.v + 4);
    u32 sf = mp_absdiff4(df, f.v, f.v + 4);
    u32 sg = mp_absdiff4(dg, g.v + 4, g.v);
    mp_mul4(zm, df, dg);
    u32 s = sf ^ sg;  // all-ones: the middle product is negative
    z1[0] = add_cc(t[0], t[8]);
#pragma unroll
    for (int i = 1; i < 8; i++) z1[i] = addc_cc(t[i], t[8 + i]);
    z1[8] = addc(0u, 0u);
    // z1 += s ? -zm : zm   (two's complement over 9 limbs; carry-in = s & 1)
    add_cc(s, s);
#pragma unroll
    for (int i = 0; i < 8; i++) z1[i] = addc_cc(z1[i], zm[i] ^ s);
    z1[8] = addc(z1[8], s);
    t[4] = add_cc(t[4], z1[0]);
#pragma unroll
    for (int i = 1; i < 9; i++) t[4 + i] = addc_cc(t[4 + i], z1[i]);
    t[13] = addc_cc(t[13], 0u);
    t[14] = addc_cc(t[14], 0u);
    t[15] = addc(t[15], 0u);
    fe_reduce512(h, t);
}
// schoolbook variant (64 + 8 products), kept for the bake-off in tools/fe_bench.cu
QQ_HD void fe_mul_school(fe& h, const fe& f, const fe& g) {
    u32 t[16];
    mp_mul8(t, f.v, g.v);
    fe_reduce512(h, t);
}

// t[2n] = 2 * off[2n] + diag(a)   (n = 8)
QQ_HD void fe_sq_inl(fe& h, const fe& f) {
    u32 off[16], t[16];
    mp_sqoff8(off, f.v);
    // double: funnel shifts (ALU pipe)
    u32 d[16];
    d[0] = 0;
#pragma unroll
    for (int k = 1; k < 16; k++) {
#if defined(__CUDA_ARCH__)
        asm("shf.l.wrap.b32 %0, %1, %2, 1;" : "=r"(d[k]) : "r"(off[k - 1]), "r"(off[k]));
#else
        d[k] = (off[k] << 1) | (off[k - 1] >> 31);
#endif
    }
    // add the diagonal a[i]^2 at limbs (2i, 2i+1)
    t[0] = mad_lo_cc(f.v[0], f.v[0], d[0]);
    t[1] = madc_hi_cc(f.v[0], f.v[0], d[1]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
        t[2 * i] = madc_lo_cc(f.v[i], f.v[i], d[2 * i]);
        t[2 * i + 1] = (i == 7) ? madc_hi(f.v[i], f.v[i], d[2 * i + 1]) : madc_hi_cc(f.v[i], f.v[i], d[2 * i + 1]);
    }
    fe_reduce512(h, t);
}

#if defined(__CUDACC__) && !defined(QQ_INLINE_FIELD_OPS)
static __device__ __noinline__ fe fe_mul_ool(fe f, fe g) {
    fe h;
#if defined(QQ_FE_MUL_KARATSUBA)
    fe_mul_inl(h, f, g);
#else
    // schoolbook (64 + 8 products) measured 4 % faster than Karatsuba (48 + 8) inside k_varbase_split: the Karatsuba
    // glue is ~40 extra carry-chain instructions that issue beside the multiplies (tools/vb_bench.cu, profiles/)
    fe_mul_school(h, f, g);
#endif
    return h;
}
static __device__ __noinline__ fe fe_sq_ool(fe f) {
    fe h;
    fe_sq_inl(h, f);
    return h;
}
static __device__ __noinline__ fe fe_sqn_ool(fe f, int n) {
#pragma unroll 1
    for (int i = 0; i < n; i++) {
        fe h;
        fe_sq_inl(h, f);
        f = h;
    }
    return f;
}
#endif
#if defined(__CUDACC__) && !defined(QQ_INLINE_FIELD_OPS)
// Several INDEPENDENT products per out-of-line call.  The group law offers them in fours (addition: 4 + 4 products,
// doubling: 4 squarings + 3..4 products); inside one function body ptxas interleaves the four carry chains, and the
// call marshalling is paid once per group of products.  Measured in k_varbase_split: 5.17e7 -> 5.48e7 scalar-mults/s
// against one product per call (tools/vb_bench.cu, profiles/).  Arguments and results travel in registers.
struct fe2 {
    fe a, b;
};
struct fe3 {
    fe a, b, c;
};
struct fe4 {
    fe a, b, c, d;
};
static __device__ __noinline__ fe3 fe_mul3_ool(fe f0, fe g0, fe f1, fe g1, fe f2, fe g2) {
    fe3 r;
    fe_mul_school(r.a, f0, g0);
    fe_mul_school(r.b, f1, g1);
    fe_mul_school(r.c, f2, g2);
    return r;
}
static __device__ __noinline__ fe4 fe_mul4_ool(fe f0, fe g0, fe f1, fe g1, fe f2, fe g2, fe f3, fe g3) {
    fe4 r;
    fe_mul_school(r.a, f0, g0);
    fe_mul_school(r.b, f1, g1);
    fe_mul_school(r.c, f2, g2);
    fe_mul_school(r.d, f3, g3);
    return r;
}
static __device__ __noinline__ fe4 fe_sq4_ool(fe f0, fe f1, fe f2, fe f3) {
    fe4 r;
    fe_sq_inl(r.a, f0);
    fe_sq_inl(r.b, f1);
    fe_sq_inl(r.c, f2);
    fe_sq_inl(r.d, f3);
    return r;
}
#endif
QQ_HD void fe_mul(fe& h, const fe& f, const fe& g) {
#if defined(__CUDA_ARCH__) && !defined(QQ_INLINE_FIELD_OPS)
    h = fe_mul_ool(f, g);
#else
    fe_mul_inl(h, f, g);
#endif
}
QQ_HD void fe_sq(fe& h, const fe& f) {
#if defined(__CUDA_ARCH__) && !defined(QQ_INLINE_FIELD_OPS)
    h = fe_sq_ool(f);
#else
    fe_sq_inl(h, f);
#endif
}
QQ_HD void fe_sqn(fe& h, const fe& f, int n) {
#if defined(__CUDA_ARCH__) && !defined(QQ_INLINE_FIELD_OPS)
    h = fe_sqn_ool(f, n);
#else
    fe_sq(h, f);
    for (int i = 1; i < n; i++) fe_sq(h, h);
#endif
}

// z^(2^252 - 3) = z^((p-5)/8)   (dalek field.rs pow_p58 / ref10 pow22523 addition chain: 251 S + 11 M)
QQ_HD void fe_pow22523(fe& out, const fe& z) {
    fe t0, t1, t2;
    fe_sq(t0, z);              // 2
    fe_sqn(t1, t0, 2);         // 8
    fe_mul(t1, z, t1);         // 9
    fe_mul(t0, t0, t1);        // 11
    fe_sq(t0, t0);             // 22
    fe_mul(t0, t1, t0);        // 31 = 2^5-1
    fe_sqn(t1, t0, 5);
    fe_mul(t0, t1, t0);        // 2^10-1
    fe_sqn(t1, t0, 10);
    fe_mul(t1, t1, t0);        // 2^20-1
    fe_sqn(t2, t1, 20);
    fe_mul(t1, t2, t1);        // 2^40-1
    fe_sqn(t1, t1, 10);
    fe_mul(t0, t1, t0);        // 2^50-1
    fe_sqn(t1, t0, 50);
    fe_mul(t1, t1, t0);        // 2^100-1
    fe_sqn(t2, t1, 100);
    fe_mul(t1, t2, t1);        // 2^200-1
    fe_sqn(t1, t1, 50);
    fe_mul(t0, t1, t0);        // 2^250-1
    fe_sqn(t0, t0, 2);         // 2^252-4
    fe_mul(out, t0, z);        // 2^252-3
}

// Canonical little-endian bytes as 8 x u32 words (fully reduced mod p).  Any input in [0, 2^256).
QQ_HD void fe_towords(u32 w[8], const fe& f) {
    u32 r[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = f.v[i];
    // fold bit 255 twice (2^255 = 19): after the first fold r < 2^255 + 19, after the second r < 2^255
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        u32 top = r[7] >> 31;
        r[7] &= 0x7fffffffu;
        r[0] = add_cc(r[0], (0u - top) & 19u);
#pragma unroll
        for (int i = 1; i < 7; i++) r[i] = addc_cc(r[i], 0u);
        r[7] = addc(r[7], 0u);
    }
    // r in [0, 2^255): r >= p  <=>  r + 19 >= 2^255
    q[0] = add_cc(r[0], 19u);
#pragma unroll
    for (int i = 1; i < 7; i++) q[i] = addc_cc(r[i], 0u);
    q[7] = addc(r[7], 0u);
    u32 ge = 0u - (q[7] >> 31);  // all-ones: take r - p = (r + 19) - 2^255
    q[7] &= 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = (q[i] & ge) | (r[i] & ~ge);
}

// Little-endian 8 x u32 words -> field element; bit 255 is ignored (dalek FieldElement::from_bytes behaviour).
QQ_HD void fe_fromwords(fe& h, const u32 w[8]) {
#pragma unroll
    for (int i = 0; i < 7; i++) h.v[i] = w[i];
    h.v[7] = w[7] & 0x7fffffffu;
}

QQ_HD u32 fe_isnegative(const fe& f) {
    u32 w[8];
    fe_towords(w, f);
    return w[0] & 1u;
}
QQ_HD u32 fe_iszero(const fe& f) {
    u32 w[8];
    fe_towords(w, f);
    u32 r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r |= w[i];
    return r == 0 ? 1u : 0u;
}
QQ_HD u32 fe_eq(const fe& f, const fe& g) {
    fe d;
    fe_sub(d, f, g);
    return fe_iszero(d);
}
// h = b ? g : h     (b in {0,1}; branch-free)
QQ_HD void fe_cmov(fe& h, const fe& g, u32 b) {
    u32 m = 0u - b;
#pragma unroll
    for (int i = 0; i < 8; i++) h.v[i] ^= m & (h.v[i] ^ g.v[i]);
}
// h = b ? -h : h
QQ_HD void fe_cneg(fe& h, u32 b) {
    fe n;
    fe_neg(n, h);
    fe_cmov(h, n, b);
}
QQ_HD void fe_abs(fe& h) { fe_cneg(h, fe_isnegative(h)); }

// ---- constants (generated by tools/gen_consts.py; checked numerically by tests/test_host_arith.py) ----
#define QQ_FE_CONST(name, a0, a1, a2, a3, a4, a5, a6, a7) \
    QQ_HD fe name() {                                     \
        fe r = {{a0, a1, a2, a3, a4, a5, a6, a7}};        \
        return r;                                         \
    }
#include "fe25519_consts.inc"

// sqrt_ratio_i (RFC 9496 4.2; dalek field.rs sqrt_ratio_i): returns was_square, r = sqrt(u/v) or sqrt(i*u/v), r >= 0.
// Branch-free (constant-time as written).
QQ_HD u32 fe_sqrt_ratio_i(fe& r, const fe& u, const fe& v) {
    fe v3, v7, t, check, uneg, unegi;
    fe_sq(v3, v);
    fe_mul(v3, v3, v);      // v^3
    fe_sq(v7, v3);
    fe_mul(v7, v7, v);      // v^7
    fe_mul(t, u, v7);       // u v^7
    fe_pow22523(t, t);
    fe_mul(r, u, v3);
    fe_mul(r, r, t);        // r = u v^3 (u v^7)^((p-5)/8)
    fe_sq(check, r);
    fe_mul(check, v, check);  // v r^2
    fe_neg(uneg, u);
    fe_mul(unegi, uneg, fe_sqrt_m1());
    u32 correct = fe_eq(check, u);
    u32 flipped = fe_eq(check, uneg);
    u32 flipped_i = fe_eq(check, unegi);
    fe ri;
    fe_mul(ri, r, fe_sqrt_m1());
    fe_cmov(r, ri, flipped | flipped_i);
    fe_abs(r);
    return correct | flipped;
}

}  // namespace qq
