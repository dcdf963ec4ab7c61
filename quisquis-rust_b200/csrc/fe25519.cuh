// GF(2^255-19) arithmetic for sm_100a.
//
// Representation: 10 unsigned limbs, radix 2^25.5 (limb i has 26 bits for even i, 25 bits for odd i), one limb per
// 32-bit register.  Every 32x32->64 partial product is one IMAD.WIDE.U32; the 19-fold for 2^255 = 19 is applied to
// one operand before the products, so a multiplication is 100 wide products + 9 small IMADs and needs NO carry
// propagation between partial products (all column sums stay below 2^64).  ptxas pairs the products into
// 3-input 64-bit adds (IADD3 + IADD3.X on the ALU pipe), which balances the FMA-pipe and ALU-pipe issue slots.
//
// Replaces (for the hot path) curve25519-dalek 3.x `backend/serial/u64/field.rs` + `field.rs`
// (FieldElement51::{mul,square,pow_p58,sqrt_ratio_i,to_bytes,from_bytes}); dalek is a dependency of the
// reference (Cargo.toml:42) and not vendored, so this restates the published arithmetic (RFC 9496 / RFC 7748).
//
// Bounds (T = "tight" = output of fe_mul/fe_sq/fe_carry): even limbs < 2^26, odd limbs < 2^25 + 2^18.
//   fe_add(tight,tight)        -> <= 2T   ("loose")
//   fe_sub(any<=3T, tight)     -> a + 2p - b
//   fe_mul(f,g): g limbs must satisfy 19*g < 2^32 (g <= 3.3T); f*g magnitude product <= ~30 T^2 per limb pair.
//   fe_sq(f):   f <= 3.3T.
// All call sites in ge25519.cuh / ristretto.cuh are annotated with the bound they rely on.
//
// The file compiles for the host too (QQ_HD empty, plain C multiply) so tests/ can unit-test the exact device
// arithmetic on CPU against the oracle; the host build is test infrastructure and is never linked into the library.
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

struct fe {
    u32 v[10];
};

#define QQ_M26 0x3ffffffu
#define QQ_M25 0x1ffffffu

QQ_HD u64 mul_wide(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
#else
    return (u64)a * b;
#endif
}
QQ_HD u64 mad_wide(u32 a, u32 b, u64 c) {
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
#else
    return (u64)a * b + c;
#endif
}

// x << K for x < 2^(32-K), written as a rotate so that ptxas keeps it on the ALU pipe (SHF) instead of emitting
// IMAD.SHL / IMAD.IADD on the FMA pipe, which is the bottleneck pipe of every kernel here.
template <int K>
QQ_HD u32 shl_alu(u32 x) {
#if defined(__CUDA_ARCH__)
    u32 r;
    asm("shf.l.wrap.b32 %0, %1, %1, %2;" : "=r"(r) : "r"(x), "n"(K));
    return r;
#else
    return x << K;
#endif
}

QQ_HD void fe_0(fe& h) {
#pragma unroll
    for (int i = 0; i < 10; i++) h.v[i] = 0;
}
QQ_HD void fe_1(fe& h) {
    h.v[0] = 1;
#pragma unroll
    for (int i = 1; i < 10; i++) h.v[i] = 0;
}
QQ_HD void fe_add(fe& h, const fe& f, const fe& g) {
#pragma unroll
    for (int i = 0; i < 10; i++) h.v[i] = f.v[i] + g.v[i];
}
// 2p in this radix
QQ_HD u32 fe_2p_limb(int i) { return i == 0 ? 0x7ffffdau : ((i & 1) ? 0x3fffffeu : 0x7fffffeu); }
// h = f - g (mod p), computed as f + 2p - g.  g must be tight (limb-wise <= 2p).
QQ_HD void fe_sub(fe& h, const fe& f, const fe& g) {
#pragma unroll
    for (int i = 0; i < 10; i++) h.v[i] = f.v[i] + fe_2p_limb(i) - g.v[i];
}
// h = f - g with g up to 2T (limb-wise <= 4p): f + 4p - g
QQ_HD void fe_sub4(fe& h, const fe& f, const fe& g) {
#pragma unroll
    for (int i = 0; i < 10; i++) h.v[i] = f.v[i] + 2u * fe_2p_limb(i) - g.v[i];
}
QQ_HD void fe_neg(fe& h, const fe& f) {
#pragma unroll
    for (int i = 0; i < 10; i++) h.v[i] = fe_2p_limb(i) - f.v[i];
}

// Parallel (single-step) weak reduction: every limb hands its excess to the next limb at once.
// Input limbs < 2^31; output is tight.  ~31 ALU ops, no serial chain.
QQ_HD void fe_carry(fe& h, const fe& f) {
    u32 c[10];
#pragma unroll
    for (int i = 0; i < 10; i++) c[i] = (i & 1) ? (f.v[i] >> 25) : (f.v[i] >> 26);
    h.v[0] = (f.v[0] & QQ_M26) + 19u * c[9];
#pragma unroll
    for (int i = 1; i < 10; i++) h.v[i] = (f.v[i] & ((i & 1) ? QQ_M25 : QQ_M26)) + c[i - 1];
}

// Carry chain on ten 64-bit column sums -> tight limbs.
QQ_HD void fe_reduce64(fe& h, u64 t[10]) {
    u64 c;
    c = t[0] >> 26; t[1] += c; t[0] &= QQ_M26;
    c = t[4] >> 26; t[5] += c; t[4] &= QQ_M26;
    c = t[1] >> 25; t[2] += c; t[1] &= QQ_M25;
    c = t[5] >> 25; t[6] += c; t[5] &= QQ_M25;
    c = t[2] >> 26; t[3] += c; t[2] &= QQ_M26;
    c = t[6] >> 26; t[7] += c; t[6] &= QQ_M26;
    c = t[3] >> 25; t[4] += c; t[3] &= QQ_M25;
    c = t[7] >> 25; t[8] += c; t[7] &= QQ_M25;
    c = t[4] >> 26; t[5] += c; t[4] &= QQ_M26;
    c = t[8] >> 26; t[9] += c; t[8] &= QQ_M26;
    c = t[9] >> 25; t[0] += c * 19u; t[9] &= QQ_M25;
    c = t[0] >> 26; t[1] += c; t[0] &= QQ_M26;
#pragma unroll
    for (int i = 0; i < 10; i++) h.v[i] = (u32)t[i];
}

// h = f * g.  Preconditions: 19*g.v[j] < 2^32 for all j (g <= 3.3T); see header for magnitude budget.
QQ_HD void fe_mul_inl(fe& h, const fe& f, const fe& g) {
    u32 g19[10], f2[10];
#pragma unroll
    for (int i = 1; i < 10; i++) g19[i] = g.v[i] * 19u;
#pragma unroll
    for (int i = 1; i < 10; i += 2) f2[i] = shl_alu<1>(f.v[i]);
    // Products are accumulated in chains of two (mul.wide + mad.wide) and the five partial sums of a column are then
    // added: ptxas keeps a 2-long chain as IMAD.WIDE with a 64-bit addend, whereas it rewrites longer chains into
    // IMAD.WIDE ..., RZ plus one 64-bit add per product (measured: 114 instead of 155 non-multiply instructions).
    u64 t[10];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        u64 part[5];
#pragma unroll
        for (int q = 0; q < 5; q++) {
            u64 acc = 0;
#pragma unroll
            for (int e = 0; e < 2; e++) {
                int i = 2 * q + e;
                int j = k - i;
                bool wrap = false;
                if (j < 0) { j += 10; wrap = true; }
                bool both_odd = (i & 1) && (j & 1);
                u32 a = both_odd ? f2[i] : f.v[i];
                u32 b = wrap ? g19[j] : g.v[j];
                acc = (e == 0) ? mul_wide(a, b) : mad_wide(a, b, acc);
            }
            part[q] = acc;
        }
        t[k] = ((part[0] + part[1]) + part[2]) + (part[3] + part[4]);
    }
    fe_reduce64(h, t);
}

// h = f^2.  Precondition: 19*f.v[j] < 2^32 (f <= 3.3T).
QQ_HD void fe_sq_inl(fe& h, const fe& f) {
    u32 f2[10], f4[10], f19[10];
#pragma unroll
    for (int i = 0; i < 10; i++) {
        f2[i] = shl_alu<1>(f.v[i]);
        f4[i] = shl_alu<2>(f.v[i]);
        f19[i] = f.v[i] * 19u;
    }
    u64 t[10];
#pragma unroll
    for (int k = 0; k < 10; k++) {
        // chains of two products (see fe_mul_inl); a column of the square has 5 or 6 distinct products
        u64 tot = 0, acc = 0;
        int len = 0, nparts = 0;
#pragma unroll
        for (int i = 0; i < 10; i++) {
            int j = k - i;
            bool wrap = false;
            if (j < 0) { j += 10; wrap = true; }
            if (j < i) continue;  // each unordered pair once
            bool both_odd = (i & 1) && (j & 1);
            int coef = (i == j ? 1 : 2) * (both_odd ? 2 : 1);  // 1, 2 or 4 on the i side; 19 on the j side
            u32 a = coef == 4 ? f4[i] : (coef == 2 ? f2[i] : f.v[i]);
            u32 b = wrap ? f19[j] : f.v[j];
            acc = (len == 0) ? mul_wide(a, b) : mad_wide(a, b, acc);
            len++;
            if (len == 2) {
                tot = (nparts == 0) ? acc : tot + acc;
                nparts++;
                len = 0;
            }
        }
        if (len) tot = (nparts == 0) ? acc : tot + acc;
        t[k] = tot;
    }
    fe_reduce64(h, t);
}

// Out-of-line copies for the device build: the big kernels call fe_mul / fe_sq thousands of times per thread, and
// inlining every call produced 150-370 KB of straight-line SASS per kernel (instruction-cache misses showed up as
// `no_instruction` stalls in ncu).  Arguments and result travel in registers (by-value ABI), so a call costs only
// CALL + RET + a few moves.  Define QQ_INLINE_FIELD_OPS to get the fully inlined code back.
#if defined(__CUDACC__) && !defined(QQ_INLINE_FIELD_OPS)
static __device__ __noinline__ fe fe_mul_ool(fe f, fe g) {
    fe h;
    fe_mul_inl(h, f, g);
    return h;
}
static __device__ __noinline__ fe fe_sq_ool(fe f) {
    fe h;
    fe_sq_inl(h, f);
    return h;
}
// n >= 1 squarings with the loop inside the callee: the square-root chains (254 squarings per decompress / compress)
// then pay the call marshalling once per run instead of once per squaring
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

// Canonical little-endian bytes as 8 x u32 words (fully reduced mod p).  Input: limbs < 2^31.
QQ_HD void fe_towords(u32 w[8], const fe& f) {
    fe t;
    fe_carry(t, f);
    fe_carry(t, t);  // now every limb within its width except possibly +small on limb 1; value < 2^255 + eps
    u32 h[10];
#pragma unroll
    for (int i = 0; i < 10; i++) h[i] = t.v[i];
    // q = floor((h + 19) / 2^255) in {0,1}: decides whether h >= p
    u32 q = (19u + h[0]) >> 26;
#pragma unroll
    for (int i = 1; i < 10; i++) q = (h[i] + q) >> ((i & 1) ? 25 : 26);
    h[0] += 19u * q;
    u32 c;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        c = h[i] >> ((i & 1) ? 25 : 26);
        h[i + 1] += c;
        h[i] &= (i & 1) ? QQ_M25 : QQ_M26;
    }
    h[9] &= QQ_M25;  // drop 2^255 * q
    w[0] = h[0] | (h[1] << 26);
    w[1] = (h[1] >> 6) | (h[2] << 19);
    w[2] = (h[2] >> 13) | (h[3] << 13);
    w[3] = (h[3] >> 19) | (h[4] << 6);
    w[4] = h[5] | (h[6] << 25);
    w[5] = (h[6] >> 7) | (h[7] << 19);
    w[6] = (h[7] >> 13) | (h[8] << 12);
    w[7] = (h[8] >> 20) | (h[9] << 6);
}

// Little-endian 8 x u32 words -> limbs; bit 255 is ignored (dalek FieldElement::from_bytes behaviour).
QQ_HD void fe_fromwords(fe& h, const u32 w[8]) {
    h.v[0] = w[0] & QQ_M26;
    h.v[1] = ((w[0] >> 26) | (w[1] << 6)) & QQ_M25;
    h.v[2] = ((w[1] >> 19) | (w[2] << 13)) & QQ_M26;
    h.v[3] = ((w[2] >> 13) | (w[3] << 19)) & QQ_M25;
    h.v[4] = (w[3] >> 6) & QQ_M26;
    h.v[5] = w[4] & QQ_M25;
    h.v[6] = ((w[4] >> 25) | (w[5] << 7)) & QQ_M26;
    h.v[7] = ((w[5] >> 19) | (w[6] << 13)) & QQ_M25;
    h.v[8] = ((w[6] >> 12) | (w[7] << 20)) & QQ_M26;
    h.v[9] = (w[7] >> 6) & QQ_M25;
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
    fe_sub(d, f, g);  // requires g tight
    return fe_iszero(d);
}
// h = b ? g : h     (b in {0,1}; branch-free)
QQ_HD void fe_cmov(fe& h, const fe& g, u32 b) {
    u32 m = 0u - b;
#pragma unroll
    for (int i = 0; i < 10; i++) h.v[i] ^= m & (h.v[i] ^ g.v[i]);
}
// h = b ? -h : h  (h tight)
QQ_HD void fe_cneg(fe& h, u32 b) {
    fe n;
    fe_neg(n, h);
    fe_cmov(h, n, b);
}
QQ_HD void fe_abs(fe& h) { fe_cneg(h, fe_isnegative(h)); }

// ---- constants (generated by tools/gen_consts.py; checked numerically by tests/test_host_arith.py) ----
#define QQ_FE_CONST(name, a0, a1, a2, a3, a4, a5, a6, a7, a8, a9) \
    QQ_HD fe name() {                                             \
        fe r = {{a0, a1, a2, a3, a4, a5, a6, a7, a8, a9}};        \
        return r;                                                 \
    }
#include "fe25519_consts.inc"

// sqrt_ratio_i (RFC 9496 4.2; dalek field.rs sqrt_ratio_i): returns was_square, r = sqrt(u/v) or sqrt(i*u/v), r >= 0.
// Branch-free (constant-time as written).  u, v tight or loose (<= 2T).
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
    fe uc;
    fe_carry(uc, u);
    fe_neg(uneg, uc);
    fe_mul(unegi, uneg, fe_sqrt_m1());
    u32 correct = fe_eq(check, uc);
    fe_carry(uneg, uneg);
    u32 flipped = fe_eq(check, uneg);
    u32 flipped_i = fe_eq(check, unegi);
    fe ri;
    fe_mul(ri, r, fe_sqrt_m1());
    fe_cmov(r, ri, flipped | flipped_i);
    fe_abs(r);
    return correct | flipped;
}

}  // namespace qq
