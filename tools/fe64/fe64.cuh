// EXPERIMENT (not part of libqq_b200.so): GF(2^255-19) and the Edwards group law on the FP64 pipe (sm_100a), for warps that
// run BESIDE the integer warps.  Measured in tools/vb64_bench.cu, result in profiles/vb64_bench_r02.jsonl and DESIGN.md
// section 4: exact and byte-identical, but the two pipes do not add up on B200, so the library keeps the integer kernels.
//
// Why: every long kernel here is bound by the integer-multiply pipe (IMAD.WIDE) while the FP64 pipe of
// the SM sits idle; B200 issues DFMA at up to 64 lanes / clk / SM, about the same multiplier area per clock as IMAD.WIDE.
// A field multiplication written ONLY with FP64 instructions (no ALU carry glue, which would compete with the integer
// warps for issue slots) lets a second group of warps do scalar multiplications on that pipe at the same time.
//
// Representation (Bernstein's floating-point Curve25519 form, re-derived for 53-bit mantissas): 12 doubles, limb i is an
// integer multiple of 2^o_i, o_i = ceil(21.25 i) = {0,22,43,64,85,107,128,149,170,192,213,234}, o_12 = 255; the value is
// the plain sum of the limbs, limbs are SIGNED.  A carried limb has |x_i| <= 2^(o_(i+1) - 1).  Every product a_i b_j is a
// multiple of 2^o_(i+j) (ceil is superadditive), products that wrap past 2^255 are scaled by 19 * 2^-255 exactly, and a
// column of 12 products of carried limbs stays below 2^(o_k + 48.96): 4.04 bits under the 53-bit mantissa, so operands may
// be sums of carried values as long as (terms of a) x (terms of b) <= 16 -- the group law below needs at most 12
// (tests/test_host_arith.py checks the bound with interval arithmetic and the arithmetic against big integers).  All
// FP64 operations are therefore EXACT (no rounding ever happens except the intended one in the carry), and the result
// is converted back to the canonical saturated form, so the encodings are byte-identical to the integer path's.
//
// One multiplication = 11 DMUL (19 * 2^-255 * b_j) + 144 DFMA + 43 DADD/DFMA of carry = 198 FP64 instructions, no ALU.
// One squaring = 22 + 78 + 43 = 143.
#pragma once
#include <string.h>
#include "../../quisquis-rust_b200/csrc/ge25519.cuh"

namespace qq {

#define QQ_F64_LIMBS 12
struct fe64 {
    double v[QQ_F64_LIMBS];
};

QQ_HD constexpr int f64_off(int i) { return (85 * i + 3) / 4; }   // ceil(21.25 i)

QQ_HD double f64_from_bits(u64 b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
QQ_HD double f64_p2(int e) { return f64_from_bits((u64)(1023 + e) << 52); }                      // 2^e
QQ_HD double f64_round_const(int e) { return f64_from_bits(((u64)(1023 + 52 + e) << 52) | (1ull << 51)); }   // 1.5 * 2^(52+e)
QQ_HD double f64_c19() { return f64_from_bits(((u64)(1023 - 251) << 52) | (3ull << 48)); }       // 19 * 2^-255
#if defined(__CUDA_ARCH__)
#define QQ_FMA(a, b, c) __fma_rn((a), (b), (c))
#define QQ_DADD(a, b) __dadd_rn((a), (b))
#define QQ_DSUB(a, b) __dsub_rn((a), (b))
#else
#define QQ_FMA(a, b, c) ((a) * (b) + (c))
#define QQ_DADD(a, b) ((a) + (b))
#define QQ_DSUB(a, b) ((a) - (b))
#endif

// Carry: limb i keeps its part below 2^o_(i+1) (round to nearest: signed remainder), the rest moves up; the carry out of
// limb 11 is a multiple of 2^255 and re-enters limb 0 times 19 * 2^-255.  A second short lap (limbs 0, 1) absorbs it.
// (x + M) - M with M = 1.5 * 2^(52 + e) rounds x to a multiple of 2^e for |x| < 2^(51 + e).
QQ_HD void fe64_carry(double r[QQ_F64_LIMBS]) {
#pragma unroll
    for (int i = 0; i < QQ_F64_LIMBS - 1; i++) {
        const double M = f64_round_const(f64_off(i + 1));
        double c = QQ_DSUB(QQ_DADD(r[i], M), M);
        r[i] = QQ_DSUB(r[i], c);
        r[i + 1] = QQ_DADD(r[i + 1], c);
    }
    {
        const double M = f64_round_const(255);
        double c = QQ_DSUB(QQ_DADD(r[11], M), M);
        r[11] = QQ_DSUB(r[11], c);
        r[0] = QQ_FMA(c, f64_c19(), r[0]);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const double M = f64_round_const(f64_off(i + 1));
        double c = QQ_DSUB(QQ_DADD(r[i], M), M);
        r[i] = QQ_DSUB(r[i], c);
        r[i + 1] = QQ_DADD(r[i + 1], c);
    }
}

QQ_HD void fe64_mul_inl(fe64& h, const fe64& a, const fe64& b) {
    double bp[QQ_F64_LIMBS], r[QQ_F64_LIMBS];
    const double c19 = f64_c19();
    bp[0] = 0.0;
#pragma unroll
    for (int j = 1; j < QQ_F64_LIMBS; j++) bp[j] = b.v[j] * c19;
#pragma unroll
    for (int k = 0; k < QQ_F64_LIMBS; k++) {
        double acc = a.v[0] * b.v[k];
#pragma unroll
        for (int i = 1; i < QQ_F64_LIMBS; i++) acc = QQ_FMA(a.v[i], (i <= k ? b.v[k - i] : bp[k + 12 - i]), acc);
        r[k] = acc;
    }
    fe64_carry(r);
#pragma unroll
    for (int k = 0; k < QQ_F64_LIMBS; k++) h.v[k] = r[k];
}

// h = f^2 * (TWICE ? 2 : 1)
template <bool TWICE = false>
QQ_HD void fe64_sq_inl(fe64& h, const fe64& a) {
    double a2[QQ_F64_LIMBS], ap[QQ_F64_LIMBS], r[QQ_F64_LIMBS];
    const double c19 = f64_c19();
#pragma unroll
    for (int j = 0; j < QQ_F64_LIMBS; j++) {
        a2[j] = a.v[j] + a.v[j];
        ap[j] = a.v[j] * c19;
    }
#pragma unroll
    for (int k = 0; k < QQ_F64_LIMBS; k++) {
        double acc = 0.0;
        bool first = true;
#pragma unroll
        for (int i = 0; i < QQ_F64_LIMBS; i++) {
            // unfolded pair (i, k - i) with i <= k - i
            int j = k - i;
            if (j >= i && j < QQ_F64_LIMBS) {
                double x = a.v[i], y = (j == i) ? a.v[j] : a2[j];
                acc = first ? x * y : QQ_FMA(x, y, acc);
                first = false;
            }
            // folded pair (i, k + 12 - i) with i <= k + 12 - i
            j = k + 12 - i;
            if (j >= i && j < QQ_F64_LIMBS) {
                double x = (j == i) ? a.v[i] : a2[i], y = ap[j];
                acc = first ? x * y : QQ_FMA(x, y, acc);
                first = false;
            }
        }
        r[k] = TWICE ? acc + acc : acc;
    }
    fe64_carry(r);
#pragma unroll
    for (int k = 0; k < QQ_F64_LIMBS; k++) h.v[k] = r[k];
}

// Device build: the products are out-of-line functions (the loop body of a scalar multiplication must stay within reach of
// the instruction caches, see kernels.cuh); operands and result travel by value.
#if defined(__CUDACC__) && !defined(QQ_F64_INLINE)
static __device__ __noinline__ fe64 fe64_mul_ool(fe64 a, fe64 b) {
    fe64 h;
    fe64_mul_inl(h, a, b);
    return h;
}
static __device__ __noinline__ fe64 fe64_sq_ool(fe64 a) {
    fe64 h;
    fe64_sq_inl<false>(h, a);
    return h;
}
static __device__ __noinline__ fe64 fe64_sq2_ool(fe64 a) {
    fe64 h;
    fe64_sq_inl<true>(h, a);
    return h;
}
#endif
QQ_HD void fe64_mul(fe64& h, const fe64& a, const fe64& b) {
#if defined(__CUDA_ARCH__) && !defined(QQ_F64_INLINE)
    h = fe64_mul_ool(a, b);
#else
    fe64_mul_inl(h, a, b);
#endif
}
template <bool TWICE = false>
QQ_HD void fe64_sq(fe64& h, const fe64& a) {
#if defined(__CUDA_ARCH__) && !defined(QQ_F64_INLINE)
    h = TWICE ? fe64_sq2_ool(a) : fe64_sq_ool(a);
#else
    fe64_sq_inl<TWICE>(h, a);
#endif
}

QQ_HD void fe64_add(fe64& h, const fe64& f, const fe64& g) {
#pragma unroll
    for (int i = 0; i < QQ_F64_LIMBS; i++) h.v[i] = f.v[i] + g.v[i];
}
QQ_HD void fe64_sub(fe64& h, const fe64& f, const fe64& g) {
#pragma unroll
    for (int i = 0; i < QQ_F64_LIMBS; i++) h.v[i] = f.v[i] - g.v[i];
}
QQ_HD void fe64_0(fe64& h) {
#pragma unroll
    for (int i = 0; i < QQ_F64_LIMBS; i++) h.v[i] = 0.0;
}
QQ_HD void fe64_1(fe64& h) {
    fe64_0(h);
    h.v[0] = 1.0;
}

// saturated integer form (any value in [0, 2^256)) -> carried fe64
QQ_HD void fe64_from_fe(fe64& h, const fe& f) {
    double r[QQ_F64_LIMBS];
#pragma unroll
    for (int i = 0; i < QQ_F64_LIMBS; i++) {
        const int o = f64_off(i), w = o >> 5, sh = o & 31;
        const int bits = (i == QQ_F64_LIMBS - 1) ? 22 : f64_off(i + 1) - o;    // the top limb also takes bit 255
        u64 two = (u64)f.v[w] | (w + 1 < 8 ? (u64)f.v[w + 1] << 32 : 0ull);
        u32 x = (u32)(two >> sh) & ((1u << bits) - 1u);
        r[i] = (double)(int)x * f64_p2(o);
    }
    fe64_carry(r);
#pragma unroll
    for (int k = 0; k < QQ_F64_LIMBS; k++) h.v[k] = r[k];
}
// fe64 with limbs of up to 16 carried terms -> saturated integer form (value in [0, 2^256), not necessarily reduced)
QQ_HD void fe64_to_fe(fe& h, const fe64& f) {
    long long q[QQ_F64_LIMBS];
    // integer limb + 16 p in limb form (p = sum (2^s_i - 1) 2^o_i - 18): every limb becomes positive
#pragma unroll
    for (int i = 0; i < QQ_F64_LIMBS; i++) {
        const int s = f64_off(i + 1) - f64_off(i);
        long long bias = 16ll * (((long long)1 << s) - 1) - (i == 0 ? 16ll * 18 : 0ll);
        q[i] = (long long)(f.v[i] * f64_p2(-f64_off(i))) + bias;
    }
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        long long carry = 0;
#pragma unroll
        for (int i = 0; i < QQ_F64_LIMBS; i++) {
            const int s = f64_off(i + 1) - f64_off(i);
            long long t = q[i] + carry;
            q[i] = t & (((long long)1 << s) - 1);
            carry = t >> s;
        }
        if (pass == 0) q[0] += 19 * carry;       // carry < 2^6; second pass: carry in {0, 1}, kept as bit 255
        else q[11] += carry << 21;
    }
#pragma unroll
    for (int w = 0; w < 8; w++) h.v[w] = 0;
#pragma unroll
    for (int i = 0; i < QQ_F64_LIMBS; i++) {
        const int o = f64_off(i), w = o >> 5, sh = o & 31;
        u64 two = (u64)q[i] << sh;
        h.v[w] |= (u32)two;
        if (w + 1 < 8) h.v[w + 1] |= (u32)(two >> 32);
    }
}

// ---- group law (same formulas as ge25519.cuh; the comments give the number of carried terms of each operand) -------------
struct ge64_p3 {
    fe64 X, Y, Z, T;
};
struct ge64_cached {
    fe64 YpX, YmX, Z2, T2d;
};
QQ_HD void ge64_identity(ge64_p3& p) {
    fe64_0(p.X);
    fe64_1(p.Y);
    fe64_1(p.Z);
    fe64_0(p.T);
}
QQ_HD void ge64_from_p3(ge64_p3& r, const ge_p3& p) {
    fe64_from_fe(r.X, p.X);
    fe64_from_fe(r.Y, p.Y);
    fe64_from_fe(r.Z, p.Z);
    fe64_from_fe(r.T, p.T);
}
QQ_HD void ge64_to_p3(ge_p3& r, const ge64_p3& p) {
    fe64_to_fe(r.X, p.X);
    fe64_to_fe(r.Y, p.Y);
    fe64_to_fe(r.Z, p.Z);
    fe64_to_fe(r.T, p.T);
}
// (Y + X, Y - X, 2 Z: two terms each; 2d T carried)
QQ_HD void ge64_to_cached(ge64_cached& c, const ge64_p3& p, const fe64& d2) {
    fe64_add(c.YpX, p.Y, p.X);
    fe64_sub(c.YmX, p.Y, p.X);
    fe64_add(c.Z2, p.Z, p.Z);
    fe64_mul(c.T2d, p.T, d2);
}
// c = b ? -c : c
QQ_HD void ge64_cached_cneg(ge64_cached& c, u32 b) {
#pragma unroll
    for (int i = 0; i < QQ_F64_LIMBS; i++) {
        double p = c.YpX.v[i], m = c.YmX.v[i], t = c.T2d.v[i];
        c.YpX.v[i] = b ? m : p;
        c.YmX.v[i] = b ? p : m;
        c.T2d.v[i] = b ? -t : t;
    }
}
// r = p + q: p carried, q as ge64_to_cached leaves it.  Operand terms: 2x2, 2x2, 1x1, 1x2, then 2x2 four times.
QQ_HD void ge64_add(ge64_p3& r, const ge64_p3& p, const ge64_cached& q) {
    fe64 a, b, c, d, e, f, g, h, t;
    fe64_sub(t, p.Y, p.X);
    fe64_mul(a, t, q.YmX);
    fe64_add(t, p.Y, p.X);
    fe64_mul(b, t, q.YpX);
    fe64_mul(c, p.T, q.T2d);
    fe64_mul(d, p.Z, q.Z2);
    fe64_sub(e, b, a);
    fe64_sub(f, d, c);
    fe64_add(g, d, c);
    fe64_add(h, b, a);
    fe64_mul(r.X, f, e);
    fe64_mul(r.Y, g, h);
    fe64_mul(r.Z, f, g);
    fe64_mul(r.T, e, h);
}
// r = 2p, p carried.  Operand terms: squarings 1, 1, 1, 2x2; cy = 2, cz = 2, cx = 3, ct = 3 (2 ZZ comes carried out of the
// doubled squaring); products 3x3, 2x2, 2x3, 3x2.
template <bool WITH_T>
QQ_HD void ge64_dbl(ge64_p3& r, const ge64_p3& p) {
    fe64 xx, yy, zz2, s, cx, cy, cz, ct, t;
    fe64_sq(xx, p.X);
    fe64_sq(yy, p.Y);
    fe64_sq<true>(zz2, p.Z);
    fe64_add(t, p.X, p.Y);
    fe64_sq(s, t);
    fe64_add(cy, yy, xx);
    fe64_sub(cz, yy, xx);
    fe64_sub(cx, s, cy);
    fe64_sub(ct, zz2, cz);
    fe64_mul(r.X, cx, ct);
    fe64_mul(r.Y, cy, cz);
    fe64_mul(r.Z, cz, ct);
    if (WITH_T) fe64_mul(r.T, cx, cy);
}

}  // namespace qq

// ---- split variable base on the FP64 pipe (same digits, same table scheme as vbs_* in scalarmult.cuh) ---------------------
#include "../../quisquis-rust_b200/csrc/scalarmult.cuh"
namespace qq {

#define QQ_VBS64_ENTRY_D 48                                                  // cached point: 4 x 12 doubles = 384 B
#define QQ_VBS64_TABLE_D (QQ_VBS_PARTS * QQ_VB_ENTRIES * QQ_VBS64_ENTRY_D)   // 13 824 B per thread

QQ_HD void fe64_store(double* dst, const fe64& a) {
#if defined(__CUDA_ARCH__)
    double2* d = reinterpret_cast<double2*>(dst);
#pragma unroll
    for (int i = 0; i < 6; i++) d[i] = make_double2(a.v[2 * i], a.v[2 * i + 1]);
#else
    for (int i = 0; i < QQ_F64_LIMBS; i++) dst[i] = a.v[i];
#endif
}
QQ_HD void fe64_load(fe64& a, const double* src) {
#if defined(__CUDA_ARCH__)
    const double2* s = reinterpret_cast<const double2*>(src);
#pragma unroll
    for (int i = 0; i < 6; i++) {
        double2 q = s[i];
        a.v[2 * i] = q.x;
        a.v[2 * i + 1] = q.y;
    }
#else
    for (int i = 0; i < QQ_F64_LIMBS; i++) a.v[i] = src[i];
#endif
}
QQ_HD void ge64_cached_store(double* dst, const ge64_cached& c) {
    fe64_store(dst, c.YpX);
    fe64_store(dst + 12, c.YmX);
    fe64_store(dst + 24, c.Z2);
    fe64_store(dst + 36, c.T2d);
}
QQ_HD void ge64_cached_load(ge64_cached& c, const double* src) {
    fe64_load(c.YpX, src);
    fe64_load(c.YmX, src + 12);
    fe64_load(c.Z2, src + 24);
    fe64_load(c.T2d, src + 36);
}
QQ_HD void vbs64_build_tables(double* tbl, const ge64_p3& p, const fe64& d2) {
    ge64_p3 q = p;
#pragma unroll 1
    for (int part = 0; part < QQ_VBS_PARTS; part++) {
        double* t = tbl + part * (QQ_VB_ENTRIES * QQ_VBS64_ENTRY_D);
        ge64_cached c0, c;
        ge64_p3 id, r;
        ge64_identity(id);
        ge64_to_cached(c, id, d2);
        ge64_cached_store(t, c);
        ge64_to_cached(c0, q, d2);
        ge64_cached_store(t + QQ_VBS64_ENTRY_D, c0);
        r = q;
#pragma unroll 1
        for (int i = 2; i <= 8; i++) {
            ge64_add(r, r, c0);
            ge64_to_cached(c, r, d2);
            ge64_cached_store(t + QQ_VBS64_ENTRY_D * i, c);
        }
        if (part + 1 < QQ_VBS_PARTS) {
#pragma unroll 1
            for (int i = 0; i < 63; i++) ge64_dbl<false>(q, q);
            ge64_dbl<true>(q, q);
        }
    }
}
QQ_HD void vbs64_scalarmult(ge64_p3& r, const double* tbl, const u32 s[8]) {
    u32 rr[9];
    sc_recode_bias<4, 64>(rr, s);
    ge64_identity(r);
#pragma unroll 1
    for (int half = 1; half >= 0; half--) {
        u32 w0 = half ? rr[1] : rr[0], w1 = half ? rr[3] : rr[2], w2 = half ? rr[5] : rr[4], w3 = half ? rr[7] : rr[6];
#pragma unroll 1
        for (int j = 7; j >= 0; j--) {
            if (!(half == 1 && j == 7)) {
#pragma unroll 1
                for (int d = 0; d < 3; d++) ge64_dbl<false>(r, r);
                ge64_dbl<true>(r, r);
            }
#pragma unroll 1
            for (int part = 0; part < QQ_VBS_PARTS; part++) {
                int d = (int)(w0 >> 28) - 8;
                w0 = (w0 << 4);
                u32 tw = w0; w0 = w1; w1 = w2; w2 = w3; w3 = tw;
                u32 neg = (u32)d >> 31;
                u32 idx = (u32)((d ^ (d >> 31)) - (d >> 31));
                ge64_cached c;
                ge64_cached_load(c, tbl + QQ_VBS64_ENTRY_D * (part * QQ_VB_ENTRIES + idx));
                ge64_cached_cneg(c, neg);
                ge64_add(r, r, c);
            }
        }
    }
}

}  // namespace qq
