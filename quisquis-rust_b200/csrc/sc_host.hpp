// Host-side arithmetic in Z/l, l = 2^252 + 27742317777372353535851937790883648493 (the Ristretto255 scalar field), for the
// verifier drivers that run next to the GPU batches: challenge powers, polynomial evaluation, the scalar checks of the
// Bayer-Groth sub-arguments (reference src/shuffle/*.rs use curve25519_dalek::scalar::Scalar for the same).
// Four 64-bit limbs, values always canonical (< l); multiplication = 4x4 schoolbook + Barrett reduction (HAC 14.42,
// b = 2^64, k = 4, mu = floor(2^512 / l)).  No secrets are handled here (verifier side): not constant time.
// Every function is also callable from device code (QQ_SC_FN): the range-proof fold kernel (rangeproof.cuh) runs the same
// arithmetic on the GPU; the constants are function-local there (namespace-scope host arrays are not visible to kernels).
#pragma once
#include <cstdint>
#include <cstring>
#include "fe25519.cuh"      // mp_mul8 / mp_mul4 and the carry-flag primitives (host: emulated) for the 32-bit limb form below
#ifdef __CUDACC__
#define QQ_SC_FN __host__ __device__ static inline
#define QQ_SC_FN_BIG __host__ __device__ __noinline__ static      // one copy of the large bodies in device code (see keccak_host.hpp)
#define QQ_SC_MEMBER __host__ __device__
#else
#define QQ_SC_FN static inline
#define QQ_SC_FN_BIG static inline
#define QQ_SC_MEMBER
#endif
#define QQ_SC_L_WORDS {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0ULL, 0x1000000000000000ULL}
#define QQ_SC_MU_WORDS {0xed9ce5a30a2c131bULL, 0x2106215d086329a7ULL, 0xffffffffffffffebULL, 0xffffffffffffffffULL, 0xfULL}

namespace qq_sc {

typedef unsigned __int128 u128;

struct sc {
    uint64_t v[4];
    QQ_SC_MEMBER bool operator==(const sc& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
    QQ_SC_MEMBER bool operator!=(const sc& o) const { return !(*this == o); }
};

static const uint64_t L[4] = QQ_SC_L_WORDS;
static const uint64_t MU[5] = QQ_SC_MU_WORDS;

QQ_SC_FN sc zero() { return sc{{0, 0, 0, 0}}; }
QQ_SC_FN sc one() { return sc{{1, 0, 0, 0}}; }
QQ_SC_FN sc from_u64(uint64_t x) { return sc{{x, 0, 0, 0}}; }
QQ_SC_FN bool is_zero(const sc& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }

// a >= l ?
QQ_SC_FN bool geq_l(const uint64_t a[4]) {
    const uint64_t L[4] = QQ_SC_L_WORDS;
    for (int i = 3; i >= 0; i--) {
        if (a[i] != L[i]) return a[i] > L[i];
    }
    return true;
}
// canonical 32 little-endian bytes -> sc; false when the value is >= l
QQ_SC_FN bool from_bytes(sc& out, const uint8_t b[32]) {
#if defined(__CUDA_ARCH__)
    if ((reinterpret_cast<uintptr_t>(b) & 15) == 0) {      // two 16-byte loads instead of 32 byte loads
        const uint4 lo = reinterpret_cast<const uint4*>(b)[0], hi = reinterpret_cast<const uint4*>(b)[1];
        out.v[0] = (uint64_t)lo.x | ((uint64_t)lo.y << 32);
        out.v[1] = (uint64_t)lo.z | ((uint64_t)lo.w << 32);
        out.v[2] = (uint64_t)hi.x | ((uint64_t)hi.y << 32);
        out.v[3] = (uint64_t)hi.z | ((uint64_t)hi.w << 32);
        return !geq_l(out.v);
    }
#endif
    memcpy(out.v, b, 32);
    return !geq_l(out.v);
}
QQ_SC_FN void to_bytes(uint8_t b[32], const sc& a) {
#if defined(__CUDA_ARCH__)
    // device: two 16-byte stores when the destination allows it (a memcpy to a byte pointer is 32 byte stores)
    if ((reinterpret_cast<uintptr_t>(b) & 15) == 0) {
        reinterpret_cast<uint4*>(b)[0] = make_uint4((uint32_t)a.v[0], (uint32_t)(a.v[0] >> 32), (uint32_t)a.v[1], (uint32_t)(a.v[1] >> 32));
        reinterpret_cast<uint4*>(b)[1] = make_uint4((uint32_t)a.v[2], (uint32_t)(a.v[2] >> 32), (uint32_t)a.v[3], (uint32_t)(a.v[3] >> 32));
        return;
    }
#endif
    memcpy(b, a.v, 32);
}
// n bytes (a multiple of 16), 16 at a time when both ends are 16-byte aligned (device); the plain memcpy otherwise
QQ_SC_FN void copy_aligned16(uint8_t* d, const uint8_t* s, size_t n) {
#if defined(__CUDA_ARCH__)
    if (((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(s) | n) & 15) == 0) {
        for (size_t i = 0; i < n / 16; i++) reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(s)[i];
        return;
    }
#endif
    memcpy(d, s, n);
}

QQ_SC_FN sc add(const sc& a, const sc& b) {
    const uint64_t L[4] = QQ_SC_L_WORDS;
    sc r;
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a.v[i] + b.v[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_l(r.v)) {   // a + b < 2 l < 2^254: no carry out of limb 3
        u128 br = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)r.v[i] - L[i] - (uint64_t)br;
            r.v[i] = (uint64_t)d;
            br = (d >> 64) & 1;
        }
    }
    return r;
}
QQ_SC_FN sc neg(const sc& a) {
    const uint64_t L[4] = QQ_SC_L_WORDS;
    if (is_zero(a)) return a;
    sc r;
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)L[i] - a.v[i] - (uint64_t)br;
        r.v[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    return r;
}
QQ_SC_FN sc sub(const sc& a, const sc& b) { return add(a, neg(b)); }

// x (8 limbs, < 2^512) mod l, through the special form l = 2^252 + c (c < 2^125, two limbs): 2^252 = -c (mod l), so
//   x = lo + 2^252 hi  =  lo - c hi;   c hi = ylo + 2^252 yhi  ->  - ylo + c yhi;   c yhi = zlo + 2^252 zhi  ->  zlo - c zhi
// with hi < 2^260, yhi < 2^133, zhi < 2^6: 18 64-bit products instead of the 39 of a Barrett reduction.
// r = lo + zlo + 2 l - ylo - c zhi lies in [0, 4 l + 2^253): a few conditional subtractions finish.
QQ_SC_FN void split252(const uint64_t* v, int limbs, uint64_t lo[4], uint64_t* hi, int hi_limbs) {
    // lo = v mod 2^252, hi = v >> 252 (v has `limbs` limbs; limbs beyond are zero)
    for (int i = 0; i < 3; i++) lo[i] = i < limbs ? v[i] : 0;
    lo[3] = limbs > 3 ? (v[3] & 0x0fffffffffffffffULL) : 0;
    for (int i = 0; i < hi_limbs; i++) {
        uint64_t a = 3 + i < limbs ? v[3 + i] : 0, b = 4 + i < limbs ? v[4 + i] : 0;
        hi[i] = (a >> 60) | (b << 4);
    }
}
// out (n + 2 limbs) = c * v (n limbs)
QQ_SC_FN void mul_c(const uint64_t* v, int n, uint64_t* out) {
    const uint64_t C[2] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL};
    for (int i = 0; i < n + 2; i++) out[i] = 0;
    for (int j = 0; j < 2; j++) {
        u128 carry = 0;
        for (int i = 0; i < n; i++) {
            carry += (u128)v[i] * C[j] + out[i + j];
            out[i + j] = (uint64_t)carry;
            carry >>= 64;
        }
        out[n + j] += (uint64_t)carry;      // no overflow: the running product fits n + 2 limbs
    }
}
QQ_SC_FN_BIG sc reduce512(const uint64_t x[8]) {
    const uint64_t L[4] = QQ_SC_L_WORDS;
    uint64_t lo[4], hi[5], y[7], ylo[4], yhi[3], z[5], zlo[4], zhi[1], w[3];
    split252(x, 8, lo, hi, 5);
    mul_c(hi, 5, y);                   // < 2^385
    split252(y, 7, ylo, yhi, 3);       // yhi < 2^133
    mul_c(yhi, 3, z);                  // < 2^258
    split252(z, 5, zlo, zhi, 1);       // zhi < 2^6
    mul_c(zhi, 1, w);                  // < 2^131
    // t = lo + zlo + 2 l - ylo - w  (5 limbs, never negative)
    uint64_t t[5];
    u128 acc = 0;
    for (int i = 0; i < 4; i++) {
        acc += (u128)lo[i] + zlo[i] + L[i] + L[i];
        t[i] = (uint64_t)acc;
        acc >>= 64;
    }
    t[4] = (uint64_t)acc;
    u128 br = 0;
    for (int i = 0; i < 5; i++) {
        u128 sub = (u128)(i < 4 ? ylo[i] : 0) + (i < 3 ? w[i] : 0) + (uint64_t)br;
        u128 d = (u128)t[i] - sub;
        t[i] = (uint64_t)d;
        br = (uint64_t)(0 - (uint64_t)(d >> 64));      // the subtrahend is below 2^66: the borrow (0, 1 or 2) is minus the high half
    }
    // t = q 2^252 + r with q < 2^5: t - q l = r - q c lies in (-2^130, 2^252), so one conditional addition of l finishes
    // (instead of up to six compare-and-subtract passes)
    const uint64_t C[2] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL};
    const uint64_t q = (t[4] << 4) | (t[3] >> 60);
    t[3] &= 0x0fffffffffffffffULL;
    const u128 qc0 = (u128)q * C[0], qc1 = (u128)q * C[1] + (uint64_t)(qc0 >> 64);
    const uint64_t sub3[3] = {(uint64_t)qc0, (uint64_t)qc1, (uint64_t)(qc1 >> 64)};
    u128 b2 = 0;
    uint64_t r[4];
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)t[i] - (i < 3 ? sub3[i] : 0) - (uint64_t)b2;
        r[i] = (uint64_t)d;
        b2 = (d >> 64) & 1;
    }
    const uint64_t m = 0 - (uint64_t)b2;      // negative: add l
    u128 cy = 0;
    for (int i = 0; i < 4; i++) {
        cy += (u128)r[i] + (L[i] & m);
        r[i] = (uint64_t)cy;
        cy >>= 64;
    }
    return sc{{r[0], r[1], r[2], r[3]}};
}
// ---- the same product with 32-bit limbs ----------------------------------------------------------------------------------
// Device code has no 64 x 64 -> 128 multiply: every (u128)a * b above becomes four 32-bit multiplies plus carry glue, ~650
// instructions per product in the one-thread-per-proof transcript kernels, whose running time IS their instruction count.
// Here the 8 x 8 product is fe25519's mp_mul8 (64 IMAD.WIDE with the carries in the flag) and the special-form reduction
// (l = 2^252 + c, c = 4 limbs) uses 4 x 4 blocks: c * hi (9 limbs), c * yhi (5 limbs), c * zhi (1 limb).  Same steps as
// reduce512, same canonical result; compiled for the host too (emulated carry flag) so tests/test_host_arith.py checks it.
QQ_SC_FN void w32_mul_c1(uint32_t out[5], uint32_t h) {      // out = c * h, h one limb
    const uint32_t C[4] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu};
    uint64_t carry = 0;
    for (int j = 0; j < 4; j++) {
        uint64_t t = (uint64_t)C[j] * h + carry;
        out[j] = (uint32_t)t;
        carry = t >> 32;
    }
    out[4] = (uint32_t)carry;
}
QQ_SC_FN sc reduce512_w32_inl(const uint32_t x[16]) {
    using namespace qq;
    const uint32_t C[4] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu};
    const uint32_t L2[8] = {0xb9eba7dau, 0xb024c634u, 0x45ef39acu, 0x29bdf3bdu, 0u, 0u, 0u, 0x20000000u};      // 2 l
    const uint32_t Lw[8] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0u, 0u, 0u, 0x10000000u};
    uint32_t hi[9], p0[8], p1[8], p2[5], y[14], yhi[5], q0[8], q1[5], z[10], w[5], t[8];
    for (int i = 0; i < 9; i++) hi[i] = (x[7 + i] >> 28) | (i < 8 ? x[8 + i] << 4 : 0u);
    // y = c * hi  (< 2^385: 13 limbs)
    mp_mul4(p0, C, hi);
    mp_mul4(p1, C, hi + 4);
    w32_mul_c1(p2, hi[8]);
    for (int i = 0; i < 4; i++) y[i] = p0[i];
    y[4] = add_cc(p0[4], p1[0]);
    y[5] = addc_cc(p0[5], p1[1]);
    y[6] = addc_cc(p0[6], p1[2]);
    y[7] = addc_cc(p0[7], p1[3]);
    y[8] = addc_cc(p1[4], 0u);
    y[9] = addc_cc(p1[5], 0u);
    y[10] = addc_cc(p1[6], 0u);
    y[11] = addc_cc(p1[7], 0u);
    y[12] = addc(0u, 0u);
    y[8] = add_cc(y[8], p2[0]);
    y[9] = addc_cc(y[9], p2[1]);
    y[10] = addc_cc(y[10], p2[2]);
    y[11] = addc_cc(y[11], p2[3]);
    y[12] = addc(y[12], p2[4]);
    y[13] = 0;
    for (int i = 0; i < 5; i++) yhi[i] = (y[7 + i] >> 28) | (y[8 + i] << 4);      // < 2^133
    // z = c * yhi  (< 2^258: 9 limbs)
    mp_mul4(q0, C, yhi);
    w32_mul_c1(q1, yhi[4]);
    for (int i = 0; i < 4; i++) z[i] = q0[i];
    z[4] = add_cc(q0[4], q1[0]);
    z[5] = addc_cc(q0[5], q1[1]);
    z[6] = addc_cc(q0[6], q1[2]);
    z[7] = addc_cc(q0[7], q1[3]);
    z[8] = addc(q1[4], 0u);
    const uint32_t zhi = (z[7] >> 28) | (z[8] << 4);      // < 2^6
    w32_mul_c1(w, zhi);                                    // < 2^131
    // t = lo + zlo + 2 l - ylo - w: in [0, 2^255)
    t[0] = add_cc(x[0], z[0]);
    for (int i = 1; i < 7; i++) t[i] = addc_cc(x[i], z[i]);
    t[7] = addc(x[7] & 0x0fffffffu, z[7] & 0x0fffffffu);
    t[0] = add_cc(t[0], L2[0]);
    for (int i = 1; i < 7; i++) t[i] = addc_cc(t[i], L2[i]);
    t[7] = addc(t[7], L2[7]);
    t[0] = sub_cc(t[0], y[0]);
    for (int i = 1; i < 7; i++) t[i] = subc_cc(t[i], y[i]);
    t[7] = subc(t[7], y[7] & 0x0fffffffu);
    t[0] = sub_cc(t[0], w[0]);
    for (int i = 1; i < 5; i++) t[i] = subc_cc(t[i], w[i]);
    for (int i = 5; i < 7; i++) t[i] = subc_cc(t[i], 0u);
    t[7] = subc(t[7], 0u);
    // t = q 2^252 + r, q < 8: r - q c in (-2^128, 2^252); one conditional addition of l finishes
    const uint32_t q = t[7] >> 28;
    t[7] &= 0x0fffffffu;
    uint32_t qc[5];
    w32_mul_c1(qc, q);
    t[0] = sub_cc(t[0], qc[0]);
    for (int i = 1; i < 5; i++) t[i] = subc_cc(t[i], qc[i]);
    for (int i = 5; i < 8; i++) t[i] = subc_cc(t[i], 0u);
    const uint32_t m = subc(0u, 0u);      // all ones when negative
    t[0] = add_cc(t[0], Lw[0] & m);
    for (int i = 1; i < 7; i++) t[i] = addc_cc(t[i], Lw[i] & m);
    t[7] = addc(t[7], Lw[7] & m);
    sc r;
    for (int i = 0; i < 4; i++) r.v[i] = (uint64_t)t[2 * i] | ((uint64_t)t[2 * i + 1] << 32);
    return r;
}
QQ_SC_FN_BIG sc reduce512_w32(const uint32_t x[16]) { return reduce512_w32_inl(x); }
// One out-of-line function per product on the device: the factors travel by value (registers), the 512-bit product stays in
// registers through the reduction (no generic loads / local stores between a caller, the product and the reduction).
QQ_SC_FN_BIG sc mul_w32(sc a, sc b) {
    uint32_t A[8], B[8], x[16];
    for (int i = 0; i < 4; i++) {
        A[2 * i] = (uint32_t)a.v[i]; A[2 * i + 1] = (uint32_t)(a.v[i] >> 32);
        B[2 * i] = (uint32_t)b.v[i]; B[2 * i + 1] = (uint32_t)(b.v[i] >> 32);
    }
    qq::mp_mul8(x, A, B);
    return reduce512_w32_inl(x);
}

// ---- lazy sums of products (the range-proof fold: sum over the proofs of a_p b_p) -----------------------------------------
// The 512-bit products are added up unreduced in 17 limbs (room for 2^38 products) and reduced once: the reduction is two
// thirds of a product's cost.
QQ_SC_FN void mul_wide_w32(uint32_t x[16], const sc& a, const sc& b) {
    uint32_t A[8], B[8];
    for (int i = 0; i < 4; i++) {
        A[2 * i] = (uint32_t)a.v[i]; A[2 * i + 1] = (uint32_t)(a.v[i] >> 32);
        B[2 * i] = (uint32_t)b.v[i]; B[2 * i + 1] = (uint32_t)(b.v[i] >> 32);
    }
    qq::mp_mul8(x, A, B);
}
QQ_SC_FN void acc17_add(uint32_t acc[17], const uint32_t x[16]) {
    using namespace qq;
    acc[0] = add_cc(acc[0], x[0]);
    for (int i = 1; i < 16; i++) acc[i] = addc_cc(acc[i], x[i]);
    acc[16] = addc(acc[16], 0u);
}
// acc (17 limbs) mod l: the low 512 bits through reduce512_w32, the top limb times 2^512 mod l
QQ_SC_FN sc acc17_reduce(const uint32_t acc[17]) {
    const sc R512 = sc{{0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL, 0xceec73d217f5be65ULL, 0x399411b7c309a3dULL}};      // 2^512 mod l
    sc lo = reduce512_w32(acc);
    if (acc[16] == 0) return lo;
    return add(lo, mul_w32(from_u64(acc[16]), R512));
}
// (a 2^k + 2^254 l - x) mod l for a canonical a, k < 64 and a 512-bit product x of canonical factors (x < 2^506 < 2^254 l):
// the difference a 2^k - x with ONE reduction
QQ_SC_FN_BIG sc shl_minus_wide(const sc& a, int k, const uint32_t x[16]) {
    using namespace qq;
    const uint32_t ML[16] = {0x0u, 0x0u, 0x0u, 0x0u, 0x0u, 0x0u, 0x0u, 0x40000000u, 0x973d74fbu, 0x960498c6u, 0xa8bde735u, 0x537be77u,
                             0x0u, 0x0u, 0x0u, 0x4000000u};      // 2^254 l
    uint32_t w[16];
    // w = a << k  (at most 317 bits)
    uint32_t A[8];
    for (int i = 0; i < 4; i++) { A[2 * i] = (uint32_t)a.v[i]; A[2 * i + 1] = (uint32_t)(a.v[i] >> 32); }
    const int ws = k >> 5, bs = k & 31;
    for (int i = 0; i < 16; i++) {
        const int s0 = i - ws, s1 = i - ws - 1;
        uint32_t lo = (s0 >= 0 && s0 < 8) ? A[s0] : 0u, below = (s1 >= 0 && s1 < 8) ? A[s1] : 0u;
        w[i] = bs ? ((lo << bs) | (below >> (32 - bs))) : lo;
    }
    w[0] = add_cc(w[0], ML[0]);
    for (int i = 1; i < 15; i++) w[i] = addc_cc(w[i], ML[i]);
    w[15] = addc(w[15], ML[15]);
    w[0] = sub_cc(w[0], x[0]);
    for (int i = 1; i < 15; i++) w[i] = subc_cc(w[i], x[i]);
    w[15] = subc(w[15], x[15]);
    return reduce512_w32(w);
}

QQ_SC_FN sc mul(const sc& a, const sc& b) {
#if defined(__CUDA_ARCH__)
    return mul_w32(a, b);
#else
    uint64_t x[8] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a.v[i] * b.v[j] + x[i + j];
            x[i + j] = (uint64_t)c;
            c >>= 64;
        }
        x[i + 4] = (uint64_t)c;
    }
    return reduce512(x);
#endif
}
// Scalar::from_bytes_mod_order_wide
QQ_SC_FN sc from_wide(const uint8_t b[64]) {
#if defined(__CUDA_ARCH__)
    uint32_t x[16];
    memcpy(x, b, 64);
    return reduce512_w32(x);
#else
    uint64_t x[8];
    memcpy(x, b, 64);
    return reduce512(x);
#endif
}
// a^(l - 2)
QQ_SC_FN_BIG sc invert(const sc& a) {
    const uint64_t L[4] = QQ_SC_L_WORDS;
    uint64_t e[4] = {L[0] - 2, L[1], L[2], L[3]};
    sc r = one();
    for (int bit = 252; bit >= 0; bit--) {
        r = mul(r, r);
        if ((e[bit >> 6] >> (bit & 63)) & 1) r = mul(r, a);
    }
    return r;
}

// a^-1 by the binary extended Euclidean algorithm (l is odd): about 3 x 253 shift / subtract steps on four limbs instead of
// the ~320 products of a^(l - 2) - 4 x fewer instructions in the one-thread-per-proof transcript kernels.  Variable time:
// for the verifiers' public challenges only.  invert_vartime(0) = 0.
QQ_SC_FN_BIG sc invert_vartime(const sc& a) {
    const uint64_t L[4] = QQ_SC_L_WORDS;
    if (is_zero(a)) return a;
    uint64_t u[4] = {a.v[0], a.v[1], a.v[2], a.v[3]}, v[4] = {L[0], L[1], L[2], L[3]};
    uint64_t x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
    // invariants: x1 a = u, x2 a = v (mod l); u, v odd-reduced towards gcd = 1
    for (int guard = 0; guard < 2048; guard++) {
        const bool u_one = (u[0] == 1) & ((u[1] | u[2] | u[3]) == 0), v_one = (v[0] == 1) & ((v[1] | v[2] | v[3]) == 0);
        if (u_one || v_one) break;
        uint64_t *w, *xw;
        const uint64_t *ws, *xs;      // subtrahends when both are odd
        bool halve;
        if (!(u[0] & 1)) { w = u; xw = x1; ws = nullptr; xs = nullptr; halve = true; }
        else if (!(v[0] & 1)) { w = v; xw = x2; ws = nullptr; xs = nullptr; halve = true; }
        else {
            bool ge = true;
            for (int i = 3; i >= 0; i--)
                if (u[i] != v[i]) { ge = u[i] > v[i]; break; }
            if (ge) { w = u; xw = x1; ws = v; xs = x2; } else { w = v; xw = x2; ws = u; xs = x1; }
            halve = false;
        }
        if (halve) {
            for (int i = 0; i < 3; i++) w[i] = (w[i] >> 1) | (w[i + 1] << 63);
            w[3] >>= 1;
            // x / 2 mod l: (x + l) / 2 when x is odd (x + l < 2^254)
            const uint64_t m = 0 - (xw[0] & 1);
            u128 c = 0;
            uint64_t t[4];
            for (int i = 0; i < 4; i++) {
                c += (u128)xw[i] + (L[i] & m);
                t[i] = (uint64_t)c;
                c >>= 64;
            }
            for (int i = 0; i < 3; i++) xw[i] = (t[i] >> 1) | (t[i + 1] << 63);
            xw[3] = t[3] >> 1;
        } else {
            u128 br = 0;
            for (int i = 0; i < 4; i++) {
                u128 d = (u128)w[i] - ws[i] - (uint64_t)br;
                w[i] = (uint64_t)d;
                br = (d >> 64) & 1;
            }
            // xw = xw - xs mod l
            br = 0;
            uint64_t t[4];
            for (int i = 0; i < 4; i++) {
                u128 d = (u128)xw[i] - xs[i] - (uint64_t)br;
                t[i] = (uint64_t)d;
                br = (d >> 64) & 1;
            }
            const uint64_t m = 0 - (uint64_t)br;      // borrowed: add l back
            u128 c = 0;
            for (int i = 0; i < 4; i++) {
                c += (u128)t[i] + (L[i] & m);
                xw[i] = (uint64_t)c;
                c >>= 64;
            }
        }
    }
    const bool u_one = (u[0] == 1) & ((u[1] | u[2] | u[3]) == 0);
    sc r;
    for (int i = 0; i < 4; i++) r.v[i] = u_one ? x1[i] : x2[i];
    return r;
}

// a^-1 without data-dependent branches or addressing: the binary extended Euclid with the subtraction and the halving merged,
// 508 fixed rounds (bit lengths of u and v sum to at most 506 and every round takes one off), every choice a mask.  For the GPU:
// invert_vartime's three-way branch diverges inside a warp and its pointer-selected operands live in local memory - it was 36 % of
// the range-proof transcript kernel (profiles/ncu_segments_*).  Same result; invert_fixed(0) = 0.
//   u even:          u /= 2,            x1 /= 2
//   u odd (u >= v after a swap):  u = (u - v) / 2,  x1 = (x1 - x2) / 2        (mod l; v stays odd)
QQ_SC_FN_BIG sc invert_fixed(const sc& a) {
    const uint64_t L[4] = QQ_SC_L_WORDS;
    const uint64_t L2[4] = {0xb024c634b9eba7daULL, 0x29bdf3bd45ef39acULL, 0ULL, 0x2000000000000000ULL};      // 2 l
    uint64_t u[4] = {a.v[0], a.v[1], a.v[2], a.v[3]}, v[4] = {L[0], L[1], L[2], L[3]};
    uint64_t x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
    for (int it = 0; it < 508; it++) {
        const uint64_t odd = 0 - (u[0] & 1);
        u128 br = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)u[i] - v[i] - (uint64_t)br;
            br = (d >> 64) & 1;
        }
        const uint64_t sw = odd & (0 - (uint64_t)br);      // u odd and u < v: swap the pairs
        for (int i = 0; i < 4; i++) {
            uint64_t t = (u[i] ^ v[i]) & sw;
            u[i] ^= t;
            v[i] ^= t;
            t = (x1[i] ^ x2[i]) & sw;
            x1[i] ^= t;
            x2[i] ^= t;
        }
        br = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)u[i] - (v[i] & odd) - (uint64_t)br;
            u[i] = (uint64_t)d;
            br = (d >> 64) & 1;
        }
        for (int i = 0; i < 3; i++) u[i] = (u[i] >> 1) | (u[i + 1] << 63);
        u[3] >>= 1;
        // t = x1 - (x2 & odd) in (-l, l); + k l with k in {0, 1, 2} making it non-negative and even (l is odd), then halved: < l
        uint64_t t[4];
        br = 0;
        for (int i = 0; i < 4; i++) {
            u128 d = (u128)x1[i] - (x2[i] & odd) - (uint64_t)br;
            t[i] = (uint64_t)d;
            br = (d >> 64) & 1;
        }
        const uint64_t neg = 0 - (uint64_t)br, par = 0 - (t[0] & 1);
        const uint64_t m1 = (neg & par) | (~neg & par);      // k == 1: parity odd (negative or not)
        const uint64_t m2 = neg & ~par;                     // k == 2: negative and even
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)t[i] + ((L[i] & m1) | (L2[i] & m2));
            t[i] = (uint64_t)c;
            c >>= 64;
        }
        for (int i = 0; i < 3; i++) x1[i] = (t[i] >> 1) | (t[i + 1] << 63);
        x1[3] = t[3] >> 1;
    }
    return sc{{x2[0], x2[1], x2[2], x2[3]}};
}

}  // namespace qq_sc
