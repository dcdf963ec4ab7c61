// Warp-cooperative group operations: ONE point per warp, one 32-bit limb per lane.
//
// The depth-bound ends of the pipelines (the 240-doubling Horner chain of the Pippenger MSM, msm.cuh; scalar multiplications
// of a 9-account anonymity set) are chains of dependent group operations executed by a handful of threads.  A single lane
// pays 7-8 field multiplications of latency per operation, the four-lane form (ge_coop.cuh) two, each a 72-product carry
// chain.  Here lane 8 q + k holds limb k of coordinate q (X, Y, Z, T) - exactly word 8 q + k of the stored point, so a
// point is one coalesced 128-byte access - and a field product is 8 lanes x 8 products:
//   * lane k gathers the operand limbs with shuffles and accumulates column k and column k + 8 of the schoolbook product
//     (8 products in all, the split between the two columns differs per lane), then folds V_k = col_k + 38 col_(k+8)
//     (2^256 = 38): a "wide limb" below 2^73 in three words, no carry has crossed a lane yet;
//   * sums and differences of products (the E, F, G, H of the addition, the four terms of the doubling) are taken on the
//     wide limbs, lane by lane, a multiple of 2p in wide-limb form keeping differences positive;
//   * warp_normalize brings wide limbs back to 32 bits: two shuffle passes move the excess to the next lane (out of lane 7,
//     cut at bit 255, times 19 into lane 0), the last carry bits are resolved with one ballot (generate / propagate
//     masks, the carry-lookahead addition trick of multi-precision GPU libraries).
// A doubling is two product rounds + two normalisations, an addition three normalisations; nothing here is a chain of 72
// dependent multiply-adds.  Values stay in the saturated form of fe25519.cuh (any value below 2^256), results are
// bit-compatible with ge25519.cuh as group elements (the encodings are canonical).
//
// Formulas: the same HWCD-2008 addition / doubling as ge25519.cuh (which replaces curve25519-dalek's curve_models); the
// doubling uses 2 X Y in place of (X + Y)^2 - X^2 - Y^2.
#pragma once
#include "scalarmult.cuh"

namespace qq {

struct w96 {
    u32 a, b, c;     // value a + 2^32 b + 2^64 c
};
#define QQ_WFULL 0xffffffffu

__device__ __forceinline__ w96 w96_shfl(const w96& v, int src) {
    w96 r;
    r.a = __shfl_sync(QQ_WFULL, v.a, src);
    r.b = __shfl_sync(QQ_WFULL, v.b, src);
    r.c = __shfl_sync(QQ_WFULL, v.c, src);
    return r;
}
__device__ __forceinline__ w96 w96_add(const w96& x, const w96& y) {
    w96 r;
    asm("add.cc.u32 %0, %3, %6; addc.cc.u32 %1, %4, %7; addc.u32 %2, %5, %8;"
        : "=r"(r.a), "=r"(r.b), "=r"(r.c) : "r"(x.a), "r"(x.b), "r"(x.c), "r"(y.a), "r"(y.b), "r"(y.c));
    return r;
}
__device__ __forceinline__ w96 w96_sub(const w96& x, const w96& y) {
    w96 r;
    asm("sub.cc.u32 %0, %3, %6; subc.cc.u32 %1, %4, %7; subc.u32 %2, %5, %8;"
        : "=r"(r.a), "=r"(r.b), "=r"(r.c) : "r"(x.a), "r"(x.b), "r"(x.c), "r"(y.a), "r"(y.b), "r"(y.c));
    return r;
}
__device__ __forceinline__ w96 w96_sel(bool c, const w96& x, const w96& y) {
    w96 r;
    r.a = c ? x.a : y.a;
    r.b = c ? x.b : y.b;
    r.c = c ? x.c : y.c;
    return r;
}
__device__ __forceinline__ w96 w96_zero() {
    w96 r;
    r.a = r.b = r.c = 0;
    return r;
}
__device__ __forceinline__ w96 w96_from_u32(u32 x) {
    w96 r;
    r.a = x;
    r.b = r.c = 0;
    return r;
}
// limb k of 2^44 * 2p = 2^44 (2^256 - 38), as a wide limb (>= 2^75.9: above any product column, see warp_mulw)
__device__ __forceinline__ w96 w96_bias(int k) {
    const u32 lo = k == 0 ? 0xffffffdau : 0xffffffffu;      // 2^32 - 38 | 2^32 - 1
    w96 r;
    r.a = 0;
    r.b = lo << 12;
    r.c = lo >> 20;
    return r;
}

// Wide limb k of A * B: A's limbs are register `a` of lanes 8 ga .. 8 ga + 7, B's limbs register `b` of lanes 8 gb .. 8 gb + 7.
// col_k = sum_{i <= k} A_i B_(k-i) < 2^67, col_(k+8) = sum_{i > k} A_i B_(k+8-i) < 7 * 2^64; V = col_k + 38 col_(k+8) < 2^72.1.
__device__ __forceinline__ w96 warp_mulw(u32 a, int ga, u32 b, int gb, int k) {
    u32 l0 = 0, l1 = 0, l2 = 0, h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const u32 ai = __shfl_sync(QQ_WFULL, a, 8 * ga + i);
        const u32 bj = __shfl_sync(QQ_WFULL, b, 8 * gb + ((k - i) & 7));
        const u32 alo = i <= k ? ai : 0u, ahi = i <= k ? 0u : ai;
        asm("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
            : "+r"(l0), "+r"(l1), "+r"(l2) : "r"(alo), "r"(bj));
        asm("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
            : "+r"(h0), "+r"(h1), "+r"(h2) : "r"(ahi), "r"(bj));
    }
    w96 v;
    asm("mad.lo.cc.u32 %0, %3, 38, %6; madc.hi.cc.u32 %1, %3, 38, %7; addc.u32 %2, %8, 0;\n\t"
        "mad.lo.cc.u32 %1, %4, 38, %1; madc.hi.u32 %2, %4, 38, %2;\n\t"
        "mad.lo.u32 %2, %5, 38, %2;"
        : "=&r"(v.a), "=&r"(v.b), "=&r"(v.c) : "r"(h0), "r"(h1), "r"(h2), "r"(l0), "r"(l1), "r"(l2));
    return v;
}

// wide limbs (each below 2^80) of one field element per 8-lane group -> 32-bit limbs of the same element mod p, value below
// 2^255 + 2^247.  k = limb index, base = first lane of the group.
__device__ __forceinline__ u32 warp_normalize(const w96& v, int k, int base) {
    const int src = base + ((k - 1) & 7);
    // pass 1: keep 32 bits (31 in lane 7: the cut is at 2^255 = 19), the rest moves to the next lane
    u32 lo = k == 7 ? (v.a & 0x7fffffffu) : v.a;
    u64 hi = k == 7 ? ((u64)(v.a >> 31) | ((u64)v.b << 1) | ((u64)v.c << 33)) : ((u64)v.b | ((u64)v.c << 32));
    u32 r0 = __shfl_sync(QQ_WFULL, (u32)hi, src), r1 = __shfl_sync(QQ_WFULL, (u32)(hi >> 32), src);
    u64 recv = (u64)r0 | ((u64)r1 << 32);           // < 2^50
    if (k == 0) recv *= 19;                          // < 2^54.3
    u64 t = recv + lo;
    // pass 2
    lo = k == 7 ? ((u32)t & 0x7fffffffu) : (u32)t;
    u32 hi2 = k == 7 ? (u32)(t >> 31) : (u32)(t >> 32);      // < 2^23.3
    u32 rr = __shfl_sync(QQ_WFULL, hi2, src);
    if (k == 0) rr *= 19;                            // < 2^27.6
    u64 y = (u64)lo + rr;
    u32 yl = (u32)y, g = (u32)(y >> 32);             // g in {0, 1}; lane 7 has lo < 2^31: no carry out of the element
    // pass 3: carry g_k enters limb k + 1; limbs that are all ones propagate
    u32 G = __ballot_sync(QQ_WFULL, g != 0), P = __ballot_sync(QQ_WFULL, yl == 0xffffffffu);
    G = (G >> base) & 0xffu;
    P = (P >> base) & 0xffu;
    u32 s = (G << 1) + P;
    return yl + (((s ^ P) >> k) & 1u);
}

// field product of the elements held by groups ga and gb (registers a, b), result limb k for this lane's group
__device__ __forceinline__ u32 warp_fe_mul(u32 a, int ga, u32 b, int gb, int k, int base) {
    return warp_normalize(warp_mulw(a, ga, b, gb, k), k, base);
}

// limb k of constants needed lane-wise
__device__ __forceinline__ u32 fe_limb(const fe& f, int k) {
    u32 r = f.v[0];
#pragma unroll
    for (int i = 1; i < 8; i++) r = k == i ? f.v[i] : r;
    return r;
}

// P = 2 P.  v: this lane's limb of P (lane 8 q + k: coordinate q, limb k).  T of the input is not read.
__device__ __forceinline__ u32 warp_dbl(u32 v, int q, int k) {
    // round 1: X X | Y Y | Z Z | X Y
    w96 V = warp_mulw(v, q == 3 ? 0 : q, v, q == 3 ? 1 : q, k);
    w96 t1 = w96_shfl(V, (q == 0 ? 24 : 0) + k);    // group 0: XY, others: XX
    w96 t2 = w96_shfl(V, 8 + k);                      // YY
    w96 t3 = w96_shfl(V, 16 + k);                     // ZZ
    // group 0: cx = 2 XY | 1: cy = YY + XX | 2: cz = YY - XX | 3: ct = 2 ZZ - YY + XX
    w96 A1 = q == 0 ? t1 : (q == 3 ? t3 : t2);
    w96 A2 = q == 2 ? w96_bias(k) : (q == 3 ? t3 : t1);
    w96 A3 = q == 3 ? w96_add(t1, w96_bias(k)) : w96_zero();
    w96 N = q == 2 ? t1 : (q == 3 ? t2 : w96_zero());
    w96 r = w96_sub(w96_add(w96_add(A1, A2), A3), N);
    u32 c = warp_normalize(r, k, 8 * q);
    // round 2: X3 = cx ct | Y3 = cy cz | Z3 = cz ct | T3 = cx cy
    const int ga = q == 3 ? 0 : q, gb = q == 0 ? 3 : (q == 1 ? 2 : (q == 2 ? 3 : 1));
    return warp_fe_mul(c, ga, c, gb, k, 8 * q);
}

// distributed (X, Y, Z, T) -> distributed cached form: Y - X | Y + X | 2 Z | 2d T
__device__ __forceinline__ u32 warp_to_cached(u32 v, int q, int k) {
    const u32 x = __shfl_sync(QQ_WFULL, v, k), y = __shfl_sync(QQ_WFULL, v, 8 + k);
    const u32 d2 = fe_limb(fe_2d(), k);
    w96 prod = warp_mulw(v, 3, d2, 3, k);           // meaningful in group 3 (every lane of a group holds its limb of 2d)
    w96 r;
    if (q == 0) r = w96_sub(w96_add(w96_from_u32(y), w96_bias(k)), w96_from_u32(x));
    else if (q == 1) r = w96_add(w96_from_u32(y), w96_from_u32(x));
    else if (q == 2) r = w96_add(w96_from_u32(v), w96_from_u32(v));
    else r = prod;
    return warp_normalize(r, k, 8 * q);
}

// P += Q, Q in the distributed cached form of warp_to_cached
__device__ __forceinline__ u32 warp_add(u32 v, u32 qc, int q, int k) {
    const u32 x = __shfl_sync(QQ_WFULL, v, k), y = __shfl_sync(QQ_WFULL, v, 8 + k);
    w96 r;
    if (q == 0) r = w96_sub(w96_add(w96_from_u32(y), w96_bias(k)), w96_from_u32(x));    // Y1 - X1
    else if (q == 1) r = w96_add(w96_from_u32(y), w96_from_u32(x));                     // Y1 + X1
    else r = w96_from_u32(v);                                                           // Z1 | T1
    const u32 a = warp_normalize(r, k, 8 * q);
    // round 1: A = (Y1 - X1)(Y2 - X2) | B = (Y1 + X1)(Y2 + X2) | D = Z1 2 Z2 | C = T1 2d T2
    w96 V = warp_mulw(a, q, qc, q, k);
    // group 0: E = B - A | 1: F = D - C | 2: G = D + C | 3: H = B + A
    const bool ba = q == 0 || q == 3;
    w96 t1 = w96_shfl(V, (ba ? 8 : 16) + k);         // B | D
    w96 t2 = w96_shfl(V, (ba ? 0 : 24) + k);         // A | C
    r = q < 2 ? w96_sub(w96_add(t1, w96_bias(k)), t2) : w96_add(t1, t2);
    const u32 c = warp_normalize(r, k, 8 * q);
    // round 2: X3 = E F | Y3 = G H | Z3 = F G | T3 = E H
    const int ga = q == 1 ? 2 : (q == 2 ? 1 : 0), gb = q == 0 ? 1 : (q == 2 ? 2 : 3);
    return warp_fe_mul(c, ga, c, gb, k, 8 * q);
}

__device__ __forceinline__ u32 warp_identity(int q, int k) { return (k == 0 && (q == 1 || q == 2)) ? 1u : 0u; }
// lane L reads / writes word L of the stored point (X | Y | Z | T, 8 words each)
__device__ __forceinline__ u32 warp_point_load(const u32x4* p) { return reinterpret_cast<const u32*>(p)[threadIdx.x & 31]; }
__device__ __forceinline__ void warp_point_store(u32x4* p, u32 v) { reinterpret_cast<u32*>(p)[threadIdx.x & 31] = v; }

}  // namespace qq
