// Four-lane cooperative group operations for the latency-bound tail of the MSM (bucket reduction, window Horner).
//
// Those phases are long dependent chains of point additions / doublings executed by few threads; a single thread pays
// 8 (addition) or 7-8 (doubling) field multiplications of latency per operation.  Here four adjacent lanes own one
// point, lane r holding coordinate r of (X, Y, Z, T); every operation is two rounds of ONE field multiplication per
// lane with warp shuffles in between, so its latency is two multiplications.  Field products are inlined (not the
// out-of-line fe_mul) so that ptxas can overlap the independent chains of the running-sum recurrences.
#pragma once
#include "scalarmult.cuh"

namespace qq {

#define QQ_COOP_MASK 0xffffffffu

__device__ __forceinline__ fe fe_shfl4(const fe& a, int src) {   // value of lane `src` (0..3) of this lane's group
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(QQ_COOP_MASK, a.v[i], src, 4);
    return r;
}
__device__ __forceinline__ fe fe_shfl_xor1(const fe& a) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(QQ_COOP_MASK, a.v[i], 1, 4);
    return r;
}
__device__ __forceinline__ fe fe_sel(u32 c, const fe& a, const fe& b) {  // c ? a : b
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
}
// lane r loads / stores coordinate r of the point at p (one 128-byte line per group)
__device__ __forceinline__ fe coop_load(const u32x4* p, int r) {
    fe a;
    int o = 2 * r;
    fe_load(p, o, a);
    return a;
}
__device__ __forceinline__ void coop_store(u32x4* p, int r, const fe& a) {
    int o = 2 * r;
    fe_store(p, o, a);
}
__device__ __forceinline__ fe coop_identity(int r) {
    fe a;
    fe_0(a);
    a.v[0] = (r == 1 || r == 2) ? 1u : 0u;
    return a;
}
// distributed (X, Y, Z, T) -> distributed cached form (Y-X, Y+X, 2Z, 2dT): lanes 2 and 3 work on their own coordinate
template <bool INL>
__device__ __forceinline__ void coop_mul(fe& h, const fe& f, const fe& g) {
    if (INL) fe_mul_school(h, f, g);
    else fe_mul(h, f, g);
}
template <bool INL>
__device__ __forceinline__ void coop_sq(fe& h, const fe& f) {
    if (INL) fe_sq_inl(h, f);
    else fe_sq(h, f);
}
template <bool INL = true>
__device__ __forceinline__ fe coop_to_cached(const fe& mine, int r) {
    fe other = fe_shfl_xor1(mine);           // lanes 0/1 swap X and Y
    fe d, s, t, z2;
    fe_sub(d, r == 0 ? other : mine, r == 0 ? mine : other);   // lane 0: Y - X
    fe_add(s, mine, other);                                    // lane 1: Y + X
    fe_add(z2, mine, mine);                                    // lane 2: 2 Z
    coop_mul<INL>(t, mine, fe_2d());                           // lane 3: 2d T
    return r == 0 ? d : (r == 1 ? s : (r == 2 ? z2 : t));
}
// P (distributed) += Q (distributed cached, from coop_to_cached): two rounds of one multiplication per lane
template <bool INL = true>
__device__ __forceinline__ fe coop_add(const fe& mine, const fe& qc, int r) {
    fe other = fe_shfl_xor1(mine);
    fe a, d, s;
    fe_sub(d, r == 0 ? other : mine, r == 0 ? mine : other);   // lane 0: Y1 - X1
    fe_add(s, mine, other);                                    // lane 1: Y1 + X1
    a = r == 0 ? d : (r == 1 ? s : mine);                      // lane 2: Z1, lane 3: T1
    fe prod;
    coop_mul<INL>(prod, a, qc);                                // lane 0: A, lane 1: B, lane 2: D = 2 Z1 Z2, lane 3: C = 2d T1 T2
    fe A = fe_shfl4(prod, 0), B = fe_shfl4(prod, 1), D = fe_shfl4(prod, 2), C = fe_shfl4(prod, 3);
    fe E, F, G, H;
    fe_sub(E, B, A);
    fe_sub(F, D, C);
    fe_add(G, D, C);
    fe_add(H, B, A);
    // X3 = E F, Y3 = G H, Z3 = F G, T3 = E H
    fe u = (r == 0 || r == 3) ? E : (r == 1 ? G : F);
    fe v = (r == 0) ? F : ((r == 1 || r == 3) ? H : G);
    fe out;
    coop_mul<INL>(out, u, v);
    return out;
}
// P (distributed) = 2 P: one squaring + one multiplication per lane
template <bool INL = true>
__device__ __forceinline__ fe coop_dbl(const fe& mine, int r) {
    fe X = fe_shfl4(mine, 0), Y = fe_shfl4(mine, 1);
    fe xy;
    fe_add(xy, X, Y);
    fe in = r == 3 ? xy : mine;                                // lane 3 squares X + Y instead of T
    fe sq;
    coop_sq<INL>(sq, in);
    fe xx = fe_shfl4(sq, 0), yy = fe_shfl4(sq, 1), zz = fe_shfl4(sq, 2), s = fe_shfl4(sq, 3);
    fe cx, cy, cz, ct, t;
    fe_add(cy, yy, xx);
    fe_sub(cz, yy, xx);
    fe_sub(cx, s, cy);
    fe_add(t, zz, zz);
    fe_sub(ct, t, cz);
    // X3 = cx ct, Y3 = cy cz, Z3 = cz ct, T3 = cx cy
    fe u = (r == 0 || r == 3) ? cx : (r == 1 ? cy : cz);
    fe v = (r == 0 || r == 2) ? ct : (r == 1 ? cz : cy);
    fe out;
    coop_mul<INL>(out, u, v);
    return out;
}

}  // namespace qq
