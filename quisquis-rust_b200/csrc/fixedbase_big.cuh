// Fixed-base scalar multiplication over LARGE window tables resident in L2 / HBM.
//
// The shared-memory kernel (k_fixedbase<6>, kernels.cuh) walks 43 windows.  A B200 has 126 MB of L2 and 180 GB of HBM3e,
// so the generators B and H (the only fixed bases of the reference: src/ristretto/constants.rs:12-21, used at
// src/elgamal/elgamal.rs:49,85 and src/accounts/accounts.rs:214) can carry far wider windows:
//     W = 16:  16 windows x 32 769 entries x 96 B =  50 MB  (L2 resident)
//     W = 22:  12 windows x  2^21+1             =  2.4 GB
//     W = 26:  10 windows x  2^25+1             = 32.2 GB  (HBM gathers: 10 x 96 B per scalar)
// One scalar multiplication is then NW - 1 mixed additions (7 M each) plus one multiplication to lift the first
// table entry to extended coordinates -- no doublings.  Table entry (k, j) = j * 2^(W k) * Base as affine Niels
// (y+x, y-x, 2dxy), j = 0 .. 2^(W-1), signed digits.  Same recoding as fb_scalarmult (scalarmult.cuh).
//
// The tables are built on the device: running sums per chunk, then ONE batch inversion per window slice
// (k_binv_*, compress_batch.cuh) to normalise Z.
#pragma once
#include "compress_batch.cuh"
#include "msm.cuh"

namespace qq {

struct fbt_geom {
    int W, NW;
    unsigned int ENT;  // 2^(W-1) + 1 entries per window (entry 0 = identity)
};

#define QQ_FBT_CHUNK 128

// ---- table construction ------------------------------------------------------------------------------------------
// bases[k] = 2^(W k) * Base, extended coordinates.  One thread (a chain of W * (NW - 1) doublings).
__global__ void k_fbt_window_bases(const u32x4* __restrict__ base_compressed, fbt_geom g, u32x4* __restrict__ bases) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    u32 w[8];
    load_words32(w, base_compressed, 0);
    ge_p3 p;
    ristretto_decompress(p, w);
    for (int k = 0; k < g.NW; k++) {
        ge_p3_store(bases + QQ_PT_Q * k, p);
        for (int i = 0; i < g.W; i++) ge_dbl<true>(p, p);
    }
}
// window k, entries j in [j0, j0 + cnt): thread c owns QQ_FBT_CHUNK consecutive j; first = j * G_k by double-and-add,
// the rest by repeated addition of G_k.  Writes extended points and their Z (for the batch inversion).
__global__ void __launch_bounds__(128) k_fbt_points(const u32x4* __restrict__ bases, int k, unsigned int j0, unsigned int cnt,
                                                    u32x4* __restrict__ ext, u32x4* __restrict__ zs) {
    unsigned int c = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int lo = c * QQ_FBT_CHUNK;
    if (lo >= cnt) return;
    unsigned int hi = lo + QQ_FBT_CHUNK < cnt ? lo + QQ_FBT_CHUNK : cnt;
    ge_p3 G;
    ge_p3_load(G, bases + QQ_PT_Q * k);
    ge_cached cg;
    ge_to_cached(cg, G);
    ge_p3 r;
    ge_identity(r);
    unsigned int j = j0 + lo;
    for (int b = 31; b >= 0; b--) {
        ge_dbl<true>(r, r);
        if ((j >> b) & 1u) ge_add(r, r, cg);
    }
    for (unsigned int i = lo; i < hi; i++) {
        ge_p3_store(ext + (size_t)QQ_PT_Q * i, r);
        fe_st(zs + 2 * (size_t)i, r.Z);
        ge_add(r, r, cg);
    }
}
// ext[i], zinv[i] -> affine Niels at tbl + 6 * i (16-byte units)
__global__ void __launch_bounds__(256) k_fbt_normalize(const u32x4* __restrict__ ext, const u32x4* __restrict__ zinv,
                                                       unsigned int cnt, u32x4* __restrict__ tbl) {
    unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    ge_p3 p;
    ge_p3_load(p, ext + (size_t)QQ_PT_Q * i);
    fe zi, x, y, t;
    fe_ld(zi, zinv + 2 * (size_t)i);
    fe_mul(x, p.X, zi);
    fe_mul(y, p.Y, zi);
    ge_niels n;
    fe_add(n.ypx, y, x);
    fe_sub(n.ymx, y, x);
    fe_mul(t, x, y);
    fe_mul(n.xy2d, t, fe_2d());
    niels_store_padded(tbl + (size_t)QQ_NIELS_STRIDE_Q * i, n);
}

// ---- the walk ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void fbt_lookup(ge_niels& n, const u32x4* __restrict__ tbl, const fbt_geom& g, const u32 r[9], int k) {
    int d = sc_digit_rt(r, g.W, k);
    u32 neg = d < 0 ? 1u : 0u;
    u32 idx = (u32)(d < 0 ? -d : d);
    niels_load_padded(n, tbl + ((size_t)k * g.ENT + idx) * QQ_NIELS_STRIDE_Q);
    ge_niels_cneg(n, neg);
}
__device__ __forceinline__ void fbt_scalarmult(ge_p3& r, const u32x4* __restrict__ tbl, const fbt_geom& g, const u32 s[8]) {
    u32 rr[9];
    sc_recode_bias_rt(rr, s, g.W, g.NW);
    // pull every table line this scalar will touch towards L2 before the dependent chain starts
    for (int k = 1; k < g.NW; k++) {
        int d = sc_digit_rt(rr, g.W, k);
        u32 idx = (u32)(d < 0 ? -d : d);
        const u32x4* e = tbl + ((size_t)k * g.ENT + idx) * QQ_NIELS_STRIDE_Q;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(e));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(e + 5));
    }
    ge_niels n;
    fbt_lookup(n, tbl, g, rr, 0);
    // affine Niels -> extended with Z = 4: X = 2 (ypx - ymx) = 4x, Y = 2 (ypx + ymx) = 4y, T = (ypx - ymx)(ypx + ymx) = 4xy
    fe dx, sy;
    fe_sub(dx, n.ypx, n.ymx);
    fe_add(sy, n.ypx, n.ymx);
    fe_add(r.X, dx, dx);
    fe_add(r.Y, sy, sy);
    fe_0(r.Z);
    r.Z.v[0] = 4;
    fe_mul(r.T, dx, sy);
#pragma unroll 1
    for (int k = 1; k < g.NW; k++) {
        fbt_lookup(n, tbl, g, rr, k);
        ge_madd(r, r, n);
    }
}
// out[t] = s_t * Base (extended, 128 B) -- for sums with other terms
// (128-thread blocks, 4 per SM, no barrier: the walk is ~25 KB of SASS and gather-latency sensitive; the 512-thread
//  lockstep form of k_varbase measured 3 % slower here)
#define QQ_FBT_BLOCK 128
__global__ void __launch_bounds__(QQ_FBT_BLOCK, 4) k_fixedbase_big(const u32x4* __restrict__ tbl, fbt_geom g,
                                                                   const u32x4* __restrict__ s, int halve,
                                                                   u32x4* __restrict__ out, size_t n) {
    size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t rounds = (n + stride - 1) / stride;
    for (size_t it = 0; it < rounds; it++) {
        size_t t = gtid + it * stride;
        bool live = t < n;
        if (!live) break;
        u32 w[8];
        load_words32(w, s, t);
        if (halve) sc_halve(w, w);
        ge_p3 r;
        fbt_scalarmult(r, tbl, g, w);
        if (live) ge_p3_store(out + QQ_PT_Q * t, r);
    }
}
// fused with the first stage of the batch encoder: Q_t = (s_t / 2) * Base, state_t, w_t, zflag_t  (k_dc_prepare's outputs)
__global__ void __launch_bounds__(QQ_FBT_BLOCK, 4) k_fixedbase_big_dc(const u32x4* __restrict__ tbl, fbt_geom g,
                                                                      const u32x4* __restrict__ s, u32x4* __restrict__ state,
                                                                      u32x4* __restrict__ wout, uint8_t* __restrict__ zflag,
                                                                      size_t n) {
    size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t rounds = (n + stride - 1) / stride;
    for (size_t it = 0; it < rounds; it++) {
        size_t t = gtid + it * stride;
        bool live = t < n;
        if (!live) break;
        u32 w[8];
        load_words32(w, s, t);
        sc_halve(w, w);
        ge_p3 r;
        fbt_scalarmult(r, tbl, g, w);
        dc_state st;
        fe wv, one;
        dc_prepare(st, wv, r);
        u32 z = fe_iszero(wv);
        fe_1(one);
        fe_cmov(wv, one, z);
        if (!live) continue;
        zflag[t] = (uint8_t)z;
        u32x4* sp = state + (size_t)QQ_DC_STATE_Q * t;
        int o = 0;
        fe_store4(sp, o, st.e, st.f);
        fe_store4(sp, o, st.g, st.h);
        fe_store4(sp, o, st.eg, st.fh);
        fe_st(wout + 2 * t, wv);
    }
}

// ---- 64-bit values (balances) ------------------------------------------------------------------------------------------
// The reference builds most committed values with Scalar::from(u64) (balances: src/accounts/accounts.rs:419-429,
// src/elgamal/elgamal.rs:285-300; negative amounts as -Scalar::from(u64)).  A 64-bit magnitude needs ceil(66 / W)
// windows instead of ceil(255 / W): 3 at W = 22.  The batch encoder wants the HALF point; (v / 2 mod l) of an odd v is a
// 252-bit scalar, so instead  v B = 2 ((v >> 1) B + (v & 1) B/2)  with the constant B/2 = ((l + 1) / 2) B added by one
// more mixed addition (identity entry when v is even: uniform control flow).  Negative v: the point is negated.
// out: the batch encoder's first-stage outputs, as k_fixedbase_big_dc.
__global__ void __launch_bounds__(QQ_FBT_BLOCK, 4) k_fixedbase_big_i64_dc(const u32x4* __restrict__ tbl, fbt_geom g,
                                                                          const long long* __restrict__ v,
                                                                          const u32x4* __restrict__ half_base,
                                                                          u32x4* __restrict__ state, u32x4* __restrict__ wout,
                                                                          uint8_t* __restrict__ zflag, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    ge_niels hb;
    niels_load_padded(hb, half_base);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        long long sv = v[t];
        unsigned long long mag = sv < 0 ? 0ull - (unsigned long long)sv : (unsigned long long)sv;
        u32 odd = (u32)(mag & 1ull);
        unsigned long long h = mag >> 1;
        u32 w[8] = {(u32)h, (u32)(h >> 32), 0, 0, 0, 0, 0, 0};
        ge_p3 r;
        fbt_scalarmult(r, tbl, g, w);                 // g.NW = windows covering 64 bits
        ge_niels add;                                 // odd ? B/2 : identity (1, 1, 0)
        fe one, zero;
        fe_1(one);
        fe_0(zero);
        add.ypx = one; add.ymx = one; add.xy2d = zero;
        fe_cmov(add.ypx, hb.ypx, odd);
        fe_cmov(add.ymx, hb.ymx, odd);
        fe_cmov(add.xy2d, hb.xy2d, odd);
        ge_madd(r, r, add);
        if (sv < 0) ge_neg(r, r);
        dc_state st;
        fe wv;
        dc_prepare(st, wv, r);
        u32 z = fe_iszero(wv);
        fe_cmov(wv, one, z);
        zflag[t] = (uint8_t)z;
        u32x4* sp = state + (size_t)QQ_DC_STATE_Q * t;
        int o = 0;
        fe_store4(sp, o, st.e, st.f);
        fe_store4(sp, o, st.g, st.h);
        fe_store4(sp, o, st.eg, st.fh);
        fe_st(wout + 2 * t, wv);
    }
}
// one extended point -> affine Niels (96 B): the constant B/2 above
__global__ void k_point_to_niels(const u32x4* __restrict__ p_ext, u32x4* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    ge_p3 p;
    ge_p3_load(p, p_ext);
    u32 w[QQ_NIELS_WORDS];
    ge_to_niels_affine(w, p);
    ge_niels nl;
    ge_niels_load(nl, w);
    niels_store_padded(out, nl);
}

}  // namespace qq
