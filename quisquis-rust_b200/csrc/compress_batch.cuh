// Batched Ristretto encoding without a square root: enc(2 Q) for many Q with ONE field inversion per batch.
//
// RistrettoPoint::compress costs an inverse square root (a 2^252-3 power, ~254 squarings) per point.  dalek 3.x also
// ships `RistrettoPoint::double_and_compress_batch` (ristretto.rs), which encodes 2Q from Q = (X:Y:Z:T) with
// 4 S + 16 M and one shared inversion (Montgomery's trick).  Every output of the hot path that is a scalar multiple
// s P (or a sum of such) can be produced as 2 ((s/2 mod l) P), because the Ristretto group has odd order l:
//   update_public_key (reference src/ristretto/keys.rs:146-148), generate_commitment (src/elgamal/elgamal.rs:41-53),
//   the pk half of update_account (src/accounts/accounts.rs:146-148), create_delta_and_epsilon_accounts (:198-220),
//   verify_account (:81-84) and the fixed-base batches.
// The encoding is canonical, so the bytes equal compress(s P) exactly (tests: test_host_arith.py, test_gpu_parity.py).
//
// Pipeline (all data-parallel, HBM-resident intermediates):
//   k_dc_prepare : per item, sum the sources, state = (e, f, g, h, eg, fh), w = eg fh (zero -> 1 + flag)
//   k_binv_up    : per chunk of C values: exclusive prefix products + chunk total     } repeated over
//   k_binv_top   : one thread: inverts the (<= C) totals of the last level            } ceil(log_C n)
//   k_binv_down  : per chunk: inv_i = run * prefix_i, run *= w_i                      } levels
//   k_dc_finish  : per item, the sign fix-ups of the encoder and the final products -> 32 bytes (or a comparison)
// Field inversions: exactly one per call, whatever n.
#pragma once
#if defined(__CUDACC__)
#include "kernels.cuh"
#else
#include "ristretto.cuh"
#include "scalarmult.cuh"
#endif

namespace qq {

struct dc_state {
    fe e, f, g, h, eg, fh;
};
#define QQ_DC_STATE_Q 12  // 6 field elements x 2 x 16 B

// Q -> state; returns w = eg * fh  (zero exactly when 2Q is in the identity class)
QQ_HD void dc_prepare(dc_state& st, fe& w, const ge_p3& q) {
    fe xx, yy, zz, dtt, t;
    fe_sq(xx, q.X);
    fe_sq(yy, q.Y);
    fe_sq(zz, q.Z);
    fe_sq(t, q.T);
    fe_mul(dtt, t, fe_d());
    fe_add(t, q.Y, q.Y);
    fe_mul(st.e, q.X, t);        // 2XY
    fe_add(st.f, zz, dtt);       // Z^2 + d T^2
    fe_add(st.g, yy, xx);        // Y^2 + X^2
    fe_sub(st.h, zz, dtt);       // Z^2 - d T^2
    fe_mul(st.eg, st.e, st.g);
    fe_mul(st.fh, st.f, st.h);
    fe_mul(w, st.eg, st.fh);
}
// state + inv = 1 / (eg fh) -> canonical encoding of 2Q
QQ_HD void dc_finish(u32 out[8], const dc_state& st, const fe& inv) {
    fe zinv, tinv, t, e, g, h, magic, me, fs;
    fe_mul(zinv, st.eg, inv);    // 1 / (f h)
    fe_mul(tinv, st.fh, inv);    // 1 / (e g)
    fe_mul(t, st.eg, zinv);
    u32 n1 = fe_isnegative(t);
    e = st.e;
    g = st.g;
    h = st.h;
    magic = fe_invsqrt_a_minus_d();
    fe_neg(me, st.e);
    fe_mul(fs, st.f, fe_sqrt_m1());
    fe_cmov(e, st.g, n1);
    fe_cmov(g, me, n1);
    fe_cmov(h, fs, n1);
    fe_cmov(magic, fe_sqrt_m1(), n1);
    fe_mul(t, h, e);
    fe_mul(t, t, zinv);
    fe_cneg(g, fe_isnegative(t));
    fe_sub(t, h, g);
    fe s;
    fe_mul(s, g, tinv);
    fe_mul(s, magic, s);
    fe_mul(s, t, s);
    fe_abs(s);
    fe_towords(out, s);
}

#if defined(__CUDACC__)
__device__ __forceinline__ void fe_ld(fe& a, const u32x4* src) {
    int o = 0;
    fe_load(src, o, a);
}
__device__ __forceinline__ void fe_st(u32x4* dst, const fe& a) {
    int o = 0;
    fe_store(dst, o, a);
}

__global__ void __launch_bounds__(256) k_dc_prepare(fin_args a, u32x4* __restrict__ state, u32x4* __restrict__ w,
                                                    uint8_t* __restrict__ zflag) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += stride) {
        ge_p3 q;
        fin_eval(q, a, t);
        dc_state st;
        fe wv, one;
        dc_prepare(st, wv, q);
        u32 z = fe_iszero(wv);
        fe_1(one);
        fe_cmov(wv, one, z);
        zflag[t] = (uint8_t)z;
        u32x4* s = state + (size_t)QQ_DC_STATE_Q * t;
        int o = 0;
        fe_store4(s, o, st.e, st.f);
        fe_store4(s, o, st.g, st.h);
        fe_store4(s, o, st.eg, st.fh);
        fe_st(w + 2 * t, wv);
    }
}

// chunk c = values [c C, min(n, (c+1) C)): prefix[i] = product of the chunk's values before i, totals[c] = chunk product
__global__ void __launch_bounds__(128) k_binv_up(const u32x4* __restrict__ vals, size_t n, int C,
                                                 u32x4* __restrict__ prefix, u32x4* __restrict__ totals) {
    size_t nchunks = (n + C - 1) / C;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += stride) {
        size_t lo = c * (size_t)C, hi = lo + C < n ? lo + C : n;
        fe acc;
        fe_1(acc);
        for (size_t i = lo; i < hi; i++) {
            fe v;
            fe_ld(v, vals + 2 * i);
            fe_st(prefix + 2 * i, acc);
            fe_mul(acc, acc, v);
        }
        fe_st(totals + 2 * c, acc);
    }
}
// single thread: vals[0..n) (n small) -> inv[i] = 1 / vals[i], one field inversion
__global__ void k_binv_top(const u32x4* __restrict__ vals, size_t n, u32x4* scratch_prefix, u32x4* inv) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    fe acc;
    fe_1(acc);
    for (size_t i = 0; i < n; i++) {
        fe v;
        fe_ld(v, vals + 2 * i);
        fe_st(scratch_prefix + 2 * i, acc);
        fe_mul(acc, acc, v);
    }
    fe run;
    fe_invert(run, acc);
    for (size_t i = n; i-- > 0;) {
        fe v, p, r;
        fe_ld(v, vals + 2 * i);
        fe_ld(p, scratch_prefix + 2 * i);
        fe_mul(r, run, p);
        fe_st(inv + 2 * i, r);
        fe_mul(run, run, v);
    }
}
// chunk c: run = inv_totals[c]; for i from the end: inv[i] = run * prefix[i]; run *= vals[i].   inv may alias prefix.
__global__ void __launch_bounds__(128) k_binv_down(const u32x4* __restrict__ vals, const u32x4* prefix,
                                                   const u32x4* __restrict__ inv_totals, size_t n, int C, u32x4* inv) {
    size_t nchunks = (n + C - 1) / C;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += stride) {
        size_t lo = c * (size_t)C, hi = lo + C < n ? lo + C : n;
        fe run;
        fe_ld(run, inv_totals + 2 * c);
        for (size_t i = hi; i-- > lo;) {
            fe v, p, r;
            fe_ld(v, vals + 2 * i);
            fe_ld(p, prefix + 2 * i);
            fe_mul(r, run, p);
            fe_st(inv + 2 * i, r);
            fe_mul(run, run, v);
        }
    }
}

// out[omap(t)] = enc(2 Q_t), zeros when the item is bad or 2 Q_t is the identity class.  With `expect` != nullptr
// nothing is written to out; instead flag[t] = (enc == expect[emap(t)]).
__global__ void __launch_bounds__(256) k_dc_finish(const u32x4* __restrict__ state, const u32x4* __restrict__ inv,
                                                   const uint8_t* __restrict__ zflag, const uint8_t* __restrict__ bad,
                                                   int bdiv, u32x4* __restrict__ out, idx_map omap,
                                                   const u32x4* __restrict__ expect, idx_map emap,
                                                   uint8_t* __restrict__ flag, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        dc_state st;
        const u32x4* s = state + (size_t)QQ_DC_STATE_Q * t;
        int o = 0;
        fe_load4(s, o, st.e, st.f);
        fe_load4(s, o, st.g, st.h);
        fe_load4(s, o, st.eg, st.fh);
        fe iv;
        fe_ld(iv, inv + 2 * t);
        u32 w[8];
        dc_finish(w, st, iv);
        if (zflag[t] != 0 || (bad != nullptr && bad[t / (size_t)bdiv] != 0)) {
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = 0;
        }
        if (expect != nullptr) {
            u32 e[8];
            load_words32(e, expect, map_index(emap, t));
            u32 d = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) d |= w[i] ^ e[i];
            flag[t] = d == 0 ? 1 : 0;
        } else {
            store_words32(out, map_index(omap, t), w);
        }
    }
}
// Small batches: the five dependent launches above are pure latency (one inversion deep each way), so every item
// inverts its own w -- prepare, fe_invert and finish in ONE kernel, one thread per item, 32-thread blocks spread over
// the SMs.  Same outputs as k_dc_prepare / k_binv_* / k_dc_finish.
__global__ void __launch_bounds__(32) k_dc_direct(fin_args a, const uint8_t* __restrict__ bad, int bdiv,
                                                  u32x4* __restrict__ out, idx_map omap,
                                                  const u32x4* __restrict__ expect, idx_map emap,
                                                  uint8_t* __restrict__ flag) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.n) return;
    ge_p3 q;
    fin_eval(q, a, t);
    dc_state st;
    fe wv, one, iv;
    dc_prepare(st, wv, q);
    u32 z = fe_iszero(wv);
    fe_1(one);
    fe_cmov(wv, one, z);
    fe_invert(iv, wv);
    u32 w[8];
    dc_finish(w, st, iv);
    if (z != 0 || (bad != nullptr && bad[t / (size_t)bdiv] != 0)) {
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = 0;
    }
    if (expect != nullptr) {
        u32 e[8];
        load_words32(e, expect, map_index(emap, t));
        u32 d = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) d |= w[i] ^ e[i];
        flag[t] = d == 0 ? 1 : 0;
    } else {
        store_words32(out, map_index(omap, t), w);
    }
}
#endif

}  // namespace qq
