// Bulletproofs range-proof batch verification: the generator scalars of the verification equation, folded over a batch
// of proofs ON THE DEVICE.
//
// RangeProof::verify_multiple (bulletproofs crate; called by the reference at src/accounts/verifier.rs:517,548) ends in one
// multiscalar multiplication whose 2 n m generator scalars are, for generator index i of proof p,
//     g_i = -z - a s_i                        h_i = z + y^-i (z^2 z^(i / n) 2^(i mod n) - b s_(nm-1-i))
// with s_i = (u_1 .. u_k)^-1 prod_{bit j of i set} u_(k-j)^2.  A batch of proofs shares the generators, so with one random
// weight rho_p per proof the whole batch needs sum_p rho_p g_(p,i) and sum_p rho_p h_(p,i): nm x proofs small products in
// Z/l -- data-parallel integer work that the host threads would spend 0.25 ms per proof on.  One thread per generator
// index walks a chunk of the proofs (records are read as warp-wide broadcasts), partial sums per chunk are added by a
// second kernel straight into the scalar array of the aggregated MSM.
#pragma once
#include "kernels.cuh"
#include "sc_host.hpp"
#include "rangeproof_verify.cuh"

namespace qq {

typedef qq_rp::record rp_record;      // written by the transcript phase (rangeproof_verify.cuh)

// Per (sub-)proof product tables: s_i and y^-j are products over the set bits of the index, so with the index cut into a low
// part (LB = min(5, lg) bits) and a high part,  s_i = allinv SHi[hi] SLo[lo]  and  y^-j = YHi[jh] YLo[jl]:  five tables of at
// most 32 entries per proof (160 short products, built once by k_rp_tables) replace the lg products every (generator, proof)
// pair paid before - 5 instead of lg + 4 = 14 products of Z/l per pair at n m = 1024.  a and b are folded into the high tables.
#define QQ_RP_TBL 32
struct rp_tables {
    qq_sc::sc sa[QQ_RP_TBL];      // rho a allinv SHi[hi]
    qq_sc::sc sb[QQ_RP_TBL];      // rho b allinv SHi[hi]
    qq_sc::sc slo[QQ_RP_TBL];     // SLo[lo]
    qq_sc::sc yhi[QQ_RP_TBL];     // YHi[jh]
    qq_sc::sc ylo[QQ_RP_TBL];     // YLo[jl]
};
// grid = number of (sub-)proofs, block = 5 x 32 threads: thread (table t, entry e)
__global__ void __launch_bounds__(160) k_rp_tables(const rp_record* __restrict__ rec, int lg, rp_tables* __restrict__ tbl) {
    const rp_record& r = rec[blockIdx.x];
    const int t = threadIdx.x / QQ_RP_TBL, e = threadIdx.x % QQ_RP_TBL;
    const int LB = lg < 5 ? lg : 5, HB = lg - LB;
    qq_sc::sc v = qq_sc::one();
    if (t == 0 || t == 1) {
        v = qq_sc::mul(t == 0 ? r.ra : r.rb, r.allinv);
        for (int k = 0; k < HB; k++)
            if ((e >> k) & 1) v = qq_sc::mul(v, r.usq[lg - 1 - (LB + k)]);
    } else if (t == 2) {
        for (int k = 0; k < LB; k++)
            if ((e >> k) & 1) v = qq_sc::mul(v, r.usq[lg - 1 - k]);
    } else if (t == 3) {
        for (int k = 0; k < HB; k++)
            if ((e >> k) & 1) v = qq_sc::mul(v, r.yinv_pow[LB + k]);
    } else {
        for (int k = 0; k < LB; k++)
            if ((e >> k) & 1) v = qq_sc::mul(v, r.yinv_pow[k]);
    }
    rp_tables& o = tbl[blockIdx.x];
    (t == 0 ? o.sa : t == 1 ? o.sb : t == 2 ? o.slo : t == 3 ? o.yhi : o.ylo)[e] = v;
}
// The points of every (sub-)proof's T MSM terms, straight from the proof and commitment bytes (the layout transcript_phase
// writes: A, S, T_1, T_2 | L_k | R_k | V_j): done before the transcripts so that their decompression can run beside them.
__global__ void __launch_bounds__(256) k_rp_points(const uint8_t* __restrict__ proofs, const uint8_t* __restrict__ commitments, size_t nsub,
                                                   unsigned int T, unsigned int lg, unsigned int m, unsigned int proof_bytes,
                                                   uint8_t* __restrict__ p_out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < nsub * T; t += stride) {
        const size_t q = t / T;
        const unsigned int s = (unsigned int)(t - q * T);
        const uint8_t* pr = proofs + q * proof_bytes;
        const uint8_t* src = s < 4 ? pr + 32 * s
                           : s < 4 + lg ? pr + 224 + 64 * (s - 4)
                           : s < 4 + 2 * lg ? pr + 224 + 64 * (s - 4 - lg) + 32
                           : commitments + 32 * ((size_t)m * q + (s - 4 - 2 * lg));
        const uint4 a = reinterpret_cast<const uint4*>(src)[0], b = reinterpret_cast<const uint4*>(src)[1];
        reinterpret_cast<uint4*>(p_out + 32 * t)[0] = a;
        reinterpret_cast<uint4*>(p_out + 32 * t)[1] = b;
    }
}
// grid = (ceil(N / block), chunks); partial: chunks x 2N scalars (g sums then h sums)
// chunk_first == nullptr: the count (sub-)proofs from `first` on are split evenly over the chunks.  Otherwise (the grouped form:
// one chunk per group, its partial sums are the group's generator scalars) chunk y covers the per_chunk (sub-)proofs from
// chunk_first[y] on, cut at first + count.
__global__ void __launch_bounds__(128) k_rp_fold(const rp_record* __restrict__ rec, const rp_tables* __restrict__ tbl, unsigned int first,
                                                 unsigned int count, int n_bits, int N, int lg, qq_sc::sc* __restrict__ partial,
                                                 unsigned int per_chunk, const unsigned int* __restrict__ chunk_first) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    unsigned int p0, p1;
    if (chunk_first != nullptr) {
        const unsigned int base = chunk_first[blockIdx.y], end = first + count;
        p0 = 0;
        p1 = base >= end ? 0 : (end - base < per_chunk ? end - base : per_chunk);
        first = base;
    } else {
        unsigned int per = (count + gridDim.y - 1) / gridDim.y;
        p0 = blockIdx.y * per;
        p1 = p0 + per < count ? p0 + per : count;
    }
    // thread i owns g_i and h_j with j = nm - 1 - i: both need the same s_i
    qq_sc::sc sum_g = qq_sc::zero(), sum_h = qq_sc::zero();
    const int j = N - 1 - i;
    const int party = j / n_bits;
    const int LB = lg < 5 ? lg : 5, lmask = (1 << LB) - 1;
    const int lo = i & lmask, hi = i >> LB, jl = j & lmask, jh = j >> LB;
    // sum_g = sum_p (neg_rz - sa slo),  sum_h = sum_p (rz + y_j (rzz_zj 2^k - sb slo)): the products sa slo and y_j u are added
    // up unreduced (17 limbs) and reduced once per thread; u itself needs one reduction (it is a factor of the next product),
    // y_j one - two reductions per (generator, proof) pair instead of five
    uint32_t accg[17], acch[17], x[16];
    for (int q = 0; q < 17; q++) accg[q] = acch[q] = 0;
    const int kbit = j % n_bits;
    for (unsigned int p = p0; p < p1; p++) {
        const rp_record& r = rec[first + p];
        const rp_tables& t = tbl[first + p];
        const qq_sc::sc slo = t.slo[lo];
        const qq_sc::sc y_j = qq_sc::mul(t.yhi[jh], t.ylo[jl]);
        sum_g = qq_sc::add(sum_g, r.neg_rz);
        qq_sc::mul_wide_w32(x, t.sa[hi], slo);
        qq_sc::acc17_add(accg, x);
        qq_sc::mul_wide_w32(x, t.sb[hi], slo);
        const qq_sc::sc u = qq_sc::shl_minus_wide(r.rzz_zj[party], kbit, x);
        sum_h = qq_sc::add(sum_h, r.rz);
        qq_sc::mul_wide_w32(x, y_j, u);
        qq_sc::acc17_add(acch, x);
    }
    sum_g = qq_sc::sub(sum_g, qq_sc::acc17_reduce(accg));
    sum_h = qq_sc::add(sum_h, qq_sc::acc17_reduce(acch));
    partial[(size_t)blockIdx.y * 2 * N + i] = sum_g;
    partial[(size_t)blockIdx.y * 2 * N + N + j] = sum_h;
}
// out[t] = sum over chunks of partial[chunk][t], t < 2N, written as the 32-byte scalars of the MSM: one block per t, the
// chunks strided over its threads, tree sum in shared memory (few generators and thousands of chunks when n m is small)
__global__ void __launch_bounds__(128) k_rp_fold_sum(const qq_sc::sc* __restrict__ partial, int chunks, int twoN,
                                                     qq_sc::sc* __restrict__ out) {
    __shared__ qq_sc::sc sh[128];
    const int t = blockIdx.x;
    qq_sc::sc s = qq_sc::zero();
    for (int c = threadIdx.x; c < chunks; c += blockDim.x) s = qq_sc::add(s, partial[(size_t)c * twoN + t]);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 64; w >= 1; w >>= 1) {
        if ((int)threadIdx.x < w) sh[threadIdx.x] = qq_sc::add(sh[threadIdx.x], sh[threadIdx.x + w]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[t] = sh[0];
}

// Grouped form of the aggregate (the failure path: which groups of transcripts do not verify?): the term list
// [groups x (2N + 2) shared terms | groups x terms_per_group proof terms] with the group of every term.  partial: groups x 2N
// generator scalars (k_rp_fold, one chunk per group), shared: groups x 2 (B_blinding, B), gen: the 2N + 2 shared points.  The
// proof terms of group g are gathered from src_sc / src_pt (T terms per (sub-)proof) starting at (sub-)proof chunk_first[g];
// slots beyond `end_sub` (a short last group) are padded with 0 * B.
__global__ void k_rp_group_terms(const qq_sc::sc* __restrict__ partial, const qq_sc::sc* __restrict__ shared, const uint8_t* __restrict__ gen,
                                 unsigned int groups, unsigned int twoN, const unsigned int* __restrict__ chunk_first, unsigned int end_sub,
                                 unsigned int T, size_t terms_per_group, const uint8_t* __restrict__ src_sc, const uint8_t* __restrict__ src_pt,
                                 uint8_t* __restrict__ sc, uint8_t* __restrict__ pt, unsigned int* __restrict__ group_of) {
    const size_t G = (size_t)twoN + 2, nshared = (size_t)groups * G, total = nshared + (size_t)groups * terms_per_group;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        uint4* ds = reinterpret_cast<uint4*>(sc + 32 * t);
        uint4* dp = reinterpret_cast<uint4*>(pt + 32 * t);
        if (t < nshared) {
            const size_t g = t / G, j = t - g * G;
            reinterpret_cast<qq_sc::sc*>(sc)[t] = j < twoN ? partial[g * twoN + j] : shared[2 * g + (j - twoN)];
            const uint4* src = reinterpret_cast<const uint4*>(gen + 32 * j);
            const uint4 p0 = src[0], p1 = src[1];
            dp[0] = p0;
            dp[1] = p1;
            group_of[t] = (unsigned int)g;
        } else {
            const size_t u = t - nshared, g = u / terms_per_group, within = u - g * terms_per_group;
            const size_t st = (size_t)chunk_first[g] * T + within;
            const bool live = st < (size_t)end_sub * T;
            const uint4* ss = reinterpret_cast<const uint4*>(src_sc + 32 * st);
            const uint4* sp = reinterpret_cast<const uint4*>((live ? src_pt + 32 * st : gen + 32 * ((size_t)twoN + 1)));
            const uint4 s0 = live ? ss[0] : z, s1 = live ? ss[1] : z, p0 = sp[0], p1 = sp[1];
            ds[0] = s0;
            ds[1] = s1;
            dp[0] = p0;
            dp[1] = p1;
            group_of[t] = (unsigned int)g;
        }
    }
}

}  // namespace qq
