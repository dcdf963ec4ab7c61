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

// grid = (ceil(N / block), chunks); partial: chunks x 2N scalars (g sums then h sums)
__global__ void __launch_bounds__(128) k_rp_fold(const rp_record* __restrict__ rec, unsigned int first, unsigned int count,
                                                 int n_bits, int N, int lg, qq_sc::sc* __restrict__ partial) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    unsigned int per = (count + gridDim.y - 1) / gridDim.y;
    unsigned int p0 = blockIdx.y * per, p1 = p0 + per < count ? p0 + per : count;
    // thread i owns g_i and h_j with j = nm - 1 - i: both need the same s_i
    qq_sc::sc sum_g = qq_sc::zero(), sum_h = qq_sc::zero();
    const int j = N - 1 - i;
    qq_sc::sc two_pow = qq_sc::zero();
    two_pow.v[0] = 1ull << (j % n_bits);
    const int party = j / n_bits;
    for (unsigned int p = p0; p < p1; p++) {
        const rp_record& r = rec[first + p];
        qq_sc::sc s_i = r.allinv, y_j = qq_sc::one();
        for (int k = 0; k < lg; k++) {
            if ((i >> k) & 1) s_i = qq_sc::mul(s_i, r.usq[lg - 1 - k]);
            else y_j = qq_sc::mul(y_j, r.yinv_pow[k]);          // bit k of j is the complement of bit k of i
        }
        sum_g = qq_sc::add(sum_g, qq_sc::sub(r.neg_rz, qq_sc::mul(r.ra, s_i)));
        qq_sc::sc t = qq_sc::sub(qq_sc::mul(r.rzz_zj[party], two_pow), qq_sc::mul(r.rb, s_i));
        sum_h = qq_sc::add(sum_h, qq_sc::add(r.rz, qq_sc::mul(y_j, t)));
    }
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

}  // namespace qq
