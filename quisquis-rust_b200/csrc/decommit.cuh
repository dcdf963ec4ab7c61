// ElGamalCommitment::decommit / decommit_value (reference src/elgamal/elgamal.rs:106-122, brute_force_decrypt :169-182).
//
// decommit:       G*v = d - sk * c                      (one variable-base multiplication + one subtraction)
// decommit_value: the reference walks v = 0, 1, 2, ... comparing v*B with G*v -- exponential in the bit length and
//                 strictly sequential.  Here it is baby-step / giant-step over the encodings:
//                   baby table  enc(j B), j < 2^20, hashed by the first 8 bytes (built once per context, 40 MB);
//                   giant steps Q_i = G*v - (i 2^20) B for i < 2^(bits - 20), every i in parallel: fixed-base table walk,
//                   subtraction, encoding, table probe;  v = i 2^20 + j.
//                 The smallest v wins (atomicMin), as in the reference's ascending loop.
#pragma once
#include "kernels.cuh"

namespace qq {

#define QQ_BSGS_BABY_BITS 20
#define QQ_BSGS_SLOT_BITS 22   // open addressing, load factor 1/4
#define QQ_BSGS_EMPTY 0xffffffffu

// scalar_i = (first + i) << shift  as 32 little-endian bytes
__global__ void k_bsgs_scalars(u32x4* __restrict__ out, size_t n, unsigned long long first, int shift) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned long long v = first + i;
        u32 w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        unsigned long long lo = shift ? (v << shift) : v;
        unsigned long long hi = shift ? (v >> (64 - shift)) : 0ull;
        w[0] = (u32)lo; w[1] = (u32)(lo >> 32); w[2] = (u32)hi; w[3] = (u32)(hi >> 32);
        store_words32(out, i, w);
    }
}
__device__ __forceinline__ u32 bsgs_hash(u32 a, u32 b) {
    unsigned long long k = ((unsigned long long)b << 32) | a;
    k *= 0x9e3779b97f4a7c15ull;
    return (u32)(k >> (64 - QQ_BSGS_SLOT_BITS));
}
__global__ void k_bsgs_insert(const u32x4* __restrict__ baby_enc, u32 n, u32* __restrict__ slots) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    u32 w[8];
    load_words32(w, baby_enc, j);
    u32 h = bsgs_hash(w[0], w[1]);
    const u32 mask = (1u << QQ_BSGS_SLOT_BITS) - 1u;
    while (atomicCAS(&slots[h], QQ_BSGS_EMPTY, j) != QQ_BSGS_EMPTY) h = (h + 1) & mask;
}
// enc[i] = encoding of Q_i;  match with baby entry j  ->  candidate (first + i) 2^20 + j
__global__ void k_bsgs_lookup(const u32x4* __restrict__ enc, size_t n, unsigned long long first,
                              const u32x4* __restrict__ baby_enc, const u32* __restrict__ slots,
                              unsigned long long* __restrict__ best) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    const u32 mask = (1u << QQ_BSGS_SLOT_BITS) - 1u;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u32 w[8];
        load_words32(w, enc, i);
        u32 h = bsgs_hash(w[0], w[1]);
        for (;;) {
            u32 j = slots[h];
            if (j == QQ_BSGS_EMPTY) break;
            u32 b[8];
            load_words32(b, baby_enc, j);
            u32 d = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) d |= b[k] ^ w[k];
            if (d == 0) {
                atomicMin(best, ((first + i) << QQ_BSGS_BABY_BITS) + j);
                break;
            }
            h = (h + 1) & mask;
        }
    }
}
// extended (not encoded) result of a finish job: out[t] = sum of the sources
__global__ void __launch_bounds__(256) k_finish_points(fin_args a, u32x4* __restrict__ out_pts) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += stride) {
        ge_p3 q;
        fin_eval(q, a, t);
        ge_p3_store(out_pts + QQ_PT_Q * t, q);
    }
}
__global__ void k_bsgs_results(const unsigned long long* __restrict__ best, const uint8_t* __restrict__ pre_status,
                               unsigned long long* __restrict__ values, uint8_t* __restrict__ status, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t st = pre_status[i];
    unsigned long long b = best[i];
    if (st == 0 && b == ~0ull) st = 5;   // QQ_ST_NOT_FOUND
    values[i] = st ? 0ull : b;
    status[i] = st;
}

}  // namespace qq
