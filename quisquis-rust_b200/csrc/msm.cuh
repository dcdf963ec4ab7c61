// Large multiscalar multiplication: Pippenger bucket method on the GPU.
//
// Replaces curve25519-dalek 3.x backend/serial/scalar_mul/pippenger.rs (reached from
// `RistrettoPoint::optional_multiscalar_mul`, reference src/accounts/verifier.rs:91-99, and from the Bulletproofs
// verification MSM called at src/accounts/verifier.rs:517,548).  Window size and bucket layout are free choices
// (the result is a canonical encoding); the `None`-on-bad-point behaviour is kept through per-term status.
//
// Pipeline for n terms, signed c-bit windows, K = ceil(256 / c) windows, NB = 2^(c-1) buckets per window:
//   k_msm_prepare     decompress point -> affine Niels (96 B, Z = 1), status, K signed digits, bucket histogram
//   k_scan_*          bucket offsets (tile totals, scan of totals, apply)
//   k_msm_scatter     counting sort of (term, sign) pairs by (window, bucket)
//   k_msm_order_*     buckets ordered by population (descending) so the lanes of a warp get equal work
//   k_msm_accumulate  one thread per bucket: mixed additions (7 M) of its points, 128-bit gathers of Niels points
//   k_msm_reduce_seg  per (window, 32-bucket segment): running-sum trick, then + base * segment total
//   k_point_sum_rows  per window: tree sum of the segment results (shared memory)
//   k_msm_horner      sum_k 2^(c k) W_k
#pragma once
#include "kernels.cuh"

namespace qq {

#define QQ_NIELS_STRIDE_Q 6  // stored affine-Niels point = 3 field elements = 6 x 16 B

__device__ __forceinline__ void niels_store_padded(u32x4* dst, const ge_niels& n) {
    int o = 0;
    fe_store(dst, o, n.ypx);
    fe_store(dst, o, n.ymx);
    fe_store(dst, o, n.xy2d);
}
__device__ __forceinline__ void niels_load_padded(ge_niels& n, const u32x4* src) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4 q;
    q = __ldg(s + 0); n.ypx.v[0] = q.x; n.ypx.v[1] = q.y; n.ypx.v[2] = q.z; n.ypx.v[3] = q.w;
    q = __ldg(s + 1); n.ypx.v[4] = q.x; n.ypx.v[5] = q.y; n.ypx.v[6] = q.z; n.ypx.v[7] = q.w;
    q = __ldg(s + 2); n.ymx.v[0] = q.x; n.ymx.v[1] = q.y; n.ymx.v[2] = q.z; n.ymx.v[3] = q.w;
    q = __ldg(s + 3); n.ymx.v[4] = q.x; n.ymx.v[5] = q.y; n.ymx.v[6] = q.z; n.ymx.v[7] = q.w;
    q = __ldg(s + 4); n.xy2d.v[0] = q.x; n.xy2d.v[1] = q.y; n.xy2d.v[2] = q.z; n.xy2d.v[3] = q.w;
    q = __ldg(s + 5); n.xy2d.v[4] = q.x; n.xy2d.v[5] = q.y; n.xy2d.v[6] = q.z; n.xy2d.v[7] = q.w;
}

// runtime-width signed recoding (see sc_recode_bias): r = s + sum_k 2^(c k + c - 1)
__device__ __forceinline__ void sc_recode_bias_rt(u32 r[9], const u32 s[8], int c, int nw) {
    u32 cst[9];
#pragma unroll
    for (int i = 0; i < 9; i++) cst[i] = 0;
    for (int k = 0; k < nw; k++) {
        int bit = c * k + (c - 1);
#pragma unroll
        for (int i = 0; i < 9; i++)
            if ((bit >> 5) == i) cst[i] |= 1u << (bit & 31);
    }
    u32 carry = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        u64 t = (u64)(i < 8 ? s[i] : 0u) + cst[i] + carry;
        r[i] = (u32)t;
        carry = (u32)(t >> 32);
    }
}
__device__ __forceinline__ int sc_digit_rt(const u32 r[9], int c, int k) {
    int bit = c * k;
    int wi = bit >> 5, sh = bit & 31;
    u32 lo = 0, hi = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        lo = (i == wi) ? r[i] : lo;
        hi = (i == wi + 1) ? r[i] : hi;
    }
    u64 two = (u64)lo | ((u64)hi << 32);
    return (int)((u32)(two >> sh) & ((1u << c) - 1u)) - (1 << (c - 1));
}

struct msm_geom {
    int c, K, NB;  // window bits, windows, buckets per window
};

__global__ void __launch_bounds__(256) k_msm_prepare(const u32x4* __restrict__ points, const u32x4* __restrict__ scalars,
                                                     size_t n, msm_geom g, u32x4* __restrict__ niels,
                                                     uint8_t* __restrict__ term_status, int16_t* __restrict__ digits,
                                                     unsigned int* __restrict__ counts) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u32 w[8];
        load_words32(w, points, i);
        ge_p3 p;
        u32 ok = ristretto_decompress(p, w);
        ge_niels nl;
        ge_to_niels_z1(nl, p);
        niels_store_padded(niels + (size_t)QQ_NIELS_STRIDE_Q * i, nl);
        u32 s[8];
        load_words32(s, scalars, i);
        u32 canon = sc_is_canonical(s);
        uint8_t st = canon ? (ok ? 0 : 1) : 2;
        term_status[i] = st;
        u32 r[9];
        sc_recode_bias_rt(r, s, g.c, g.K);
        for (int k = 0; k < g.K; k++) {
            int d = st ? 0 : sc_digit_rt(r, g.c, k);
            digits[(size_t)k * n + i] = (int16_t)d;
            if (d != 0) atomicAdd(&counts[(size_t)k * g.NB + (size_t)((d < 0 ? -d : d) - 1)], 1u);
        }
    }
}

// ---- exclusive scan of `total` counters -----------------------------------------------------------------------------
// Tile = 4096 counters per 1024-thread block (4 per thread, warp-shuffle scan).  PASS 0 writes each tile's total,
// PASS 1 (one block) scans the tile totals in place, PASS 2 rescans every tile and adds its tile offset.
__device__ __forceinline__ unsigned int block_scan_1024(unsigned int tsum, unsigned int* warp_sums, unsigned int& block_total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned int x = tsum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        unsigned int y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        unsigned int w = warp_sums[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            unsigned int y = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += y;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    block_total = warp_sums[31];
    unsigned int excl = (wid ? warp_sums[wid - 1] : 0u) + (x - tsum);
    __syncthreads();
    return excl;  // exclusive prefix of this thread's tsum within the block
}
__global__ void __launch_bounds__(1024) k_scan_tile_totals(const unsigned int* __restrict__ in, size_t total,
                                                           unsigned int* __restrict__ tile_tot) {
    __shared__ unsigned int warp_sums[32];
    size_t i0 = (size_t)blockIdx.x * 4096 + (size_t)threadIdx.x * 4;
    unsigned int tsum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) tsum += (i0 + j < total) ? in[i0 + j] : 0u;
    unsigned int bt;
    block_scan_1024(tsum, warp_sums, bt);
    if (threadIdx.x == 0) tile_tot[blockIdx.x] = bt;
}
// single block: exclusive scan of up to 4096 values in place
__global__ void __launch_bounds__(1024) k_scan_small(unsigned int* __restrict__ v, size_t n) {
    __shared__ unsigned int warp_sums[32];
    size_t i0 = (size_t)threadIdx.x * 4;
    unsigned int x[4], tsum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { x[j] = (i0 + j < n) ? v[i0 + j] : 0u; tsum += x[j]; }
    unsigned int bt;
    unsigned int excl = block_scan_1024(tsum, warp_sums, bt);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (i0 + j < n) v[i0 + j] = excl;
        excl += x[j];
    }
}
__global__ void __launch_bounds__(1024) k_scan_apply(const unsigned int* __restrict__ in, size_t total,
                                                     const unsigned int* __restrict__ tile_off,
                                                     unsigned int* __restrict__ out) {
    __shared__ unsigned int warp_sums[32];
    size_t i0 = (size_t)blockIdx.x * 4096 + (size_t)threadIdx.x * 4;
    unsigned int x[4], tsum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { x[j] = (i0 + j < total) ? in[i0 + j] : 0u; tsum += x[j]; }
    unsigned int bt;
    unsigned int excl = block_scan_1024(tsum, warp_sums, bt) + tile_off[blockIdx.x];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (i0 + j < total) out[i0 + j] = excl;
        excl += x[j];
    }
}
// host-side helper: exclusive scan of `total` <= 4096 * 4096 counters; tile_tmp holds ceil(total / 4096) words
static inline void launch_scan_exclusive(const unsigned int* in, unsigned int* out, size_t total, unsigned int* tile_tmp,
                                         cudaStream_t st) {
    unsigned tiles = (unsigned)((total + 4095) / 4096);
    k_scan_tile_totals<<<tiles, 1024, 0, st>>>(in, total, tile_tmp);
    k_scan_small<<<1, 1024, 0, st>>>(tile_tmp, tiles);
    k_scan_apply<<<tiles, 1024, 0, st>>>(in, total, tile_tmp, out);
}

__global__ void __launch_bounds__(256) k_msm_scatter(const int16_t* __restrict__ digits, size_t n, msm_geom g,
                                                     const unsigned int* __restrict__ offsets,
                                                     unsigned int* __restrict__ cursor, unsigned int* __restrict__ sorted) {
    size_t total = (size_t)g.K * n;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        int d = digits[t];
        if (d == 0) continue;
        size_t k = t / n;
        size_t i = t - k * n;
        size_t key = k * g.NB + (size_t)((d < 0 ? -d : d) - 1);
        unsigned int slot = atomicAdd(&cursor[key], 1u);
        sorted[offsets[key] + slot] = (unsigned int)i | (d < 0 ? 0x80000000u : 0u);
    }
}

// ---- bucket ordering by population (descending), counting sort on min(count, 2047) -----------------------------
#define QQ_ORDER_BINS 2048
__global__ void k_msm_order_hist(const unsigned int* __restrict__ counts, size_t total, unsigned int* __restrict__ hist) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += stride) {
        unsigned int c = counts[b];
        unsigned int key = QQ_ORDER_BINS - 1 - (c < QQ_ORDER_BINS - 1 ? c : QQ_ORDER_BINS - 1);
        atomicAdd(&hist[key], 1u);
    }
}
__global__ void k_msm_order_scatter(const unsigned int* __restrict__ counts, size_t total,
                                    const unsigned int* __restrict__ hist_off, unsigned int* __restrict__ cursor,
                                    unsigned int* __restrict__ order) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += stride) {
        unsigned int c = counts[b];
        unsigned int key = QQ_ORDER_BINS - 1 - (c < QQ_ORDER_BINS - 1 ? c : QQ_ORDER_BINS - 1);
        unsigned int slot = atomicAdd(&cursor[key], 1u);
        order[hist_off[key] + slot] = (unsigned int)b;
    }
}

// ---- bucket accumulation: thread t sums the points of bucket order[t] ---------------------------------------------
__global__ void __launch_bounds__(128) k_msm_accumulate(const u32x4* __restrict__ niels,
                                                        const unsigned int* __restrict__ sorted,
                                                        const unsigned int* __restrict__ offsets,
                                                        const unsigned int* __restrict__ counts,
                                                        const unsigned int* __restrict__ order, size_t total,
                                                        u32x4* __restrict__ buckets) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    unsigned int b = order[t];
    unsigned int cnt = counts[b];
    const unsigned int* ent = sorted + offsets[b];
    ge_p3 acc;
    ge_identity(acc);
    for (unsigned int e = 0; e < cnt; e++) {
        unsigned int v = __ldg(ent + e);
        ge_niels nl;
        niels_load_padded(nl, niels + (size_t)QQ_NIELS_STRIDE_Q * (v & 0x7fffffffu));
        ge_niels_cneg(nl, v >> 31);
        ge_madd(acc, acc, nl);
    }
    ge_p3_store(buckets + QQ_PT_Q * (size_t)b, acc);
}

// ---- bucket reduction ------------------------------------------------------------------------------------------------
// thread (k, s): segment of SEG buckets starting at base = s * SEG of window k:
//   run = sum B_j, W = sum (j - base + 1) B_j  (running-sum trick, high to low), result = W + base * run
__global__ void __launch_bounds__(128) k_msm_reduce_seg(const u32x4* __restrict__ buckets, msm_geom g, int SEG,
                                                        u32x4* __restrict__ seg_out) {
    int nseg = g.NB / SEG;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.K * nseg) return;
    int k = t / nseg, s = t - k * nseg;
    int base = s * SEG;
    const u32x4* bk = buckets + QQ_PT_Q * ((size_t)k * g.NB + base);
    ge_p3 run, sum;
    ge_identity(run);
    ge_identity(sum);
    for (int j = SEG - 1; j >= 0; j--) {
        ge_p3 p;
        ge_p3_load(p, bk + QQ_PT_Q * j);
        ge_cached c;
        ge_to_cached(c, p);
        ge_add(run, run, c);
        ge_to_cached(c, run);
        ge_add(sum, sum, c);
    }
    if (base != 0) {
        // sum += base * run  (double-and-add over the bits of base, at most 15 bits)
        ge_cached cr;
        ge_to_cached(cr, run);
        ge_p3 m;
        ge_identity(m);
        for (int bit = 15; bit >= 0; bit--) {
            ge_dbl<true>(m, m);
            if ((base >> bit) & 1) ge_add(m, m, cr);
        }
        ge_cached cm;
        ge_to_cached(cm, m);
        ge_add(sum, sum, cm);
    }
    ge_p3_store(seg_out + QQ_PT_Q * (size_t)t, sum);
}
// block r sums in[r * row_len .. (r + 1) * row_len) -> out[r]
__global__ void __launch_bounds__(128) k_point_sum_rows(const u32x4* __restrict__ in, int row_len,
                                                        u32x4* __restrict__ out) {
    __shared__ u32x4 sm[128 * QQ_PT_Q];
    ge_p3 acc;
    ge_identity(acc);
    const u32x4* row = in + QQ_PT_Q * (size_t)blockIdx.x * row_len;
    for (int t = threadIdx.x; t < row_len; t += blockDim.x) {
        ge_p3 p;
        ge_p3_load(p, row + QQ_PT_Q * t);
        ge_cached c;
        ge_to_cached(c, p);
        ge_add(acc, acc, c);
    }
    ge_p3_store(sm + QQ_PT_Q * threadIdx.x, acc);
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            ge_p3 p;
            ge_p3_load(p, sm + QQ_PT_Q * (threadIdx.x + s));
            ge_cached c;
            ge_to_cached(c, p);
            ge_add(acc, acc, c);
            ge_p3_store(sm + QQ_PT_Q * threadIdx.x, acc);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) ge_p3_store(out + QQ_PT_Q * blockIdx.x, acc);
}
// result = sum_k 2^(c k) win[k].  A chain of c (K - 1) = 240 dependent doublings: latency, not throughput.  Four lanes of one
// warp cooperate on every doubling -- lane l squares one of (X, Y, Z, X + Y), the four squares are exchanged through
// shared memory, every lane forms the completed-point terms and multiplies out one output coordinate -- so a doubling
// costs one squaring + one multiplication of latency instead of four + four.
__global__ void __launch_bounds__(32) k_msm_horner(const u32x4* __restrict__ win, msm_geom g, u32x4* __restrict__ result) {
    __shared__ fe sq_s[4];
    __shared__ fe co_s[4];
    const int lane = threadIdx.x;
    if (blockIdx.x != 0 || lane >= 4) return;
    const unsigned mask = 0xfu;
    ge_p3 acc;
    ge_p3_load(acc, win + QQ_PT_Q * (size_t)(g.K - 1));
    // lane l keeps coordinate l of the running point in `mine` (0: X, 1: Y, 2: Z, 3: T)
    fe mine = lane == 0 ? acc.X : (lane == 1 ? acc.Y : (lane == 2 ? acc.Z : acc.T));
    for (int k = g.K - 2; k >= 0; k--) {
        for (int i = 0; i < g.c; i++) {
            co_s[lane] = mine;
            __syncwarp(mask);
            fe in;
            if (lane == 3) fe_add(in, co_s[0], co_s[1]);   // X + Y
            else in = mine;
            fe sq;
            fe_sq(sq, in);
            sq_s[lane] = sq;
            __syncwarp(mask);
            fe xx = sq_s[0], yy = sq_s[1], zz = sq_s[2], s = sq_s[3];
            fe cx, cy, cz, ct, t;
            fe_add(cy, yy, xx);
            fe_sub(cz, yy, xx);
            fe_sub(cx, s, cy);
            fe_add(t, zz, zz);
            fe_sub(ct, t, cz);
            // X3 = cx ct, Y3 = cy cz, Z3 = cz ct, T3 = cx cy
            fe a = (lane == 0 || lane == 3) ? cx : (lane == 1 ? cy : cz);
            fe b = (lane == 0 || lane == 2) ? ct : (lane == 1 ? cz : cy);
            fe_mul(mine, a, b);
            __syncwarp(mask);
        }
        // add window k: lane 0 gathers the point, performs the addition, and redistributes
        co_s[lane] = mine;
        __syncwarp(mask);
        if (lane == 0) {
            ge_p3 r, p;
            r.X = co_s[0]; r.Y = co_s[1]; r.Z = co_s[2]; r.T = co_s[3];
            ge_p3_load(p, win + QQ_PT_Q * (size_t)k);
            ge_cached c;
            ge_to_cached(c, p);
            ge_add(r, r, c);
            co_s[0] = r.X; co_s[1] = r.Y; co_s[2] = r.Z; co_s[3] = r.T;
        }
        __syncwarp(mask);
        mine = co_s[lane];
        __syncwarp(mask);
    }
    co_s[lane] = mine;
    __syncwarp(mask);
    if (lane == 0) {
        ge_p3 r;
        r.X = co_s[0]; r.Y = co_s[1]; r.Z = co_s[2]; r.T = co_s[3];
        ge_p3_store(result, r);
    }
}

}  // namespace qq
