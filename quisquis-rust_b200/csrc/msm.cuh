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
//   k_msm_vcount/vfill buckets above `cap` entries are cut into virtual buckets (bounded work per thread)
//   k_msm_order_*     virtual buckets ordered by population (descending) so the lanes of a warp get equal work
//   k_msm_accumulate  one thread per virtual bucket: mixed additions (7 M), next gather issued before each addition
//   k_msm_reduce0/_level  recursive running-sum reduction, 8 elements per thread and level, no scalar multiplications
//   k_msm_sum_levels  tree sums of the per-level W arrays;  k_msm_window_totals  Horner over the levels
//   k_msm_horner      sum_k 2^(c k) T_k
#pragma once
#include "kernels.cuh"
#include "ge_coop.cuh"
#include "ge_warp.cuh"

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

// PRE = false: decompress the compressed points into `niels` (and ok into pre_ok-less status).
// PRE = true : the points were decompressed earlier by k_msm_points_prepare (generator sets reused across MSMs);
//              only the scalars are recoded.  `pre_ok[i]` is the validity recorded then.
template <bool PRE>
__global__ void __launch_bounds__(256, 2) k_msm_prepare(const u32x4* __restrict__ points, const u32x4* __restrict__ scalars,
                                                     size_t n, size_t n_dec, msm_geom g, u32x4* __restrict__ niels,
                                                     const uint8_t* __restrict__ pre_ok,
                                                     uint8_t* __restrict__ term_status, int16_t* __restrict__ digits,
                                                     unsigned int* __restrict__ slots, unsigned int* __restrict__ counts,
                                                     unsigned int win_stride, const unsigned int* __restrict__ group_of) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        // grouped form (several independent MSMs in one pass): group g owns the windows [g K, (g + 1) K) of the bucket array
        const size_t kb = group_of != nullptr ? (size_t)group_of[i] * (size_t)g.K : 0;
        // Scalar side first: the K histogram atomics (which also hand out each entry's position inside its bucket, so
        // the scatter needs no atomics) are issued before the decompression and collected after it -- their round
        // trips hide behind the square-root chain.  Digits do not depend on the point's validity: a bad term fails
        // the whole MSM (status), whatever its bucket.
        u32 s[8];
        load_words32(s, scalars, i);
        u32 canon = sc_is_canonical(s);
        u32 r[9];
        sc_recode_bias_rt(r, s, g.c, g.K);
        unsigned int slot[16];
        const bool fast = g.K <= 16;
        if (fast) {
#pragma unroll
            for (int k = 0; k < 16; k++) {
                slot[k] = 0;
                if (k < g.K) {
                    int d = canon ? sc_digit_rt(r, g.c, k) : 0;
                    digits[(size_t)k * n + i] = (int16_t)d;
                    if (d != 0) slot[k] = atomicAdd(&counts[(kb + k) * win_stride + (size_t)((d < 0 ? -d : d) - 1)], 1u);
                }
            }
        } else {
            for (int k = 0; k < g.K; k++) {
                int d = canon ? sc_digit_rt(r, g.c, k) : 0;
                digits[(size_t)k * n + i] = (int16_t)d;
                if (d != 0) slots[(size_t)k * n + i] = atomicAdd(&counts[(kb + k) * win_stride + (size_t)((d < 0 ? -d : d) - 1)], 1u);
            }
        }
        u32 ok;
        if (PRE) {
            ok = pre_ok[i];
        } else if (i >= n_dec) {
            ok = 1;       // decompressed by k_msm_decompress, which also downgrades the status
        } else {
            u32 w[8];
            load_words32(w, points, i);
            ge_p3 p;
            ok = ristretto_decompress(p, w);
            ge_niels nl;
            ge_to_niels_z1(nl, p);
            niels_store_padded(niels + (size_t)QQ_NIELS_STRIDE_Q * i, nl);
        }
        term_status[i] = canon ? (ok ? 0 : 1) : 2;
        if (fast) {
#pragma unroll
            for (int k = 0; k < 16; k++)
                if (k < g.K) slots[(size_t)k * n + i] = slot[k];
        }
    }
}
// The point side alone (decompression to affine Niels) for the terms k_msm_prepare left out (n_dec <= i < n): it runs while
// the counting sort of the digits proceeds on a second, high-priority stream -- the sort is latency / atomics / DRAM
// bound, the decompression multiply bound.  (Splitting ALL the decompression off was measured slower: the histogram
// atomics of a scalar-only kernel are no longer hidden behind the square-root chain.)
__global__ void __launch_bounds__(256, 2) k_msm_decompress(const u32x4* __restrict__ points, size_t n, u32x4* __restrict__ niels,
                                                           uint8_t* __restrict__ term_status) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u32 w[8];
        load_words32(w, points, i);
        ge_p3 p;
        u32 ok = ristretto_decompress(p, w);
        ge_niels nl;
        ge_to_niels_z1(nl, p);
        niels_store_padded(niels + (size_t)QQ_NIELS_STRIDE_Q * i, nl);
        if (!ok && term_status[i] == 0) term_status[i] = 1;
    }
}
// compressed points -> the MSM's point form (affine Niels, 96 B) + validity, once, for reuse
__global__ void __launch_bounds__(256) k_msm_points_prepare(const u32x4* __restrict__ points, size_t n,
                                                            u32x4* __restrict__ niels, uint8_t* __restrict__ ok_out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        u32 w[8];
        load_words32(w, points, i);
        ge_p3 p;
        u32 ok = ristretto_decompress(p, w);
        ge_niels nl;
        ge_to_niels_z1(nl, p);
        niels_store_padded(niels + (size_t)QQ_NIELS_STRIDE_Q * i, nl);
        ok_out[i] = (uint8_t)ok;
    }
}

// the same for slots that were rewritten after an early k_msm_points_prepare: slot i belongs to owner i / cap; owners whose flag
// says "unchanged" (flag == 0 when keep_is_zero, flag != 0 otherwise) are left alone (batched verifiers: a proof that left
// the aggregate got its points replaced)
__global__ void __launch_bounds__(256) k_msm_points_reprepare(const u32x4* __restrict__ points, size_t n, unsigned int cap,
                                                              const uint8_t* __restrict__ flag, int keep_is_zero,
                                                              u32x4* __restrict__ niels, uint8_t* __restrict__ ok_out) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const bool unchanged = keep_is_zero ? flag[i / cap] == 0 : flag[i / cap] != 0;
        if (unchanged) continue;
        u32 w[8];
        load_words32(w, points, i);
        ge_p3 p;
        u32 ok = ristretto_decompress(p, w);
        ge_niels nl;
        ge_to_niels_z1(nl, p);
        niels_store_padded(niels + (size_t)QQ_NIELS_STRIDE_Q * i, nl);
        ok_out[i] = (uint8_t)ok;
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

// win_stride = NB, idx_stride = 0: one bucket set per window, entries name the point.  win_stride = 0, idx_stride = N (shifted
// point sets, see k_msm_shift_build): ONE bucket set for all windows, the entry of (window k, point i) names 2^(c k) P_i.
__global__ void __launch_bounds__(256) k_msm_scatter(const int16_t* __restrict__ digits, size_t n, msm_geom g,
                                                     const unsigned int* __restrict__ offsets,
                                                     const unsigned int* __restrict__ slots, unsigned int* __restrict__ sorted,
                                                     unsigned int win_stride, unsigned int idx_stride,
                                                     const unsigned int* __restrict__ group_of) {
    size_t total = (size_t)g.K * n;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        int d = digits[t];
        if (d == 0) continue;
        size_t k = t / n;
        size_t i = t - k * n;
        size_t kb = group_of != nullptr ? (size_t)group_of[i] * (size_t)g.K : 0;
        size_t key = (kb + k) * win_stride + (size_t)((d < 0 ? -d : d) - 1);
        sorted[offsets[key] + slots[t]] = (unsigned int)(i + k * idx_stride) | (d < 0 ? 0x80000000u : 0u);
    }
}
// ---- bucket ordering by population (descending), counting sort on min(count, 2047) -----------------------------
// Populations cluster on a few dozen values, so the histogram and the slot reservation are privatised per block in
// shared memory (one global atomic per (block, occupied bin) instead of one per bucket).
#define QQ_ORDER_BINS 2048
#define QQ_COMBINE_HEAVY 64   // parts of one bucket beyond which a whole block sums it (k_msm_combine)
#define QQ_ORDER_TILE 4096   // elements per block iteration (256 threads x 16)
__device__ __forceinline__ unsigned int order_key(unsigned int c) {
    return QQ_ORDER_BINS - 1 - (c < QQ_ORDER_BINS - 1 ? c : QQ_ORDER_BINS - 1);
}
__global__ void __launch_bounds__(256) k_msm_order_hist(const unsigned int* __restrict__ counts, size_t total,
                                                        unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[QQ_ORDER_BINS];
    for (int i = threadIdx.x; i < QQ_ORDER_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += stride) atomicAdd(&sh[order_key(counts[b])], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < QQ_ORDER_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
__global__ void __launch_bounds__(256) k_msm_order_scatter(const unsigned int* __restrict__ counts, size_t total,
                                                           const unsigned int* __restrict__ hist_off,
                                                           unsigned int* __restrict__ cursor, unsigned int* __restrict__ order) {
    __shared__ unsigned int sh_cnt[QQ_ORDER_BINS];   // tile histogram, then running local rank
    __shared__ unsigned int sh_base[QQ_ORDER_BINS];  // global slot of the tile's first element of each bin
    for (size_t tile = (size_t)blockIdx.x * QQ_ORDER_TILE; tile < total; tile += (size_t)gridDim.x * QQ_ORDER_TILE) {
        for (int i = threadIdx.x; i < QQ_ORDER_BINS; i += blockDim.x) sh_cnt[i] = 0;
        __syncthreads();
        size_t end = tile + QQ_ORDER_TILE < total ? tile + QQ_ORDER_TILE : total;
        for (size_t b = tile + threadIdx.x; b < end; b += blockDim.x) atomicAdd(&sh_cnt[order_key(counts[b])], 1u);
        __syncthreads();
        for (int i = threadIdx.x; i < QQ_ORDER_BINS; i += blockDim.x) {
            unsigned int c = sh_cnt[i];
            if (c) sh_base[i] = hist_off[i] + atomicAdd(&cursor[i], c);
            sh_cnt[i] = 0;
        }
        __syncthreads();
        for (size_t b = tile + threadIdx.x; b < end; b += blockDim.x) {
            unsigned int key = order_key(counts[b]);
            unsigned int r = atomicAdd(&sh_cnt[key], 1u);
            order[sh_base[key] + r] = (unsigned int)b;
        }
        __syncthreads();
    }
}

// ---- pipelined tail: the windows are cut into `ng` ranks (rank 0 = the top windows) whose buckets are accumulated by separate
// launches, so that the depth-bound reduction / Horner chain of one rank runs (on the high-priority stream) under the
// accumulation of the next.  Virtual buckets are laid out in bucket order, so a rank is a contiguous range of them:
// vb[0] = virtual buckets in use, vb[r] (1 <= r < ng) = end of rank r's range = first virtual bucket of window kw[r].
struct msm_ranks {
    int ng;
    unsigned int W;          // ordering bins per rank (ng * W == QQ_ORDER_BINS)
    int kw[5];               // rank r covers windows [kw[r + 1], kw[r]); kw[0] = K, kw[ng] = 0
};
__global__ void k_msm_rank_bounds(const unsigned int* __restrict__ voff, const unsigned int* __restrict__ nsub, size_t total, int NB,
                                  msm_ranks rk, unsigned int* __restrict__ vb) {
    int r = threadIdx.x;
    if (r == 0) vb[0] = voff[total - 1] + nsub[total - 1];
    else if (r < rk.ng) vb[r] = voff[(size_t)rk.kw[r] * NB];
}
__device__ __forceinline__ unsigned int order_key_g(unsigned int c, unsigned int v, const unsigned int* vb, const msm_ranks& rk) {
    unsigned int r = 0;
    for (int i = 1; i < rk.ng; i++) r = v < vb[i] ? (unsigned int)i : r;
    return r * rk.W + (rk.W - 1 - (c < rk.W - 1 ? c : rk.W - 1));
}
__global__ void __launch_bounds__(256) k_msm_order_hist_g(const unsigned int* __restrict__ counts, const unsigned int* __restrict__ vb,
                                                          msm_ranks rk, unsigned int* __restrict__ hist) {
    __shared__ unsigned int sh[QQ_ORDER_BINS];
    for (int i = threadIdx.x; i < QQ_ORDER_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const size_t total = vb[0];
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += stride)
        atomicAdd(&sh[order_key_g(counts[b], (unsigned int)b, vb, rk)], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < QQ_ORDER_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
__global__ void __launch_bounds__(256) k_msm_order_scatter_g(const unsigned int* __restrict__ counts, const unsigned int* __restrict__ vb,
                                                             msm_ranks rk, const unsigned int* __restrict__ hist_off,
                                                             unsigned int* __restrict__ cursor, unsigned int* __restrict__ order) {
    __shared__ unsigned int sh_cnt[QQ_ORDER_BINS];
    __shared__ unsigned int sh_base[QQ_ORDER_BINS];
    const size_t total = vb[0];
    for (size_t tile = (size_t)blockIdx.x * QQ_ORDER_TILE; tile < total; tile += (size_t)gridDim.x * QQ_ORDER_TILE) {
        for (int i = threadIdx.x; i < QQ_ORDER_BINS; i += blockDim.x) sh_cnt[i] = 0;
        __syncthreads();
        size_t end = tile + QQ_ORDER_TILE < total ? tile + QQ_ORDER_TILE : total;
        for (size_t b = tile + threadIdx.x; b < end; b += blockDim.x) atomicAdd(&sh_cnt[order_key_g(counts[b], (unsigned int)b, vb, rk)], 1u);
        __syncthreads();
        for (int i = threadIdx.x; i < QQ_ORDER_BINS; i += blockDim.x) {
            unsigned int c = sh_cnt[i];
            if (c) sh_base[i] = hist_off[i] + atomicAdd(&cursor[i], c);
            sh_cnt[i] = 0;
        }
        __syncthreads();
        for (size_t b = tile + threadIdx.x; b < end; b += blockDim.x) {
            unsigned int key = order_key_g(counts[b], (unsigned int)b, vb, rk);
            unsigned int r = atomicAdd(&sh_cnt[key], 1u);
            order[sh_base[key] + r] = (unsigned int)b;
        }
        __syncthreads();
    }
}

// ---- virtual buckets ---------------------------------------------------------------------------------------------------
// A bucket with more than `cap` entries is cut into ceil(cnt / cap) virtual buckets, each summed by its own thread, so
// that skewed digit distributions (the short top window: scalars < 2^253 leave it only 2^(253 - c (K-1)) distinct
// digits; small balances; repeated scalars) do not serialise on one thread.  nsub[b] = number of parts of bucket b.
__global__ void k_msm_vcount(const unsigned int* __restrict__ counts, size_t total, unsigned int cap,
                             unsigned int* __restrict__ nsub) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += stride)
        nsub[b] = (counts[b] + cap - 1) / cap;
}
__global__ void k_msm_vfill(const unsigned int* __restrict__ counts, const unsigned int* __restrict__ offsets,
                            const unsigned int* __restrict__ voff, size_t total, unsigned int cap,
                            unsigned int* __restrict__ vstart, unsigned int* __restrict__ vcnt,
                            unsigned int* __restrict__ multi, unsigned int* __restrict__ nmulti) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += stride) {
        unsigned int c = counts[b], o = offsets[b], v = voff[b];
        if (c > cap) multi[atomicAdd(nmulti, 1u)] = (unsigned int)b;
        if (c > cap * QQ_COMBINE_HEAVY) nmulti[1] = 1u;      // some bucket has more than QQ_COMBINE_HEAVY parts (k_msm_combine's second pass)
        for (unsigned int done = 0; done < c; done += cap, v++) {
            vstart[v] = o + done;
            vcnt[v] = c - done < cap ? c - done : cap;
        }
    }
}

// affine Niels (Z = 1) -> extended with Z = 4: X = 2 (ypx - ymx) = 4x, Y = 2 (ypx + ymx) = 4y, T = (ypx - ymx)(ypx + ymx) = 4xy
__device__ __forceinline__ void ge_from_niels(ge_p3& r, const ge_niels& n) {
    fe dx, sy;
    fe_sub(dx, n.ypx, n.ymx);
    fe_add(sy, n.ypx, n.ymx);
    fe_add(r.X, dx, dx);
    fe_add(r.Y, sy, sy);
    fe_0(r.Z);
    r.Z.v[0] = 4;
    fe_mul(r.T, dx, sy);
}

// Point sets that are reused (qq_msm_points_prepare): window k of the shifted form holds 2^(c k) P_i, so that every digit of
// every scalar lands in ONE set of 2^(c-1) buckets and the per-window reductions and the Horner chain of 240 doublings
// disappear from the MSM.  One step: ext_i = 2^c * prev_i (affine Niels in, extended out + Z for the batch inversion).
__global__ void __launch_bounds__(256) k_msm_shift_build(const u32x4* __restrict__ prev, size_t n, int c, u32x4* __restrict__ ext,
                                                         u32x4* __restrict__ zs) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        ge_niels nl;
        niels_load_padded(nl, prev + (size_t)QQ_NIELS_STRIDE_Q * i);
        ge_p3 p;
        ge_from_niels(p, nl);
        for (int k = 0; k < c; k++) ge_dbl<true>(p, p);
        ge_p3_store(ext + (size_t)QQ_PT_Q * i, p);
        u32x4 q;
        q.x = p.Z.v[0]; q.y = p.Z.v[1]; q.z = p.Z.v[2]; q.w = p.Z.v[3]; zs[2 * i] = q;
        q.x = p.Z.v[4]; q.y = p.Z.v[5]; q.z = p.Z.v[6]; q.w = p.Z.v[7]; zs[2 * i + 1] = q;
    }
}

// ---- bucket accumulation: thread t sums the entries of virtual bucket order[t] ----------------------------------------
// The gather of entry e + 1 (index, then 96 B Niels point) is issued before the mixed addition of entry e.
__device__ __forceinline__ void msm_accumulate_one(const u32x4* __restrict__ niels, const unsigned int* __restrict__ sorted,
                                                   const unsigned int* __restrict__ vstart, const unsigned int* __restrict__ vcnt,
                                                   unsigned int v, u32x4* __restrict__ partial) {
    unsigned int cnt = vcnt[v];
    if (cnt == 0) return;
    const unsigned int* ent = sorted + vstart[v];
    unsigned int cur = __ldg(ent);
    ge_niels nl, nx;
    niels_load_padded(nl, niels + (size_t)QQ_NIELS_STRIDE_Q * (cur & 0x7fffffffu));
    ge_niels_cneg(nl, cur >> 31);
    ge_p3 acc;
    ge_from_niels(acc, nl);
    if (cnt > 1) {
        cur = __ldg(ent + 1);
        niels_load_padded(nx, niels + (size_t)QQ_NIELS_STRIDE_Q * (cur & 0x7fffffffu));
    }
    for (unsigned int e = 1; e < cnt; e++) {
        nl = nx;
        unsigned int sign = cur >> 31;
        if (e + 1 < cnt) {
            cur = __ldg(ent + e + 1);
            niels_load_padded(nx, niels + (size_t)QQ_NIELS_STRIDE_Q * (cur & 0x7fffffffu));
            if (e + 2 < cnt) {   // pull the point after that towards L2
                const u32x4* pf = niels + (size_t)QQ_NIELS_STRIDE_Q * (__ldg(ent + e + 2) & 0x7fffffffu);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + 5));
            }
        }
        ge_niels_cneg(nl, sign);
        ge_madd(acc, acc, nl);
    }
    ge_p3_store(partial + QQ_PT_Q * (size_t)v, acc);
}
__global__ void __launch_bounds__(128) k_msm_accumulate(const u32x4* __restrict__ niels,
                                                        const unsigned int* __restrict__ sorted,
                                                        const unsigned int* __restrict__ vstart,
                                                        const unsigned int* __restrict__ vcnt,
                                                        const unsigned int* __restrict__ order, size_t vmax,
                                                        u32x4* __restrict__ partial) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= vmax) return;
    msm_accumulate_one(niels, sorted, vstart, vcnt, order[t], partial);
}
// one rank of the pipelined form: entries [*lo, *hi) of `order` (bounds known on the device only; hi == nullptr: *lo + count)
__global__ void __launch_bounds__(128) k_msm_accumulate_rank(const u32x4* __restrict__ niels,
                                                             const unsigned int* __restrict__ sorted,
                                                             const unsigned int* __restrict__ vstart,
                                                             const unsigned int* __restrict__ vcnt,
                                                             const unsigned int* __restrict__ order,
                                                             const unsigned int* __restrict__ lo, const unsigned int* __restrict__ hi,
                                                             u32x4* __restrict__ partial) {
    size_t t = (size_t)*lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)*hi) return;
    msm_accumulate_one(niels, sorted, vstart, vcnt, order[t], partial);
}

// Buckets that were cut into several virtual buckets: one warp per such bucket sums its partial sums (lanes stride over
// the parts, then a shared-memory tree) into the first part, so that the reduction reads exactly one point per bucket.
// `multi` lists those buckets (appended by k_msm_vfill), *nmulti is their number.
__global__ void __launch_bounds__(128) k_msm_combine(u32x4* __restrict__ partial, const unsigned int* __restrict__ voff,
                                                     const unsigned int* __restrict__ nsub,
                                                     const unsigned int* __restrict__ multi,
                                                     const unsigned int* __restrict__ nmulti, unsigned int b_lo, unsigned int b_hi) {
    __shared__ u32x4 sm[128 * QQ_PT_Q];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    u32x4* my = sm + QQ_PT_Q * (size_t)(wib * 32);
    unsigned int cnt = *nmulti;
    for (unsigned int w = blockIdx.x * 4 + wib; w < cnt; w += gridDim.x * 4) {
        unsigned int b = multi[w];
        if (b < b_lo || b >= b_hi) continue;      // warp-uniform: another rank's bucket
        unsigned int v0 = voff[b], np = nsub[b];
        if (np > QQ_COMBINE_HEAVY) continue;      // the second pass below
        ge_p3 acc;
        ge_identity(acc);
        for (unsigned int q = lane; q < np; q += 32) {
            ge_p3 p;
            ge_p3_load(p, partial + QQ_PT_Q * (size_t)(v0 + q));
            ge_cached c;
            ge_to_cached(c, p);
            ge_add(acc, acc, c);
        }
        ge_p3_store(my + QQ_PT_Q * lane, acc);
        __syncwarp();
        for (int s = 16; s > 0; s >>= 1) {
            if (lane < s && (unsigned int)(lane + s) < np) {
                ge_p3 p;
                ge_p3_load(p, my + QQ_PT_Q * (lane + s));
                ge_cached c;
                ge_to_cached(c, p);
                ge_add(acc, acc, c);
                ge_p3_store(my + QQ_PT_Q * lane, acc);
            }
            __syncwarp();
        }
        if (lane == 0) ge_p3_store(partial + QQ_PT_Q * (size_t)v0, acc);
        __syncwarp();
    }
    // Heavy buckets (more than QQ_COMBINE_HEAVY parts: the digit that the top bit or the recoding carry of a class of scalars
    // lands in - 128-bit weights put half of their terms into ONE bucket of window 8): the whole block sums one, np / 128 + 7
    // dependent additions instead of np / 32 + 5 on a warp (2 400 parts: 175 -> 60 us).
    if (nmulti[1] == 0) return;      // no heavy bucket in this MSM (block-uniform)
    __syncthreads();
    for (unsigned int w = blockIdx.x; w < cnt; w += gridDim.x) {
        unsigned int b = multi[w];
        if (b < b_lo || b >= b_hi) continue;      // block-uniform
        unsigned int v0 = voff[b], np = nsub[b];
        if (np <= QQ_COMBINE_HEAVY) continue;
        ge_p3 acc;
        ge_identity(acc);
        for (unsigned int q = threadIdx.x; q < np; q += blockDim.x) {
            ge_p3 p;
            ge_p3_load(p, partial + QQ_PT_Q * (size_t)(v0 + q));
            ge_cached c;
            ge_to_cached(c, p);
            ge_add(acc, acc, c);
        }
        ge_p3_store(sm + QQ_PT_Q * threadIdx.x, acc);
        __syncthreads();
        for (int s2 = 64; s2 > 0; s2 >>= 1) {
            if ((int)threadIdx.x < s2) {
                ge_p3 p;
                ge_p3_load(p, sm + QQ_PT_Q * (threadIdx.x + s2));
                ge_cached c;
                ge_to_cached(c, p);
                ge_add(acc, acc, c);
                ge_p3_store(sm + QQ_PT_Q * threadIdx.x, acc);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) ge_p3_store(partial + QQ_PT_Q * (size_t)v0, acc);
        __syncthreads();
    }
}

// ---- bucket reduction ------------------------------------------------------------------------------------------------
// Window value T = sum_{j=1..NB} j B_j.  With A_i = B_{i+1}:  T = P0(A) + sum A,  P0(A) = sum_i i A_i, and for segments
// of S elements  P0(A) = sum_s W_s + S * P0(run),  W_s = sum_{i<S} i A_{sS+i},  run_s = sum_{i<S} A_{sS+i}
// (running-sum trick, 2S - 1 additions per segment).  Applied recursively the element count shrinks by S per level:
//     T = sum_l S^l SW_l + R,   SW_l = sum_s W^l_s,   R = the single run left at the top.
// No scalar multiplications by segment bases, and every level is data-parallel over (window, segment).
#define QQ_MSM_RSEG 8
// k_msm_reduce0/_level, k_msm_window_totals and k_msm_horner are four-lane cooperative (ge_coop.cuh): group = 4 adjacent lanes.
// One running-sum step: run += A_i; W += run   (skipped for i = 0).
template <bool INL>
__device__ __forceinline__ void coop_runsum_step(fe& run, fe& w, const fe& a, bool add_w, int r) {
    fe ca = coop_to_cached<INL>(a, r);
    run = coop_add<INL>(run, ca, r);
    if (add_w) {
        fe cr = coop_to_cached<INL>(run, r);
        w = coop_add<INL>(w, cr, r);
    }
}
// level 0: element (k, j) = bucket k * NB + j; after k_msm_combine its sum is the first of its partial sums
__global__ void __launch_bounds__(128) k_msm_reduce0(const u32x4* __restrict__ partial, const unsigned int* __restrict__ voff,
                                                     const unsigned int* __restrict__ nsub, msm_geom g, int mout,
                                                     u32x4* __restrict__ run_out, u32x4* __restrict__ w_out, int k0) {
    int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, r = threadIdx.x & 3;
    bool act = gid < g.K * mout;
    int t = act ? gid : 0;
    int k = t / mout, s = t - k * mout;
    fe run = coop_identity(r), w = coop_identity(r);
#pragma unroll 1
    for (int i = QQ_MSM_RSEG - 1; i >= 0; i--) {
        int j = s * QQ_MSM_RSEG + i;
        fe a = coop_identity(r);
        if (j < g.NB) {
            size_t b = (size_t)(k0 + k) * g.NB + j;       // g.K = windows of this launch, k0 = its first window
            if (nsub[b] != 0) a = coop_load(partial + QQ_PT_Q * (size_t)voff[b], r);
        }
        coop_runsum_step<true>(run, w, a, i > 0, r);
    }
    if (act) {
        coop_store(run_out + QQ_PT_Q * (size_t)t, r, run);
        coop_store(w_out + QQ_PT_Q * (size_t)t, r, w);
    }
}
// level >= 1: in[k][0..m) -> run_out[k][0..mout), w_out[k][0..mout),  mout = ceil(m / S).  Four lanes per segment,
// out-of-line field products and a rolled loop: these launches are pure latency (a few thousand threads), and the
// unrolled, inlined form spends its time fetching instructions (64-77 us per level against 30 us here; one thread per
// segment: 55 us).
__global__ void __launch_bounds__(128) k_msm_reduce_level(const u32x4* __restrict__ in, int K, int m, int mout,
                                                               u32x4* __restrict__ run_out, u32x4* __restrict__ w_out) {
    int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, r = threadIdx.x & 3;
    bool act = gid < K * mout;
    int t = act ? gid : 0;
    int k = t / mout, s = t - k * mout;
    fe run = coop_identity(r), w = coop_identity(r);
#pragma unroll 1
    for (int i = QQ_MSM_RSEG - 1; i >= 0; i--) {
        int j = s * QQ_MSM_RSEG + i;
        fe a = coop_identity(r);
        if (j < m) a = coop_load(in + QQ_PT_Q * ((size_t)k * m + j), r);
        coop_runsum_step<false>(run, w, a, i > 0, r);
    }
    if (act) {
        coop_store(run_out + QQ_PT_Q * (size_t)t, r, run);
        coop_store(w_out + QQ_PT_Q * (size_t)t, r, w);
    }
}
// SW_l for every (level, window): block (l, k) sums row k of level l's W array (w_all + off[l], rows of len[l] points)
#define QQ_MSM_MAXLEVELS 8
struct msm_levels {
    int L;
    int len[QQ_MSM_MAXLEVELS];
    unsigned int off[QQ_MSM_MAXLEVELS];
};
__global__ void __launch_bounds__(512) k_msm_sum_levels(const u32x4* __restrict__ w_all, msm_levels lv, int K, int l_first,
                                                        u32x4* __restrict__ sw, int Kg, int k0) {
    extern __shared__ __align__(16) u32x4 sm[];   // blockDim.x points
    int l = l_first + blockIdx.x / Kg, k = k0 + blockIdx.x % Kg;      // windows [k0, k0 + Kg) of K
    int row_len = lv.len[l];
    const u32x4* row = w_all + QQ_PT_Q * ((size_t)lv.off[l] + (size_t)k * row_len);
    ge_p3 acc;
    ge_identity(acc);
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
        if (threadIdx.x < s && threadIdx.x + s < row_len) {
            ge_p3 p;
            ge_p3_load(p, sm + QQ_PT_Q * (threadIdx.x + s));
            ge_cached c;
            ge_to_cached(c, p);
            ge_add(acc, acc, c);
            ge_p3_store(sm + QQ_PT_Q * threadIdx.x, acc);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) ge_p3_store(sw + QQ_PT_Q * ((size_t)l * K + k), acc);
}
// wins[k] = R_k + sum_l S^l SW_l[k]   (Horner over the levels: log2(S) doublings + one addition per level)
__global__ void __launch_bounds__(128) k_msm_window_totals(const u32x4* __restrict__ sw, const u32x4* __restrict__ run_top,
                                                           msm_levels lv, int K, u32x4* __restrict__ wins, int Kg, int k0) {
    int gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 2, r = threadIdx.x & 3;
    bool act = gid < Kg;
    int k = k0 + (act ? gid : 0);
    fe t = coop_load(sw + QQ_PT_Q * ((size_t)(lv.L - 1) * K + k), r);
    for (int l = lv.L - 2; l >= 0; l--) {
        t = coop_dbl(t, r);
        t = coop_dbl(t, r);
        t = coop_dbl(t, r);
        fe p = coop_load(sw + QQ_PT_Q * ((size_t)l * K + k), r);
        t = coop_add(t, coop_to_cached(p, r), r);
    }
    fe p = coop_load(run_top + QQ_PT_Q * k, r);
    t = coop_add(t, coop_to_cached(p, r), r);
    if (act) coop_store(wins + QQ_PT_Q * k, r, t);
}
// result = sum_k 2^(c k) win[k]: a chain of c (K - 1) dependent doublings (240 for c = 16), each one squaring + one
// multiplication of latency.  One warp per MSM (block b: the windows [b K, (b + 1) K) of the grouped form); every group of four
// lanes computes the same chain, the first one stores it.
__global__ void __launch_bounds__(32) k_msm_horner(const u32x4* __restrict__ win, msm_geom g, u32x4* __restrict__ result) {
    int r = threadIdx.x & 3;
    win += QQ_PT_Q * (size_t)blockIdx.x * g.K;
    result += QQ_PT_Q * (size_t)blockIdx.x;
    fe acc = coop_load(win + QQ_PT_Q * (size_t)(g.K - 1), r);
    for (int k = g.K - 2; k >= 0; k--) {
        for (int i = 0; i < g.c; i++) acc = coop_dbl(acc, r);
        fe p = coop_load(win + QQ_PT_Q * (size_t)k, r);
        acc = coop_add(acc, coop_to_cached(p, r), r);
    }
    if (threadIdx.x < 4) coop_store(result, r, acc);
}

// The same chain with one limb per lane (ge_warp.cuh): a doubling is two rounds of 8 products per lane and two carry
// normalisations instead of two 72-product carry chains per lane.
__global__ void __launch_bounds__(32) k_msm_horner_warp(const u32x4* __restrict__ win, msm_geom g, u32x4* __restrict__ result) {
    const int q = (threadIdx.x >> 3) & 3, k = threadIdx.x & 7;
    win += QQ_PT_Q * (size_t)blockIdx.x * g.K;
    result += QQ_PT_Q * (size_t)blockIdx.x;
    u32 acc = warp_point_load(win + QQ_PT_Q * (size_t)(g.K - 1));
    for (int w = g.K - 2; w >= 0; w--) {
        u32 p = warp_point_load(win + QQ_PT_Q * (size_t)w);      // issued ahead of the doublings
        u32 pc = warp_to_cached(p, q, k);
#pragma unroll 1
        for (int i = 0; i < g.c; i++) acc = warp_dbl(acc, q, k);
        acc = warp_add(acc, pc, q, k);
    }
    warp_point_store(result, acc);
}
// A segment of the chain for the pipelined tail: windows k_hi - 1 .. k_lo; first: start from win[k_hi - 1], else continue from
// the value in `result` (the ranks above).
__global__ void __launch_bounds__(32) k_msm_horner_warp_seg(const u32x4* __restrict__ win, int c, int k_hi, int k_lo, int first,
                                                            u32x4* __restrict__ result) {
    const int q = (threadIdx.x >> 3) & 3, k = threadIdx.x & 7;
    int w = k_hi - 1;
    u32 acc;
    if (first) acc = warp_point_load(win + QQ_PT_Q * (size_t)w--);
    else acc = warp_point_load(result);
    for (; w >= k_lo; w--) {
        u32 p = warp_point_load(win + QQ_PT_Q * (size_t)w);
        u32 pc = warp_to_cached(p, q, k);
#pragma unroll 1
        for (int i = 0; i < c; i++) acc = warp_dbl(acc, q, k);
        acc = warp_add(acc, pc, q, k);
    }
    warp_point_store(result, acc);
}
// parity hook for the warp-cooperative operations (qq_warp_ops_selftest): warp j computes 2 P_j and P_j + Q_j
__global__ void __launch_bounds__(128) k_warp_selftest(const u32x4* __restrict__ P, const u32x4* __restrict__ Q, size_t n,
                                                       u32x4* __restrict__ dbl_out, u32x4* __restrict__ add_out) {
    const int q = (threadIdx.x >> 3) & 3, k = threadIdx.x & 7;
    size_t j = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= n) return;      // whole warps
    u32 p = warp_point_load(P + QQ_PT_Q * j), qq_ = warp_point_load(Q + QQ_PT_Q * j);
    warp_point_store(dbl_out + QQ_PT_Q * j, warp_dbl(p, q, k));
    warp_point_store(add_out + QQ_PT_Q * j, warp_add(p, warp_to_cached(qq_, q, k), q, k));
}

// ---- grouped form: G independent MSMs in one pass ---------------------------------------------------------------------------
// group_of[i] for terms given as CSR segments (offsets[G + 1]): one thread per term, binary search
__global__ void k_msm_group_of(const unsigned int* __restrict__ offsets, unsigned int G, size_t n, unsigned int* __restrict__ group_of) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned int lo = 0, hi = G;      // offsets[lo] <= i < offsets[hi]
        while (hi - lo > 1) {
            unsigned int mid = (lo + hi) >> 1;
            if (offsets[mid] <= i) lo = mid;
            else hi = mid;
        }
        group_of[i] = lo;
    }
}
// per group: the status of its first failing term (keys preset to ~0), then one status byte and one "is the identity" flag
__global__ void k_msm_group_first_bad(const uint8_t* __restrict__ term_status, const unsigned int* __restrict__ group_of, size_t n,
                                      unsigned long long* __restrict__ keys) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint8_t s = term_status[i];
        if (s) atomicMin(&keys[group_of[i]], (unsigned long long)i * 4 + s);
    }
}
__global__ void k_msm_group_verdicts(const u32x4* __restrict__ results, const unsigned long long* __restrict__ keys, unsigned int G,
                                     u32x4* __restrict__ compressed, uint8_t* __restrict__ status, uint8_t* __restrict__ is_identity) {
    unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= G) return;
    ge_p3 p;
    ge_p3_load(p, results + QQ_PT_Q * (size_t)t);
    const uint8_t st = keys[t] == ~0ull ? 0 : (uint8_t)(keys[t] & 3);
    u32 w[8];
    ristretto_compress(w, p);
    if (st) {
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = 0;
    }
    if (compressed != nullptr) store_words32(compressed, t, w);
    status[t] = st;
    is_identity[t] = (uint8_t)(st == 0 && ge_ristretto_is_identity(p));
}

}  // namespace qq
