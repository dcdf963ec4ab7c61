// CUDA kernels (sm_100a) for the quisquis hot path.  One thread owns one group element; every kernel is a grid-stride
// loop so the grid can be sized to a multiple of the SM count (148 on B200).  Points travel between kernels as
// extended coordinates, 4 x 8 limbs = 128 B (QQ_PT_Q x 16 B), read and written with 128-bit accesses; the compressed API
// buffers (32 B per point / scalar) are read with two 128-bit loads per thread.
//
// The work is bound by integer-multiply issue (IMAD.WIDE.U32 carry chains), not by HBM: see DESIGN.md.
#pragma once
#include "ristretto.cuh"
#include "scalarmult.cuh"
#include "ge_coop.cuh"

namespace qq {

struct idx_map {  // element t -> source index (t / ppi) * pstride + off[t % ppi]
    int ppi, pstride;
    int off[4];
};
__device__ __forceinline__ size_t map_index(const idx_map& m, size_t t) {
    size_t g = t / (size_t)m.ppi;
    int s = (int)(t - g * (size_t)m.ppi);
    int o = m.off[0];
    o = s == 1 ? m.off[1] : o;
    o = s == 2 ? m.off[2] : o;
    o = s == 3 ? m.off[3] : o;
    return g * (size_t)m.pstride + (size_t)o;
}

__device__ __forceinline__ void load_words32(u32 w[8], const u32x4* src, size_t i) {
    const uint4* s = reinterpret_cast<const uint4*>(src) + 2 * i;
    uint4 a = __ldg(s), b = __ldg(s + 1);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
}
__device__ __forceinline__ void store_words32(u32x4* dst, size_t i, const u32 w[8]) {
    uint4* d = reinterpret_cast<uint4*>(dst) + 2 * i;
    d[0] = make_uint4(w[0], w[1], w[2], w[3]);
    d[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// ---- decompress: compressed points (mapped) -> extended points + validity flags -------------------------------
__global__ void __launch_bounds__(256) k_decompress(const u32x4* __restrict__ in, idx_map map, u32x4* __restrict__ pts,
                                                    uint8_t* __restrict__ ok, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        u32 w[8];
        load_words32(w, in, map_index(map, t));
        ge_p3 p;
        u32 v = ristretto_decompress(p, w);
        ge_p3_store(pts + QQ_PT_Q * t, p);
        ok[t] = (uint8_t)v;
    }
}

// ---- variable-base scalar multiplication ------------------------------------------------------------------------
// item t: point pts[map(t)], scalars s0[t / sdiv] (and s1[t / sdiv] when NS == 2); out0[t] = s0 * P, out1[t] = s1 * P.
// The per-thread window table lives in `scratch` (gridDim.x * blockDim.x * 72 x 16 B), so it stays L2-resident
// while the grid-stride loop walks the batch.
struct vb_args {
    const u32x4* pts;
    idx_map map;
    const u32x4* s0;
    const u32x4* s1;
    int sdiv;
    int halve0, halve1;   // multiply by s/2 mod l instead of s (outputs feed the double-and-compress encoder)
    u32x4* out0;
    u32x4* out1;
    u32x4* scratch;
    size_t n;
};
// One 512-thread block per SM, all of whose warps are kept at the same code position by a barrier per item.
// The loop body of a scalar multiplication is ~60 KB of SASS, far beyond the per-scheduler instruction cache; warps that
// drift apart (they do, through table-load latency) each fetch their own copy of it and the kernel stalls on
// instruction fetch: ncu `no_instruction` 3 % at 7 items per thread, 13 % at 55, with the multiply pipe dropping from
// 78 % to 71 % active.  With the barrier the rate is independent of the batch size: 5.06e7 -> 6.18e7 scalar-mults/s at
// 2^21 points (tools/vb_bench.cu, profiles/vb_bench_sync_r01.jsonl).  Every thread runs the same number of rounds
// (the barrier sits at the top of the round; threads past the end skip the work, so a 9-account call is not slowed
// down by 500 idle lanes repeating it).
#define QQ_VB_BLOCK 512
template <int NS, bool SECRET = false>
__global__ void __launch_bounds__(QQ_VB_BLOCK, 1) k_varbase(vb_args a) {
    size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    u32x4* tbl = a.scratch + gtid * (QQ_VB_ENTRIES * QQ_PT_Q);
    size_t rounds = (a.n + stride - 1) / stride;
    for (size_t it = 0; it < rounds; it++) {
        size_t t = gtid + it * stride;
        __syncthreads();
        if (t >= a.n) continue;      // still meets the barrier of every later round
        ge_p3 p, r;
        ge_p3_load(p, a.pts + QQ_PT_Q * map_index(a.map, t));
        vb_build_table(tbl, p);
        u32 s[8];
        load_words32(s, a.s0, t / (size_t)a.sdiv);
        if (a.halve0) sc_halve(s, s);
        vb_scalarmult_t<false, SECRET>(r, tbl, s);
        ge_p3_store(a.out0 + QQ_PT_Q * t, r);
        if (NS == 2) {
            load_words32(s, a.s1, t / (size_t)a.sdiv);
            if (a.halve1) sc_halve(s, s);
            vb_scalarmult_t<false, SECRET>(r, tbl, s);
            ge_p3_store(a.out1 + QQ_PT_Q * t, r);
        }
    }
}

// Two scalars per point through the split tables (vbs_*): 312 doublings per point instead of 504.  The four tables of
// a thread (4.6 KB) are written once and read 128 times; they stream through L2.
template <bool SECRET = false>
__global__ void __launch_bounds__(QQ_VB_BLOCK, 1) k_varbase_split(vb_args a) {
    size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    u32x4* tbl = a.scratch + gtid * QQ_VBS_TABLE_Q;
    size_t rounds = (a.n + stride - 1) / stride;
    for (size_t it = 0; it < rounds; it++) {
        size_t t = gtid + it * stride;
        __syncthreads();
        if (t >= a.n) continue;      // still meets the barrier of every later round
        ge_p3 p, r;
        ge_p3_load(p, a.pts + QQ_PT_Q * map_index(a.map, t));
        vbs_build_tables(tbl, p);
        u32 s[8];
        load_words32(s, a.s0, t / (size_t)a.sdiv);
        if (a.halve0) sc_halve(s, s);
        vbs_scalarmult<SECRET>(r, tbl, s);
        ge_p3_store(a.out0 + QQ_PT_Q * t, r);
        load_words32(s, a.s1, t / (size_t)a.sdiv);
        if (a.halve1) sc_halve(s, s);
        vbs_scalarmult<SECRET>(r, tbl, s);
        ge_p3_store(a.out1 + QQ_PT_Q * t, r);
    }
}

// Small batches (9-account anonymity sets, a block's worth of transactions): the chain of 316 group operations of one
// scalar multiplication is pure latency and most of the GPU's lanes are idle, so FOUR adjacent lanes share one
// (point, scalar) job (ge_coop.cuh: lane r owns coordinate r, every group operation is two rounds of one field
// multiplication per lane instead of 8 multiplications in a row).  Job j = point j / ns with scalar j % ns; each
// group builds its own cached table {1..8}P in shared memory (1 KB per group).  Same digits (signed radix 16), same
// results as k_varbase / k_varbase_split; chosen by launch_varbase while the jobs leave lanes idle.
#define QQ_VBC_GROUP_Q 64   // 8 entries x 4 lanes x 2 u32x4
__global__ void __launch_bounds__(128) k_varbase_coop(vb_args a, int ns) {
    extern __shared__ __align__(16) u32x4 vbc_tbl[];
    const int r = threadIdx.x & 3;
    u32x4* tb = vbc_tbl + (threadIdx.x >> 2) * QQ_VBC_GROUP_Q;
    size_t job = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2, njobs = a.n * (size_t)ns;
    bool act = job < njobs;
    size_t j = act ? job : 0;     // idle groups of the last warp repeat job 0 (the shuffles are warp-wide) and store nothing
    size_t t = j / (size_t)ns;
    int which = (int)(j - t * (size_t)ns);
    fe p = coop_load(a.pts + QQ_PT_Q * map_index(a.map, t), r);
    fe c1 = coop_to_cached(p, r), q = p;
    {
        int o = 2 * r;
        fe_store(tb, o, c1);
    }
#pragma unroll 1
    for (int i = 2; i <= 8; i++) {
        q = coop_add(q, c1, r);
        fe c = coop_to_cached(q, r);
        int o = ((i - 1) * 4 + r) * 2;
        fe_store(tb, o, c);
    }
    __syncwarp();
    u32 s[8], rr[9], w[8];
    load_words32(s, which ? a.s1 : a.s0, t / (size_t)a.sdiv);
    if (which ? a.halve1 : a.halve0) sc_halve(s, s);
    sc_recode_bias<4, 64>(rr, s);        // rr[8] == 0 for s < 2^253
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = rr[i];
    fe acc = coop_identity(r);
#pragma unroll 1
    for (int k = 63; k >= 0; k--) {
        if (k != 63) {
#pragma unroll 1
            for (int d = 0; d < 4; d++) acc = coop_dbl(acc, r);
        }
        int d = (int)(w[7] >> 28) - 8;     // signed digit in [-8, 8); then shift the register left by one nibble
#pragma unroll
        for (int i = 7; i > 0; i--) w[i] = (w[i] << 4) | (w[i - 1] >> 28);
        w[0] <<= 4;
        u32 neg = d < 0 ? 1u : 0u;
        int idx = d < 0 ? -d : d;
        // cached lanes: 0 = Y - X, 1 = Y + X, 2 = 2 Z, 3 = 2d T.  Negation swaps lanes 0 / 1 and negates lane 3.
        int lane = r < 2 ? (r ^ (int)neg) : r;
        fe c;
        int o = ((idx > 0 ? idx - 1 : 0) * 4 + lane) * 2;
        fe_load(tb, o, c);
        if (r == 3) fe_cneg(c, neg);
        fe idc;
        fe_0(idc);
        idc.v[0] = r == 2 ? 2u : (r == 3 ? 0u : 1u);
        c = fe_sel(idx == 0, idc, c);
        acc = coop_add(acc, c, r);
    }
    if (act) coop_store((which ? a.out1 : a.out0) + QQ_PT_Q * t, r, acc);
}

// ---- fixed-base scalar multiplication: table staged in shared memory ---------------------------------------------
template <int W, bool SECRET = false>
__global__ void __launch_bounds__(512) k_fixedbase(const u32* __restrict__ tbl_g, const u32x4* __restrict__ s, int halve,
                                                   u32x4* __restrict__ out, size_t n) {
    extern __shared__ __align__(16) u32 tbl_s[];
    const int words = ((256 + W - 1) / W) * ((1 << (W - 1)) + 1) * QQ_NIELS_WORDS;
    {
        const uint4* g = reinterpret_cast<const uint4*>(tbl_g);
        uint4* d = reinterpret_cast<uint4*>(tbl_s);
        for (int i = threadIdx.x; i < words / 4; i += blockDim.x) d[i] = __ldg(g + i);
        for (int i = (words / 4) * 4 + threadIdx.x; i < words; i += blockDim.x) tbl_s[i] = tbl_g[i];
    }
    __syncthreads();
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        u32 w[8];
        load_words32(w, s, t);
        if (halve) sc_halve(w, w);
        ge_p3 r;
        fb_scalarmult<W, SECRET>(r, tbl_s, w);
        ge_p3_store(out + QQ_PT_Q * t, r);
    }
}

// one thread per table entry (k, j): tbl[(k * ENT + j) * 30 ..] = (j << (W k)) * base
__global__ void k_fb_build(u32* tbl, const u32x4* base_compressed, int W, int nw, int ent) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nw * ent) return;
    u32 w[8];
    load_words32(w, base_compressed, 0);
    ge_p3 b;
    ristretto_decompress(b, w);
    fb_build_entry(tbl + (size_t)t * QQ_NIELS_WORDS, b, W, t / ent, t % ent);
}

// ---- finish kernels: combine + compress -------------------------------------------------------------------------
// A finish job produces one compressed point per item t:  enc( sum of up to 3 extended points ).
// Source k of output t is  src[k].base + QQ_PT_Q * map(src[k].map, t)  (null base = absent), optionally negated.
struct fin_src {
    const u32x4* base;
    idx_map map;
    int negate;
};
struct fin_args {
    fin_src src[3];
    u32x4* out;        // compressed outputs
    idx_map omap;      // where output t is written (index into `out`, in 32-byte units)
    const uint8_t* bad;  // per-element status (indexed by t / bdiv): non-zero -> write zeros instead
    int bdiv;
    size_t n;
};
__device__ __forceinline__ void fin_eval(ge_p3& q, const fin_args& a, size_t t) {
    ge_p3_load(q, a.src[0].base + QQ_PT_Q * map_index(a.src[0].map, t));
    if (a.src[0].negate) ge_neg(q, q);
#pragma unroll 1
    for (int k = 1; k < 3; k++) {
        if (a.src[k].base == nullptr) break;
        ge_p3 p;
        ge_p3_load(p, a.src[k].base + QQ_PT_Q * map_index(a.src[k].map, t));
        if (a.src[k].negate) ge_neg(p, p);
        ge_cached c;
        ge_to_cached(c, p);
        ge_add(q, q, c);
    }
}
__global__ void __launch_bounds__(256, 2) k_finish_compress(fin_args a) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < a.n; t += stride) {
        ge_p3 q;
        fin_eval(q, a, t);
        u32 w[8];
        ristretto_compress(w, q);
        if (a.bad != nullptr && a.bad[t / (size_t)a.bdiv] != 0) {
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = 0;
        }
        store_words32(a.out, map_index(a.omap, t), w);
    }
}
// projective Ristretto equality of two extended points: flag[t] = 1 if equal
__global__ void __launch_bounds__(256) k_points_equal(const u32x4* __restrict__ a, const u32x4* __restrict__ b,
                                                      uint8_t* __restrict__ flag, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        ge_p3 p, q;
        ge_p3_load(p, a + QQ_PT_Q * t);
        ge_p3_load(q, b + QQ_PT_Q * t);
        flag[t] = (uint8_t)ge_ristretto_eq(p, q);
    }
}

// ---- hash-to-group tail: 64 uniform bytes -> compressed point (RistrettoPoint::from_uniform_bytes) ---------------------
__global__ void __launch_bounds__(256, 2) k_from_uniform(const u32x4* __restrict__ in, u32x4* __restrict__ out, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        u32 w[16];
        load_words32(w, in, 2 * t);
        load_words32(w + 8, in, 2 * t + 1);
        ge_p3 p;
        ristretto_from_uniform(p, w);
        u32 o[8];
        ristretto_compress(o, p);
        store_words32(out, t, o);
    }
}

// ---- status assembly ---------------------------------------------------------------------------------------------
// status[i] = BAD_SCALAR if any of the (up to 3) scalar arrays holds a non-canonical scalar at i,
//             else BAD_POINT if any of the `npts` validity flags ok[i * npts + j] is 0, else 0.
struct st_args {
    const u32x4* sc[3];
    const uint8_t* ok;
    int npts;
    uint8_t* status;
    size_t n;
};
__global__ void k_status(st_args a) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
        uint8_t st = 0;
        for (int j = 0; j < a.npts; j++)
            if (!a.ok[i * a.npts + j]) st = 1;
        for (int k = 0; k < 3; k++) {
            if (a.sc[k] == nullptr) continue;
            u32 w[8];
            load_words32(w, a.sc[k], i);
            if (!sc_is_canonical(w)) st = 2;
        }
        a.status[i] = st;
    }
}
// verify_account verdict (reference order: keypair check, then commitment; src/accounts/accounts.rs:81-84):
// pre = status from k_status over (sk, bl) and ok flags [gr, c]; eqflag[i] = keypair equal, eqflag[n + i] = commitment equal.
__global__ void k_verify_account_status(const uint8_t* pre_scalar, const uint8_t* ok2, const uint8_t* eqflag,
                                        uint8_t* status, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint8_t st = 0;
        if (pre_scalar[i] == 2) st = 2;
        else if (!ok2[2 * i]) st = 1;
        else if (!eqflag[i]) st = 3;
        else if (!ok2[2 * i + 1]) st = 1;
        else if (!eqflag[n + i]) st = 4;
        status[i] = st;
    }
}
// pairs of flags -> status: pre[i] != 0 wins, else (f[2i] & f[2i+1]) ? 0 : fail_code
__global__ void k_pair_status(const uint8_t* pre, const uint8_t* f, uint8_t fail_code, uint8_t* status, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        status[i] = pre[i] ? pre[i] : ((f[2 * i] && f[2 * i + 1]) ? 0 : fail_code);
}
__global__ void k_or_status(uint8_t* dst, const uint8_t* a, const uint8_t* b, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint8_t x = a[i], y = b[i];
        dst[i] = (x == 2 || y == 2) ? 2 : ((x | y) ? (x ? x : y) : 0);
    }
}

// delta / epsilon pk halves (reference src/accounts/accounts.rs:210,215): delta_i.pk = acc_i.pk, epsilon_i.pk = base_pk
struct pk_bytes {
    uint4 q[4];
};
__global__ void k_copy_pk(const u32x4* __restrict__ acc, u32x4* __restrict__ delta, u32x4* __restrict__ eps,
                          pk_bytes base, const uint8_t* __restrict__ status, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint4* a = reinterpret_cast<const uint4*>(acc) + 8 * i;
        uint4* d = reinterpret_cast<uint4*>(delta) + 8 * i;
        uint4* e = reinterpret_cast<uint4*>(eps) + 8 * i;
        bool bad = status[i] != 0;
        uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            d[k] = bad ? z : __ldg(a + k);
            e[k] = bad ? z : base.q[k];
        }
    }
}
// first failing term of an MSM (the reference stops at the first None / the oracle at the first bad term):
// key = index * 4 + code, minimum wins
__global__ void k_first_bad(const uint8_t* __restrict__ term_status, size_t n, unsigned long long* key) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long best = ~0ull;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint8_t s = term_status[i];
        if (s) {
            unsigned long long k = (unsigned long long)i * 4 + s;
            best = k < best ? k : best;
        }
    }
    if (best != ~0ull) atomicMin(key, best);
}
__global__ void k_key_to_status(const unsigned long long* key, uint8_t* status) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *status = (*key == ~0ull) ? 0 : (uint8_t)(*key & 3);
}

// ---- point-sum reduction -------------------------------------------------------------------------------------------
// out[b] = sum over a strided slice of in[0..n) (block b handles elements b, b + gridDim, ...; threads stride inside)
// followed by a shared-memory tree over the block.  Used for the identity check and for MSM fallbacks.
__global__ void __launch_bounds__(128) k_point_sum(const u32x4* __restrict__ in, idx_map map, size_t n,
                                                   u32x4* __restrict__ out) {
    __shared__ u32x4 sm[128 * QQ_PT_Q];
    ge_p3 acc;
    ge_identity(acc);
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        ge_p3 p;
        ge_p3_load(p, in + QQ_PT_Q * map_index(map, t));
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
// single thread: out_xyzt (4 x 32 canonical bytes), out compressed, identity flag, from one extended point
__global__ void k_point_export(const u32x4* in, u32x4* xyzt, u32x4* compressed, uint8_t* is_identity) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    ge_p3 p;
    ge_p3_load(p, in);
    u32 w[8];
    if (xyzt != nullptr) {
        fe_towords(w, p.X); store_words32(xyzt, 0, w);
        fe_towords(w, p.Y); store_words32(xyzt, 1, w);
        fe_towords(w, p.Z); store_words32(xyzt, 2, w);
        fe_towords(w, p.T); store_words32(xyzt, 3, w);
    }
    if (compressed != nullptr) {
        ristretto_compress(w, p);
        store_words32(compressed, 0, w);
    }
    if (is_identity != nullptr) *is_identity = (uint8_t)ge_ristretto_is_identity(p);
}
// device-side delivery of an MSM result: nbytes of `src` (zeros when *status != 0) and the status byte
__global__ void k_emit_result(const uint8_t* __restrict__ src, int nbytes, const uint8_t* __restrict__ status,
                              uint8_t* __restrict__ out, uint8_t* __restrict__ status_out) {
    int t = threadIdx.x;
    uint8_t st = *status;
    if (t < nbytes) out[t] = st ? 0 : src[t];
    if (t == 0) *status_out = st;
}
// k points given as canonical X,Y,Z,T bytes -> extended limbs
__global__ void k_point_import(const u32x4* xyzt, u32x4* pts, size_t k) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= k) return;
    ge_p3 p;
    u32 w[8];
    load_words32(w, xyzt, 4 * t + 0); fe_fromwords(p.X, w);
    load_words32(w, xyzt, 4 * t + 1); fe_fromwords(p.Y, w);
    load_words32(w, xyzt, 4 * t + 2); fe_fromwords(p.Z, w);
    load_words32(w, xyzt, 4 * t + 3); fe_fromwords(p.T, w);
    ge_p3_store(pts + QQ_PT_Q * t, p);
}
// segmented sum: instance j = sum of terms offsets[j]..offsets[j+1]-1, compressed; bad if any term flag set
__global__ void __launch_bounds__(256, 2) k_segment_sum_compress(const u32x4* __restrict__ terms,
                                                              const uint32_t* __restrict__ offsets,
                                                              const uint8_t* __restrict__ term_status,
                                                              u32x4* __restrict__ out, uint8_t* __restrict__ status,
                                                              size_t m) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
        ge_p3 acc;
        ge_identity(acc);
        uint8_t st = 0;
        for (uint32_t t = offsets[j]; t < offsets[j + 1]; t++) {
            ge_p3 p;
            ge_p3_load(p, terms + QQ_PT_Q * (size_t)t);
            ge_cached c;
            ge_to_cached(c, p);
            ge_add(acc, acc, c);
            uint8_t s = term_status[t];
            st = (s == 2 || st == 2) ? 2 : (st | s);
        }
        u32 w[8];
        ristretto_compress(w, acc);
        if (st) {
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = 0;
        }
        store_words32(out, j, w);
        status[j] = st;
    }
}

// ---- Straus (interleaved windows) for many small MSMs -----------------------------------------------------------------
// One thread per instance: the 2..9 terms of an instance (reference call sites: src/accounts/verifier.rs:165-880,
// src/shuffle/*.rs) share ONE doubling chain -- 252 doublings + 64 table additions per term instead of
// 252 doublings per term.  Per-term window tables and recoded scalars live in a per-thread global scratch slab
// (QQ_STRAUS_KMAX terms x 74 x 16 B); instances with more terms are processed in chunks of QQ_STRAUS_KMAX.
// dalek counterpart: backend/serial/scalar_mul/straus.rs.
#define QQ_STRAUS_KMAX 10
#define QQ_STRAUS_TABLE_Q (QQ_VB_ENTRIES * QQ_PT_Q)
#define QQ_STRAUS_TERM_Q (QQ_STRAUS_TABLE_Q + 2)   // 72 x 16 B table + 2 x 16 B recoded scalar
struct straus_args {
    const u32x4* pts;          // decompressed terms
    const u32x4* scalars;
    const uint32_t* offsets;   // m + 1
    const uint8_t* term_status;
    u32x4* out;                // m compressed points (classic encoder), unused when half_out != nullptr
    u32x4* half_out;           // m extended points = sum (s_i / 2) P_i, for the double-and-compress batch encoder
    uint8_t* status;           // m
    u32x4* scratch;
    const unsigned int* order; // instances sorted by term count (descending) so a warp's lanes do equal work
    size_t m;
    int kc;                    // k_straus_coop: terms per pass (1 .. QQ_STC_KC); its shared memory is 8 x kc term slots per warp
};
// signed 64-bit values -> canonical scalars (v mod l), branch-free: |v| and l - |v| are both formed, the sign selects
__global__ void k_i64_to_scalars(const long long* __restrict__ v, u32x4* __restrict__ out, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        long long x = v[i];
        u32 neg = (u32)((unsigned long long)x >> 63);
        unsigned long long mag = ((unsigned long long)x ^ (0ull - (unsigned long long)neg)) + neg;      // |x| (2^63 for INT64_MIN)
        u32 a[8] = {(u32)mag, (u32)(mag >> 32), 0, 0, 0, 0, 0, 0}, b[8];
        u32 borrow = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            u64 d = (u64)sc_l_word(k) - a[k] - borrow;
            b[k] = (u32)d;
            borrow = (u32)(d >> 63);
        }
        u32 use_b = neg & (u32)(mag != 0);
        u32 m = 0u - use_b;
        u32 w[8];
#pragma unroll
        for (int k = 0; k < 8; k++) w[k] = a[k] ^ (m & (a[k] ^ b[k]));
        store_words32(out, i, w);
    }
}
// status of an MSM evaluated in `group` parts: a non-canonical scalar wins over an undecodable point
__global__ void k_group_status(const uint8_t* __restrict__ part, int group, uint8_t* __restrict__ status, size_t mo) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < mo; j += stride) {
        uint8_t st = 0;
        for (int g = 0; g < group; g++) {
            uint8_t s = part[j * group + g];
            st = (s == 2 || st == 2) ? 2 : (st | s);
        }
        status[j] = st;
    }
}
__global__ void k_seg_counts(const uint32_t* __restrict__ offsets, size_t m, unsigned int* __restrict__ counts) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) counts[j] = offsets[j + 1] - offsets[j];
}
template <int BLOCK, int MINB, bool SYNC>
__global__ void __launch_bounds__(BLOCK, MINB) k_straus(straus_args a) {
    size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    u32x4* slab = a.scratch + gtid * (size_t)(QQ_STRAUS_KMAX * QQ_STRAUS_TERM_Q);
    size_t rounds = (a.m + stride - 1) / stride;
    for (size_t it = 0; it < rounds; it++) {
        size_t jj = gtid + it * stride;
        if (SYNC) __syncthreads();   // instances are ordered by term count: the warps of a block restart together
        if (jj >= a.m) continue;
        size_t j = a.order[jj];
        uint32_t lo = a.offsets[j], hi = a.offsets[j + 1];
        ge_p3 total;
        ge_identity(total);
        uint8_t st = 0;
        for (uint32_t c0 = lo; c0 < hi; c0 += QQ_STRAUS_KMAX) {
            int k = (int)((hi - c0) < QQ_STRAUS_KMAX ? (hi - c0) : QQ_STRAUS_KMAX);
            for (int t = 0; t < k; t++) {
                ge_p3 p;
                ge_p3_load(p, a.pts + QQ_PT_Q * (size_t)(c0 + t));
                u32x4* tb = slab + (size_t)t * QQ_STRAUS_TERM_Q;
                vb_build_table(tb, p);
                u32 s[8], rr[9];
                load_words32(s, a.scalars, c0 + t);
                if (a.half_out != nullptr) sc_halve(s, s);
                sc_recode_bias<4, 64>(rr, s);
                store_words32(tb + QQ_STRAUS_TABLE_Q, 0, rr);
                uint8_t ts = a.term_status[c0 + t];
                st = (ts == 2 || st == 2) ? 2 : (st | ts);
            }
            ge_p3 r;
            ge_identity(r);
#pragma unroll 1
            for (int w = 63; w >= 0; w--) {
                if (w != 63) {
                    ge_dbl<false>(r, r);
                    ge_dbl<false>(r, r);
                    ge_dbl<false>(r, r);
                    ge_dbl<true>(r, r);
                }
#pragma unroll 1
                for (int t = 0; t < k; t++) {
                    const u32x4* tb = slab + (size_t)t * QQ_STRAUS_TERM_Q;
                    u32 word = reinterpret_cast<const u32*>(tb + QQ_STRAUS_TABLE_Q)[w >> 3];
                    int d = (int)((word >> ((w & 7) * 4)) & 15u) - 8;
                    u32 neg = d < 0 ? 1u : 0u;
                    u32 idx = (u32)(d < 0 ? -d : d);
                    ge_cached c;
                    ge_cached_load(c, tb + QQ_PT_Q * idx);
                    ge_cached_cneg(c, neg);
                    ge_add(r, r, c);
                }
            }
            ge_cached cr;
            ge_to_cached(cr, r);
            ge_add(total, total, cr);
        }
        a.status[j] = st;
        if (a.half_out != nullptr) {
            ge_p3_store(a.half_out + QQ_PT_Q * j, total);
            continue;
        }
        u32 wds[8];
        ristretto_compress(wds, total);
        if (st) {
#pragma unroll
            for (int i = 0; i < 8; i++) wds[i] = 0;
        }
        store_words32(a.out, j, wds);
    }
}

// Few instances (the 18-36 MSMs of ONE sigma proof, the ~42 of one shuffle proof): four lanes per instance, as in
// k_varbase_coop.  A 32-thread block holds 8 instances; the per-term tables {1..8}P (cached, lane-sliced) and recoded
// scalars of a chunk of up to QQ_STC_KC terms live in shared memory.  The shuffles are warp-wide, so the 8 instances of
// a warp run the term count of the longest one (absent terms add the identity); no ordering pass is needed.
// Output: half_out (the scalars are halved, the batch encoder doubles), as in k_straus.
#define QQ_STC_KC 9
#define QQ_STC_TERM_Q (QQ_VBC_GROUP_Q + 2)                 // table + 8 recoded words
#define QQ_STC_SMEM_BYTES (8 * QQ_STC_KC * QQ_STC_TERM_Q * 16)
// a.kc terms per pass: an instance with more terms takes several passes (each pays the 252 doublings).  The caller sizes kc to the
// largest instance when it knows it: with the full 9 slots a block needs 76 KB of shared memory and only two warps fit an SM -
// 8 192 two-term instances (the exact MSMs of 4 096 shuffle proofs) then run in four waves, 1.09 ms; with kc = 2 (17 KB) in one.
__global__ void __launch_bounds__(32) k_straus_coop(straus_args a) {
    extern __shared__ __align__(16) u32x4 stc_smem[];
    const int r = threadIdx.x & 3, grp = threadIdx.x >> 2;
    const int KC = a.kc;
    u32x4* slab = stc_smem + (size_t)grp * ((size_t)KC * QQ_STC_TERM_Q);
    size_t inst = (size_t)blockIdx.x * 8 + grp;
    bool act = inst < a.m;
    size_t j = act ? inst : 0;
    uint32_t lo = a.offsets[j], nt = a.offsets[j + 1] - lo;
    uint32_t ntmax = nt;
#pragma unroll
    for (int sft = 4; sft < 32; sft <<= 1) {
        uint32_t o = __shfl_xor_sync(QQ_COOP_MASK, ntmax, sft);
        ntmax = o > ntmax ? o : ntmax;
    }
    fe total = coop_identity(r);
    uint8_t st = 0;
    for (uint32_t c0 = 0; c0 < ntmax; c0 += (uint32_t)KC) {
        int kw = (int)(ntmax - c0 < (uint32_t)KC ? ntmax - c0 : (uint32_t)KC);               // warp-uniform
        int k = c0 < nt ? (int)(nt - c0 < (uint32_t)KC ? nt - c0 : (uint32_t)KC) : 0;        // this instance's
#pragma unroll 1
        for (int t = 0; t < kw; t++) {
            bool have = t < k;
            size_t ti = (size_t)lo + c0 + (have ? t : 0);
            u32x4* tb = slab + (size_t)t * QQ_STC_TERM_Q;
            fe p = coop_identity(r);
            u32 rr[9];
#pragma unroll
            for (int i = 0; i < 9; i++) rr[i] = 0x88888888u;     // every digit 0
            if (have) {
                p = coop_load(a.pts + QQ_PT_Q * ti, r);
                u32 sc[8];
                load_words32(sc, a.scalars, ti);
                sc_halve(sc, sc);
                sc_recode_bias<4, 64>(rr, sc);
                uint8_t ts = a.term_status[ti];
                st = (ts == 2 || st == 2) ? 2 : (st | ts);
            }
            if (r == 0) store_words32(tb + QQ_VBC_GROUP_Q, 0, rr);
            fe c1 = coop_to_cached(p, r), q = p;
            {
                int o = 2 * r;
                fe_store(tb, o, c1);
            }
#pragma unroll 1
            for (int i = 2; i <= 8; i++) {
                q = coop_add(q, c1, r);
                fe c = coop_to_cached(q, r);
                int o = ((i - 1) * 4 + r) * 2;
                fe_store(tb, o, c);
            }
        }
        __syncwarp();
        fe acc = coop_identity(r);
#pragma unroll 1
        for (int w = 63; w >= 0; w--) {
            if (w != 63) {
#pragma unroll 1
                for (int d = 0; d < 4; d++) acc = coop_dbl(acc, r);
            }
#pragma unroll 1
            for (int t = 0; t < kw; t++) {
                const u32x4* tb = slab + (size_t)t * QQ_STC_TERM_Q;
                u32 word = reinterpret_cast<const u32*>(tb + QQ_VBC_GROUP_Q)[w >> 3];
                int d = (int)((word >> ((w & 7) * 4)) & 15u) - 8;
                u32 neg = d < 0 ? 1u : 0u;
                int idx = d < 0 ? -d : d;
                int lane = r < 2 ? (r ^ (int)neg) : r;
                fe c;
                int o = ((idx > 0 ? idx - 1 : 0) * 4 + lane) * 2;
                fe_load(tb, o, c);
                if (r == 3) fe_cneg(c, neg);
                fe idc;
                fe_0(idc);
                idc.v[0] = r == 2 ? 2u : (r == 3 ? 0u : 1u);
                c = fe_sel(idx == 0, idc, c);
                acc = coop_add(acc, c, r);
            }
        }
        total = coop_add(total, coop_to_cached(acc, r), r);
        __syncwarp();
    }
    if (act) {
        coop_store(a.half_out + QQ_PT_Q * j, r, total);
        if (r == 0) a.status[j] = st;
    }
}

// ---- integer-pipe peak micro-benchmark (roofline denominator) ------------------------------------------------------
// 8 independent dependent chains per thread and nothing else in the loop body:
//   MODE 0: c = c * a + b          -> IMAD          (32-bit multiply-add, the "IMAD unit" of the cost model)
//   MODE 1: w = lo(w) * a + w      -> IMAD.WIDE.U32 (32x32->64 product; issues at half the IMAD rate)
template <int MODE>
__global__ void __launch_bounds__(256) k_imad_peak(u32* out, u32 seed, int outer) {
    u32 a[8], b[8], c[8];
    u64 w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = (seed * 2654435761u + i * 40503u + threadIdx.x) | 1u;
        b[i] = seed + i * 7919u + blockIdx.x;
        c[i] = i + seed;
        w[i] = ((u64)c[i] << 32) | b[i];
    }
    for (int o = 0; o < outer; o++) {
#pragma unroll
        for (int k = 0; k < 32; k++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(c[i]) : "r"(a[i]), "r"(b[i]));
                if (MODE == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((u32)(w[i] >> 32)), "r"(a[i]));
            }
        }
    }
    u32 acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= c[i] ^ (u32)w[i] ^ (u32)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

}  // namespace qq
