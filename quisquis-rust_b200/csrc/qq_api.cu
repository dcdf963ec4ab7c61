// libqq_b200.so -- host side of the C ABI declared in include/qq_b200.h.
// Owns the device workspace, the fixed-base tables and the stream; sequences the kernels of kernels.cuh / msm.cuh.
// No CPU fallback: every entry point fails with QQ_ERR_NODEVICE / QQ_ERR_CUDA when the GPU path is unavailable.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <random>
#include <deque>
#include <condition_variable>
#include <memory>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/qq_b200.h"
#include "kernels.cuh"
#include "msm.cuh"
#include "compress_batch.cuh"
#include "fixedbase_big.cuh"
#include "keccak_host.hpp"
#include "merlin_host.hpp"
#include "sc_host.hpp"
#include "decommit.cuh"
#include "rangeproof.cuh"
#include "shuffle_verify.cuh"
#include "sigma_verify.cuh"

using namespace qq;

// The large-window table (15 additions at W = 16) is used at every batch size when it exists: measured 0.05 ms against
// 0.10 (n = 9) ... 0.24 ms (n = 1000) for the shared-memory table, whose 136 KB staging per block is what a small batch
// pays for (profiles/small_batch_r01.jsonl).  The shared-memory kernel remains for qq_fixed_base_set_window(.., 0).
#define QQ_FBT_MIN_BATCH 1
// Batch encoder: up to this many items every thread inverts its own denominator (k_dc_direct, one launch) instead of the
// multi-level shared inversion (five dependent launches)
#define QQ_DC_DIRECT_MAX ((size_t)ctx->sms * 128)
#define QQ_FB_W 6  // window width of the shared-memory fixed-base tables (43 windows x 33 entries x 96 B = 136 KB)
enum { FAM_DEC = 0, FAM_VB = 1, FAM_FB = 2, FAM_FIN = 3, FAM_MSM_BUCKET = 4, FAM_MSM_REDUCE = 5, FAM_TRANSCRIPT = 6, FAM_COUNT = 7 };

// Grow-only pool of page-locked host buffers for the job lists the batched verifiers upload (acquired by the caller's
// thread, released by the GPU worker thread of the pipelined shuffle verifier: guarded by a mutex).
struct pinned_pool {
    struct blk { uint8_t* p; size_t cap; bool used; };
    std::mutex mu;
    std::vector<blk> blocks;
    uint8_t* acquire(size_t bytes) {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& b : blocks)
            if (!b.used && b.cap >= bytes) {
                b.used = true;
                return b.p;
            }
        void* p = nullptr;
        size_t cap = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
        if (cudaHostAlloc(&p, cap, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;      // the caller falls back to pageable memory
        }
        blocks.push_back({(uint8_t*)p, cap, true});
        return (uint8_t*)p;
    }
    void release(uint8_t* p) {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& b : blocks)
            if (b.p == p) b.used = false;
    }
    void destroy() {
        for (auto& b : blocks) cudaFreeHost(b.p);
        blocks.clear();
    }
};

struct qq_ctx {
    pinned_pool pin;
    int device = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;   // upload / download streams of the pipelined host entry points
    cudaEvent_t msm_ev[12] = {nullptr};
    cudaStream_t msm_hi = nullptr;           // high-priority stream: the MSM's counting sort, concurrent with the decompression
    long msm_split_min = 1 << 17;            // QQ_MSM_SPLIT_MIN / QQ_MSM_TAIL_PCT in the environment override
    int msm_sort_bpsm = 3;                   // blocks per SM of the sort kernels while they run under the decompression
    int msm_tail_pct = 30;                   // share of the points decompressed under the counting sort
    std::vector<cudaEvent_t> pipe_ev;
    char* ws = nullptr;
    size_t ws_cap = 0, ws_off = 0;
    char* io = nullptr;      // staging slab for the host-pointer entry points (grow-only)
    size_t io_cap = 0;
    cudaEvent_t user_ev[8] = {nullptr};
    u32* fb_tbl[2] = {nullptr, nullptr};      // shared-memory-sized tables (W = QQ_FB_W)
    u32x4* bsgs_enc = nullptr;                // baby-step table of decommit_value: enc(j B), j < 2^20 (built on first use)
    u32* bsgs_slots = nullptr;
    u32x4* msm_res = nullptr;                 // per-call MSM result point + export bytes (outside the workspace slab)
    uint8_t* msm_small = nullptr;
    uint8_t* fb = nullptr;                    // scratch of the verifiers' failure paths (grouped checks), grown on demand
    size_t fb_cap = 0;
    u32x4* fbt[2] = {nullptr, nullptr};       // large-window tables in L2 / HBM (fixedbase_big.cuh), optional
    fbt_geom fbt_g[2] = {{0, 0, 0}, {0, 0, 0}};
    u32x4* half_base[2] = {nullptr, nullptr};  // ((l + 1) / 2) * Base as affine Niels, for the 64-bit fixed-base path
    uint8_t base_pk[64];
    uint64_t launches = 0;
    float last_ms = 0.f;
    float breakdown[FAM_COUNT] = {0};
    std::string err;
    std::vector<cudaEvent_t> ev_pool;
    struct span { int fam; int e0, e1; };
    std::vector<span> spans;
    int ev_used = 0;
    int vb_blocks_per_sm[3] = {0, 0, 0};
    bool xpc_ready = false;                    // VectorPedersenGens::new(4): H, G_vec (compressed), derived on first use
    uint8_t xpc_h[32], xpc_g[96];
    bool bp_ready = false;                     // BulletproofGens::new(64, 16) (compressed, party-major), derived on first use
    std::vector<uint8_t> bp_g, bp_h;
    // qq_transcript_capture arms (pointer, capacity in states) for the NEXT entry point only: ENTER() of every entry point moves
    // them to capture_live / capture_cap (and disarms); only sigma_verify reads the live pair, so a call that returns early, a
    // verifier that keeps no transcript or an exception between the two calls can never leave a stale pointer behind
    uint8_t* transcript_capture = nullptr;
    size_t transcript_capture_cap = 0;
    uint8_t* capture_live = nullptr;
    size_t capture_cap = 0;
    bool stc_ready = false;                    // k_straus_coop's shared-memory opt-in done
    int straus_minb = 4;                       // k_straus build for more than one wave of instances (QQ_STRAUS_MINB)
    int stc_per_sm = 64;                       // segmented MSMs: four-lane cooperative kernel up to this many MSMs per SM (QQ_STRAUS_COOP_PER_SM)
    // qq_msm_set_shifted: memory one prepared point set may take for its shifted form.  Default = what stays L2-resident
    // (126 MB L2): measured, shifted / plain ms: 2^10-2^12 points 0.44 / 0.62, 2^16 0.61 / 0.76, 2^18 (403 MB) 1.19 / 1.14, 2^20 (1.6 GB)
    // 3.29 / 2.35 - beyond L2 the 16 x larger gather footprint costs more than the reductions and the Horner chain save.
    size_t msm_shift_budget = (size_t)112 << 20;
    bool msm_use_shifted = true;
    // pipelined tail of the large MSM: ranks of windows (QQ_MSM_PIPE_RANKS = 2..4), from msm_pipe_min terms on.  Off by default:
    // measured 3.94 against 3.97 ms at 2^20 and 55.2 against 54.4 ms at 2^24 (the reductions it hides are real multiply work that
    // then competes with the accumulation; profiles/msm_pipelined_tail_r02.jsonl)
    int msm_pipe_ranks = 1;
    long msm_pipe_min = 1 << 17;
    unsigned transcript_lanes = 32;  // threads per block of the one-thread-per-proof transcript kernels (QQ_TRANSCRIPT_LANES: 8, 16, 32)
    // range-proof verifier: the MSM's points decompressed on a copy stream beside the transcripts.  -1 (default): for batches of up to
    // 1 024 transcripts (1 / 64 / 512 proofs: -0.07 ms, 5 %); 1: always (4 096 x 16 values 2.85 -> 2.68 ms, but 4 096 single-value proofs
    // 1.79 -> 2.65 ms: with a transcript warp on every SM the decompression slows the transcript kernel down by more than it hides);
    // 0: never.  QQ_VERIFY_EARLY_DECOMPRESS.
    int verify_early_decompress = -1;
    bool small_fanout = true;        // small batches: independent launches of one call on the copy streams (QQ_SMALL_FANOUT)
    bool msm_horner_warp = true;     // window Horner with one limb per lane (ge_warp.cuh); false: the four-lane form (A/B knob)
    bool secret_mode = false;                  // qq_set_secret_mode: constant-time table access for scalars that are secrets
    int vb_blocks_per_sm_secret[3] = {0, 0, 0};
    bool verify_aggregate = true;              // qq_verify_set_aggregation: identity equations of the shuffle proofs in one weighted Pippenger MSM
    bool verify_host_transcripts = false;      // qq_verify_set_transcripts(ctx, 0): per-proof phases of the shuffle verifier on the host threads
    uint8_t* d_shuffle_gens = nullptr;         // B | B_blinding | H | G[0..3) of VectorPedersenGens::new(4), device copy for the transcript kernels
    long vbc_max_jobs = -1;                    // four-lane cooperative variable base up to this many scalar mults (< 0: sms * 160)
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                         \
            return e_ == cudaErrorMemoryAllocation ? QQ_ERR_NOMEM : QQ_ERR_CUDA;                   \
        }                                                                                          \
    } while (0)
#define CKQ(call)                 \
    do {                          \
        int r_ = (call);          \
        if (r_ != QQ_OK) return r_; \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- workspace: one growable slab, bump-allocated per call -----------------------------------------------------
static int ws_begin(qq_ctx* ctx, size_t bytes) {
    bytes = align_up(bytes, 256) + 4096;
    if (bytes > ctx->ws_cap) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->ws) CK(cudaFree(ctx->ws));
        ctx->ws = nullptr;
        ctx->ws_cap = 0;
        size_t want = bytes + bytes / 8;
        CK(cudaMalloc((void**)&ctx->ws, want));
        ctx->ws_cap = want;
    }
    ctx->ws_off = 0;
    return QQ_OK;
}
template <class T>
static T* ws_take(qq_ctx* ctx, size_t bytes) {
    size_t off = align_up(ctx->ws_off, 256);
    ctx->ws_off = off + bytes;
    return reinterpret_cast<T*>(ctx->ws + off);
}
static size_t ws_need(std::initializer_list<size_t> parts) {
    size_t t = 0;
    for (size_t p : parts) t += align_up(p, 256) + 256;
    return t;
}

// ---- timing spans ------------------------------------------------------------------------------------------------
static int ev_get(qq_ctx* ctx) {
    if (ctx->ev_used == (int)ctx->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->ev_pool.push_back(e);
    }
    return ctx->ev_used++;
}
static void span_begin(qq_ctx* ctx, int fam) {
    qq_ctx::span s;
    s.fam = fam;
    s.e0 = ev_get(ctx);
    s.e1 = -1;
    cudaEventRecord(ctx->ev_pool[s.e0], ctx->stream);
    ctx->spans.push_back(s);
}
static void span_end(qq_ctx* ctx) {
    qq_ctx::span& s = ctx->spans.back();
    s.e1 = ev_get(ctx);
    cudaEventRecord(ctx->ev_pool[s.e1], ctx->stream);
}
static void call_begin(qq_ctx* ctx) {
    ctx->spans.clear();
    ctx->ev_used = 0;
}
static int call_end(qq_ctx* ctx) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    for (int i = 0; i < FAM_COUNT; i++) ctx->breakdown[i] = 0.f;
    ctx->last_ms = 0.f;
    for (auto& s : ctx->spans) {
        float ms = 0.f;
        if (s.e1 >= 0) cudaEventElapsedTime(&ms, ctx->ev_pool[s.e0], ctx->ev_pool[s.e1]);
        ctx->breakdown[s.fam] += ms;
        ctx->last_ms += ms;
    }
    return QQ_OK;
}

static inline int grid_for(size_t n, int block, int cap) {
    size_t g = (n + block - 1) / block;
    if (g > (size_t)cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}
static idx_map imap(int ppi, int pstride, int o0 = 0, int o1 = 1, int o2 = 2, int o3 = 3) {
    idx_map m;
    m.ppi = ppi;
    m.pstride = pstride;
    m.off[0] = o0; m.off[1] = o1; m.off[2] = o2; m.off[3] = o3;
    return m;
}
static const idx_map IDENT = {1, 1, {0, 0, 0, 0}};

// ---- kernel launch helpers --------------------------------------------------------------------------------------
static int launch_decompress(qq_ctx* ctx, const void* in, idx_map map, u32x4* pts, uint8_t* ok, size_t n) {
    if (n == 0) return QQ_OK;
    span_begin(ctx, FAM_DEC);
    k_decompress<<<grid_for(n, 256, ctx->sms * 64), 256, 0, ctx->stream>>>((const u32x4*)in, map, pts, ok, n);
    span_end(ctx);
    ctx->launches++;
    CK(cudaGetLastError());
    return QQ_OK;
}
static size_t vb_scratch_bytes(qq_ctx* ctx, int ns) {
    int bps = ctx->vb_blocks_per_sm[ns] > ctx->vb_blocks_per_sm_secret[ns] ? ctx->vb_blocks_per_sm[ns] : ctx->vb_blocks_per_sm_secret[ns];
    return (size_t)ctx->sms * bps * QQ_VB_BLOCK * (ns == 2 ? QQ_VBS_TABLE_WORDS : QQ_VB_TABLE_WORDS) * 4;
}
static int launch_varbase(qq_ctx* ctx, int ns, const u32x4* pts, idx_map map, const void* s0, const void* s1, int sdiv,
                          u32x4* out0, u32x4* out1, u32x4* scratch, size_t n, int halve0 = 0, int halve1 = 0) {
    if (n == 0) return QQ_OK;
    vb_args a;
    a.pts = pts; a.map = map; a.s0 = (const u32x4*)s0; a.s1 = (const u32x4*)s1; a.sdiv = sdiv;
    a.halve0 = halve0; a.halve1 = halve1;
    a.out0 = out0; a.out1 = out1; a.scratch = scratch; a.n = n;
    // large batches: one 512-thread block per SM (lockstep, see kernels.cuh).  Small batches are latency-bound chains:
    // narrower blocks spread the few warps over all schedulers (one warp per scheduler up to 4 * sms warps).
    // While the lanes are not all busy, four lanes per scalar multiplication (k_varbase_coop): measured faster up to
    // ~19 000 scalar mults (0.66 against 0.98 ms), slower from ~38 000 (1.23 against 1.01 ms); switch at 160 per SM.
    size_t coop_max = ctx->vbc_max_jobs < 0 ? (size_t)ctx->sms * 160 : (size_t)ctx->vbc_max_jobs;
    // secret mode: only the kernels whose table access is a masked scan of the whole table (the cooperative kernel reads its
    // shared-memory table by digit)
    if (!ctx->secret_mode && n * (size_t)ns <= coop_max) {
        size_t threads = n * (size_t)ns * 4;
        int cb = 32;
        while (cb < 128 && (threads + cb - 1) / cb > (size_t)ctx->sms * 4) cb *= 2;
        span_begin(ctx, FAM_VB);
        k_varbase_coop<<<(unsigned)((threads + cb - 1) / cb), cb, (size_t)(cb / 4) * QQ_VBC_GROUP_Q * 16, ctx->stream>>>(a, ns);
        span_end(ctx);
        ctx->launches++;
        CK(cudaGetLastError());
        return QQ_OK;
    }
    int block = QQ_VB_BLOCK;
    int grid = ctx->sms * (ctx->secret_mode ? ctx->vb_blocks_per_sm_secret[ns] : ctx->vb_blocks_per_sm[ns]);
    if ((size_t)grid * QQ_VB_BLOCK > n) {
        block = 32;
        while (block < QQ_VB_BLOCK && (n + block - 1) / block > (size_t)ctx->sms * 4) block *= 2;
        grid = (int)((n + block - 1) / block);
    }
    span_begin(ctx, FAM_VB);
    if (ctx->secret_mode) {
        if (ns == 1) k_varbase<1, true><<<grid, block, 0, ctx->stream>>>(a);
        else k_varbase_split<true><<<grid, block, 0, ctx->stream>>>(a);
    } else {
        if (ns == 1) k_varbase<1><<<grid, block, 0, ctx->stream>>>(a);
        else k_varbase_split<false><<<grid, block, 0, ctx->stream>>>(a);
    }
    span_end(ctx);
    ctx->launches++;
    CK(cudaGetLastError());
    return QQ_OK;
}
static size_t fb_table_words() { return (size_t)fb_num_windows(QQ_FB_W) * fb_entries(QQ_FB_W) * QQ_NIELS_WORDS; }
static int launch_fixedbase(qq_ctx* ctx, int which, const void* s, u32x4* out, size_t n, int halve = 0) {
    if (n == 0) return QQ_OK;
    if (!ctx->secret_mode && ctx->fbt[which] != nullptr && n >= QQ_FBT_MIN_BATCH) {
        span_begin(ctx, FAM_FB);
        k_fixedbase_big<<<grid_for(n, QQ_FBT_BLOCK, ctx->sms * 4), QQ_FBT_BLOCK, 0, ctx->stream>>>(ctx->fbt[which], ctx->fbt_g[which],
                                                                                (const u32x4*)s, halve, out, n);
        span_end(ctx);
        ctx->launches++;
        CK(cudaGetLastError());
        return QQ_OK;
    }
    size_t smem = fb_table_words() * 4;
    int grid = (int)((n + 511) / 512);
    if (grid > ctx->sms) grid = ctx->sms;
    span_begin(ctx, FAM_FB);
    // secret mode: every entry of a window of the shared-memory table is read and masked (33 entries per window)
    if (ctx->secret_mode) k_fixedbase<QQ_FB_W, true><<<grid, 512, smem, ctx->stream>>>(ctx->fb_tbl[which], (const u32x4*)s, halve, out, n);
    else k_fixedbase<QQ_FB_W><<<grid, 512, smem, ctx->stream>>>(ctx->fb_tbl[which], (const u32x4*)s, halve, out, n);
    span_end(ctx);
    ctx->launches++;
    CK(cudaGetLastError());
    return QQ_OK;
}
static fin_src fsrc(const u32x4* base, idx_map map, int negate = 0) {
    fin_src s;
    s.base = base; s.map = map; s.negate = negate;
    return s;
}
static const fin_src FNONE = {nullptr, {1, 1, {0, 0, 0, 0}}, 0};
static int launch_finish(qq_ctx* ctx, fin_src a, fin_src b, fin_src c, void* out, idx_map omap, const uint8_t* bad,
                         size_t n) {
    if (n == 0) return QQ_OK;
    fin_args f;
    f.src[0] = a; f.src[1] = b; f.src[2] = c;
    f.out = (u32x4*)out; f.omap = omap; f.bad = bad; f.bdiv = 1; f.n = n;
    span_begin(ctx, FAM_FIN);
    k_finish_compress<<<grid_for(n, 256, ctx->sms * 64), 256, 0, ctx->stream>>>(f);
    span_end(ctx);
    ctx->launches++;
    CK(cudaGetLastError());
    return QQ_OK;
}
// ---- double-and-compress finish (compress_batch.cuh): enc(2 * sum of sources) with one inversion per launch ------
#define QQ_BINV_C 64
struct dc_ws {
    u32x4 *state, *w, *prefix, *lv_vals, *lv_prefix;
    uint8_t* zflag;
};
static size_t dc_levels_q(size_t n) {  // u32x4 units needed by the upper levels of the inversion tree
    size_t q = 0;
    while (n > QQ_BINV_C) {
        n = (n + QQ_BINV_C - 1) / QQ_BINV_C;
        q += 2 * n;
    }
    return q + 2;
}
static size_t dc_scratch_bytes(size_t n) {
    if (n == 0) n = 1;
    return ws_need({n * QQ_DC_STATE_Q * 16, n * 32, n * 32, dc_levels_q(n) * 16, dc_levels_q(n) * 16, n});
}
static dc_ws dc_take(qq_ctx* ctx, size_t n) {
    if (n == 0) n = 1;
    dc_ws d;
    d.state = ws_take<u32x4>(ctx, n * QQ_DC_STATE_Q * 16);
    d.w = ws_take<u32x4>(ctx, n * 32);
    d.prefix = ws_take<u32x4>(ctx, n * 32);
    d.lv_vals = ws_take<u32x4>(ctx, dc_levels_q(n) * 16);
    d.lv_prefix = ws_take<u32x4>(ctx, dc_levels_q(n) * 16);
    d.zflag = ws_take<uint8_t>(ctx, n);
    return d;
}
// inv[i] = 1 / vals[i] for n nonzero field elements; `prefix` (n) receives the inverses
static int launch_batch_invert(qq_ctx* ctx, const dc_ws& d, size_t n) {
    const u32x4* vals[12];
    u32x4* pre[12];
    size_t cnt[12];
    vals[0] = d.w; pre[0] = d.prefix; cnt[0] = n;
    int L = 0;
    size_t off = 0;
    while (cnt[L] > QQ_BINV_C) {
        cnt[L + 1] = (cnt[L] + QQ_BINV_C - 1) / QQ_BINV_C;
        u32x4* tot = d.lv_vals + off;
        pre[L + 1] = d.lv_prefix + off;
        off += 2 * cnt[L + 1];
        k_binv_up<<<grid_for(cnt[L + 1], 128, ctx->sms * 8), 128, 0, ctx->stream>>>(vals[L], cnt[L], QQ_BINV_C, pre[L], tot);
        vals[L + 1] = tot;
        L++;
        ctx->launches++;
    }
    k_binv_top<<<1, 32, 0, ctx->stream>>>(vals[L], cnt[L], pre[L], pre[L]);
    ctx->launches++;
    for (int l = L - 1; l >= 0; l--) {
        k_binv_down<<<grid_for(cnt[l + 1], 128, ctx->sms * 8), 128, 0, ctx->stream>>>(vals[l], pre[l], pre[l + 1], cnt[l], QQ_BINV_C, pre[l]);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    return QQ_OK;
}
// out[omap(t)] = enc(2 * (a + b + c)(t)); with expect != nullptr: flag[t] = (enc == expect[emap(t)]) instead
static int launch_finish_dbl(qq_ctx* ctx, const dc_ws& d, fin_src a, fin_src b, fin_src c, void* out, idx_map omap,
                             const uint8_t* bad, int bdiv, size_t n, const void* expect = nullptr,
                             idx_map emap = {1, 1, {0, 0, 0, 0}}, uint8_t* flag = nullptr) {
    if (n == 0) return QQ_OK;
    fin_args f;
    f.src[0] = a; f.src[1] = b; f.src[2] = c;
    f.out = (u32x4*)out; f.omap = omap; f.bad = bad; f.bdiv = bdiv; f.n = n;
    span_begin(ctx, FAM_FIN);
    if (n <= QQ_DC_DIRECT_MAX && ctx->vbc_max_jobs != 0) {
        k_dc_direct<<<(unsigned)((n + 31) / 32), 32, 0, ctx->stream>>>(f, bad, bdiv, (u32x4*)out, omap, (const u32x4*)expect, emap, flag);
        span_end(ctx);
        ctx->launches++;
        CK(cudaGetLastError());
        return QQ_OK;
    }
    k_dc_prepare<<<grid_for(n, 256, ctx->sms * 64), 256, 0, ctx->stream>>>(f, d.state, d.w, d.zflag);
    ctx->launches++;
    CKQ(launch_batch_invert(ctx, d, n));
    k_dc_finish<<<grid_for(n, 256, ctx->sms * 64), 256, 0, ctx->stream>>>(d.state, d.prefix, d.zflag, bad, bdiv, (u32x4*)out, omap,
                                                                        (const u32x4*)expect, emap, flag, n);
    span_end(ctx);
    ctx->launches++;
    CK(cudaGetLastError());
    return QQ_OK;
}
static int launch_status(qq_ctx* ctx, const void* s0, const void* s1, const void* s2, const uint8_t* ok, int npts,
                         uint8_t* status, size_t n) {
    if (n == 0) return QQ_OK;
    st_args a;
    a.sc[0] = (const u32x4*)s0; a.sc[1] = (const u32x4*)s1; a.sc[2] = (const u32x4*)s2;
    a.ok = ok; a.npts = npts; a.status = status; a.n = n;
    k_status<<<grid_for(n, 256, ctx->sms * 16), 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return QQ_OK;
}
// sum of n extended points (mapped) -> one extended point at `result` (device, 128 B)
static int launch_point_sum(qq_ctx* ctx, const u32x4* pts, idx_map map, size_t n, u32x4* partials /*>= sms*2*10*/,
                            u32x4* result) {
    int blocks = grid_for(n, 128, ctx->sms * 2);
    span_begin(ctx, FAM_MSM_REDUCE);
    k_point_sum<<<blocks, 128, 0, ctx->stream>>>(pts, map, n, partials);
    k_point_sum<<<1, 128, 0, ctx->stream>>>(partials, IDENT, (size_t)blocks, result);
    span_end(ctx);
    ctx->launches += 2;
    CK(cudaGetLastError());
    return QQ_OK;
}

// ---- large-window fixed-base tables (fixedbase_big.cuh) ---------------------------------------------------------------
static int launch_batch_invert(qq_ctx* ctx, const dc_ws& d, size_t n);
static const uint8_t QQ_BASE_PK_BYTES[64] = {
    0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
    0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76,
    0x8c, 0x92, 0x40, 0xb4, 0x56, 0xa9, 0xe6, 0xdc, 0x65, 0xc3, 0x77, 0xa1, 0x04, 0x8d, 0x74, 0x5f,
    0x94, 0xa0, 0x8c, 0xdb, 0x7f, 0x44, 0xcb, 0xcd, 0x7b, 0x46, 0xf3, 0x40, 0x48, 0x87, 0x11, 0x34};
// device buffers released on every exit path of a builder function
struct dev_bufs {
    std::vector<void*> p;
    template <class T>
    cudaError_t alloc(T** out, size_t bytes) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, bytes ? bytes : 16);
        if (e == cudaSuccess) p.push_back(q);
        *out = (T*)q;
        return e;
    }
    void release(void* keep) {   // hand one buffer over to the caller
        for (auto& q : p)
            if (q == keep) q = nullptr;
    }
    ~dev_bufs() {
        for (void* q : p)
            if (q) cudaFree(q);
    }
};
static int fbt_build(qq_ctx* ctx, int which, int W) {
    if (ctx->fbt[which]) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFree(ctx->fbt[which]));
        ctx->fbt[which] = nullptr;
        ctx->fbt_g[which] = fbt_geom{0, 0, 0};
    }
    if (W == 0) return QQ_OK;
    fbt_geom g;
    g.W = W;
    g.NW = (255 + W - 1) / W;
    g.ENT = (1u << (W - 1)) + 1u;
    size_t entries = (size_t)g.NW * g.ENT;
    dev_bufs bufs;
    u32x4* tbl = nullptr;
    CK(bufs.alloc(&tbl, entries * QQ_NIELS_STRIDE_Q * 16));
    const unsigned int SLICE = 1u << 22;
    size_t slice = g.ENT < SLICE ? g.ENT : SLICE;
    u32x4 *dbase = nullptr, *bases = nullptr, *ext = nullptr;
    dc_ws d;
    d.state = nullptr; d.zflag = nullptr;
    CK(bufs.alloc(&dbase, 64));
    CK(bufs.alloc(&bases, (size_t)g.NW * QQ_PT_BYTES));
    CK(bufs.alloc(&ext, slice * QQ_PT_BYTES));
    CK(bufs.alloc(&d.w, slice * 32));
    CK(bufs.alloc(&d.prefix, slice * 32));
    CK(bufs.alloc(&d.lv_vals, dc_levels_q(slice) * 16));
    CK(bufs.alloc(&d.lv_prefix, dc_levels_q(slice) * 16));
    CK(cudaMemcpyAsync(dbase, QQ_BASE_PK_BYTES, 64, cudaMemcpyHostToDevice, ctx->stream));
    k_fbt_window_bases<<<1, 32, 0, ctx->stream>>>(dbase + 2 * which, g, bases);
    ctx->launches++;
    for (int k = 0; k < g.NW; k++) {
        for (size_t j0 = 0; j0 < g.ENT; j0 += slice) {
            unsigned int cnt = (unsigned int)(g.ENT - j0 < slice ? g.ENT - j0 : slice);
            unsigned int chunks = (cnt + QQ_FBT_CHUNK - 1) / QQ_FBT_CHUNK;
            k_fbt_points<<<(chunks + 127) / 128, 128, 0, ctx->stream>>>(bases, k, (unsigned int)j0, cnt, ext, d.w);
            CKQ(launch_batch_invert(ctx, d, cnt));
            k_fbt_normalize<<<(cnt + 255) / 256, 256, 0, ctx->stream>>>(ext, d.prefix, cnt,
                                                                      tbl + ((size_t)k * g.ENT + j0) * QQ_NIELS_STRIDE_Q);
            ctx->launches += 2;
        }
    }
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    bufs.release(tbl);
    ctx->fbt[which] = tbl;
    ctx->fbt_g[which] = g;
    return QQ_OK;
}
// Secret-scalar mode (the wallet's update / commitment scalars and the provers' blindings are secrets).
extern "C" int qq_set_secret_mode(qq_ctx* ctx, int on) {
    if (!ctx) return QQ_ERR_ARG;
    ctx->secret_mode = on != 0;
    return QQ_OK;
}
extern "C" int qq_secret_mode(const qq_ctx* ctx) { return ctx && ctx->secret_mode ? 1 : 0; }
extern "C" int qq_varbase_set_coop_limit(qq_ctx* ctx, long max_scalar_mults) {
    if (!ctx) return QQ_ERR_ARG;
    ctx->vbc_max_jobs = max_scalar_mults;
    return QQ_OK;
}
extern "C" int qq_fixed_base_set_window(qq_ctx* ctx, int which, int window_bits) {
    if (!ctx || (which != 0 && which != 1) || (window_bits != 0 && (window_bits < 8 || window_bits > 28))) return QQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    return fbt_build(ctx, which, window_bits);
}
extern "C" int qq_fixed_base_window(const qq_ctx* ctx, int which) {
    return (ctx && (which == 0 || which == 1)) ? ctx->fbt_g[which].W : 0;
}

// =================================================================================================================
// context
// =================================================================================================================
extern "C" int qq_init(qq_ctx** out, int device) {
    if (!out) return QQ_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return QQ_ERR_NODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return QQ_ERR_NODEVICE;
    if (prop.major != 10) {
        fprintf(stderr, "qq_b200: device %d is sm_%d%d; this library contains sm_100a code only\n", device, prop.major,
                prop.minor);
        return QQ_ERR_NODEVICE;
    }
    qq_ctx* ctx = new qq_ctx();
    ctx->device = device;
    ctx->sms = prop.multiProcessorCount;
    auto fail = [&](int code) {
        fprintf(stderr, "qq_b200: init failed: %s\n", ctx->err.c_str());
        delete ctx;
        return code;
    };
    auto body = [&]() -> int {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        {
            int lo_pr = 0, hi_pr = 0;
            CK(cudaDeviceGetStreamPriorityRange(&lo_pr, &hi_pr));
            CK(cudaStreamCreateWithPriority(&ctx->msm_hi, cudaStreamNonBlocking, hi_pr));
            if (const char* e = getenv("QQ_MSM_SPLIT_MIN")) ctx->msm_split_min = atol(e);
            if (const char* e = getenv("QQ_STRAUS_COOP_PER_SM")) ctx->stc_per_sm = atoi(e);
            if (const char* e = getenv("QQ_STRAUS_MINB")) ctx->straus_minb = atoi(e);
            if (const char* e = getenv("QQ_MSM_TAIL_PCT")) ctx->msm_tail_pct = atoi(e);
            if (const char* e = getenv("QQ_MSM_SORT_BPSM")) ctx->msm_sort_bpsm = atoi(e);
            if (const char* e = getenv("QQ_MSM_SHIFT_BUDGET_MB")) ctx->msm_shift_budget = (size_t)atol(e) << 20;
            if (const char* e = getenv("QQ_MSM_HORNER_WARP")) ctx->msm_horner_warp = atoi(e) != 0;
            if (const char* e = getenv("QQ_SMALL_FANOUT")) ctx->small_fanout = atoi(e) != 0;
            if (const char* e = getenv("QQ_VERIFY_EARLY_DECOMPRESS")) ctx->verify_early_decompress = atoi(e);
            if (const char* e = getenv("QQ_TRANSCRIPT_LANES")) { int v = atoi(e); if (v == 4 || v == 8 || v == 16 || v == 32) ctx->transcript_lanes = (unsigned)v; }
            if (const char* e = getenv("QQ_MSM_PIPE_RANKS")) { int v = atoi(e); if (v >= 1 && v <= 4) ctx->msm_pipe_ranks = v; }
            if (const char* e = getenv("QQ_MSM_PIPE_MIN")) { long v = atol(e); if (v >= 1) ctx->msm_pipe_min = v; }
            if (const char* e = getenv("QQ_VERIFY_HOST_TRANSCRIPTS")) ctx->verify_host_transcripts = atoi(e) != 0;
            if (const char* e = getenv("QQ_VERIFY_AGGREGATE")) ctx->verify_aggregate = atoi(e) != 0;
        }
        for (int i = 0; i < 12; i++) CK(cudaEventCreateWithFlags(&ctx->msm_ev[i], cudaEventDisableTiming));
        CK(cudaFuncSetAttribute(k_msm_sum_levels, cudaFuncAttributeMaxDynamicSharedMemorySize, 512 * QQ_PT_BYTES));
        CK(cudaFuncSetAttribute(k_fixedbase<QQ_FB_W>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(fb_table_words() * 4)));
        CK(cudaFuncSetAttribute(k_fixedbase<QQ_FB_W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(fb_table_words() * 4)));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_varbase<1>, QQ_VB_BLOCK, 0));
        ctx->vb_blocks_per_sm[1] = occ > 0 ? occ : 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_varbase_split<false>, QQ_VB_BLOCK, 0));
        ctx->vb_blocks_per_sm[2] = occ > 0 ? occ : 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_varbase<1, true>, QQ_VB_BLOCK, 0));
        ctx->vb_blocks_per_sm_secret[1] = occ > 0 ? occ : 1;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_varbase_split<true>, QQ_VB_BLOCK, 0));
        ctx->vb_blocks_per_sm_secret[2] = occ > 0 ? occ : 1;
        if (const char* e = getenv("QQ_SECRET_MODE")) ctx->secret_mode = atoi(e) != 0;
        // fixed-base tables for B and H, built on the device from their compressed encodings
        const uint8_t* BASE_PK = QQ_BASE_PK_BYTES;
        memcpy(ctx->base_pk, BASE_PK, 64);
        u32x4* dbase = nullptr;
        CK(cudaMalloc((void**)&dbase, 64));
        CK(cudaMemcpy(dbase, BASE_PK, 64, cudaMemcpyHostToDevice));
        int nw = fb_num_windows(QQ_FB_W), ent = fb_entries(QQ_FB_W);
        for (int b = 0; b < 2; b++) {
            CK(cudaMalloc((void**)&ctx->fb_tbl[b], align_up(fb_table_words() * 4, 256)));
            k_fb_build<<<(nw * ent + 63) / 64, 64, 0, ctx->stream>>>(ctx->fb_tbl[b], dbase + 2 * b, QQ_FB_W, nw, ent);
            ctx->launches++;
        }
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaGetLastError());
        CK(cudaFree(dbase));
        CK(cudaMalloc((void**)&ctx->msm_res, QQ_PT_BYTES));
        CK(cudaMalloc((void**)&ctx->msm_small, 256));
        // large-window tables: 16-bit windows (50 MB per base, L2 resident) unless QQ_FB_WINDOW says otherwise
        int W = 16;
        if (const char* e = getenv("QQ_FB_WINDOW")) W = atoi(e);
        if (W != 0 && (W < 8 || W > 28)) { ctx->err = "QQ_FB_WINDOW must be 0 or in [8, 28]"; return QQ_ERR_ARG; }
        for (int b = 0; b < 2; b++) CKQ(fbt_build(ctx, b, W));
        // B/2 and H/2 = ((l + 1) / 2) * Base through the shared-memory table, normalised to affine Niels
        {
            static const uint8_t HALF_L1[32] = {0xf7, 0xe9, 0x7a, 0x2e, 0x8d, 0x31, 0x09, 0x2c, 0x6b, 0xce, 0x7b, 0x51, 0xef, 0x7c, 0x6f, 0x0a,
                                                0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0x08};   // (l + 1) / 2, little endian
            u32x4 *ds = nullptr, *dp = nullptr;
            CK(cudaMalloc((void**)&ds, 32));
            CK(cudaMalloc((void**)&dp, QQ_PT_BYTES));
            CK(cudaMemcpy(ds, HALF_L1, 32, cudaMemcpyHostToDevice));
            for (int b = 0; b < 2; b++) {
                CK(cudaMalloc((void**)&ctx->half_base[b], QQ_NIELS_STRIDE_Q * 16));
                size_t smem = fb_table_words() * 4;
                k_fixedbase<QQ_FB_W><<<1, 512, smem, ctx->stream>>>(ctx->fb_tbl[b], ds, 0, dp, 1);
                k_point_to_niels<<<1, 32, 0, ctx->stream>>>(dp, ctx->half_base[b]);
                ctx->launches += 2;
            }
            CK(cudaStreamSynchronize(ctx->stream));
            CK(cudaGetLastError());
            cudaFree(ds);
            cudaFree(dp);
        }
        return QQ_OK;
    };
    int r = body();
    if (r != QQ_OK) return fail(r);
    *out = ctx;
    return QQ_OK;
}

extern "C" void qq_destroy(qq_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->io) cudaFree(ctx->io);
    if (ctx->bsgs_enc) cudaFree(ctx->bsgs_enc);
    if (ctx->bsgs_slots) cudaFree(ctx->bsgs_slots);
    if (ctx->msm_res) cudaFree(ctx->msm_res);
    if (ctx->msm_small) cudaFree(ctx->msm_small);
    if (ctx->fb) cudaFree(ctx->fb);
    if (ctx->d_shuffle_gens) cudaFree(ctx->d_shuffle_gens);
    for (int i = 0; i < 8; i++)
        if (ctx->user_ev[i]) cudaEventDestroy(ctx->user_ev[i]);
    for (int b = 0; b < 2; b++) {
        if (ctx->fb_tbl[b]) cudaFree(ctx->fb_tbl[b]);
        if (ctx->fbt[b]) cudaFree(ctx->fbt[b]);
        if (ctx->half_base[b]) cudaFree(ctx->half_base[b]);
    }
    for (int i = 0; i < 12; i++)
        if (ctx->msm_ev[i]) cudaEventDestroy(ctx->msm_ev[i]);
    for (auto e : ctx->pipe_ev) cudaEventDestroy(e);
    ctx->pin.destroy();
    if (ctx->msm_hi) cudaStreamDestroy(ctx->msm_hi);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" const char* qq_last_error(const qq_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
extern "C" int qq_device_sm_count(const qq_ctx* ctx) { return ctx ? ctx->sms : 0; }
extern "C" uint64_t qq_launch_count(const qq_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" float qq_last_kernel_ms(const qq_ctx* ctx) { return ctx ? ctx->last_ms : 0.f; }
extern "C" int qq_last_kernel_breakdown(const qq_ctx* ctx, float* ms, int cap) {
    if (!ctx || !ms) return 0;
    int k = cap < FAM_COUNT ? cap : FAM_COUNT;
    for (int i = 0; i < k; i++) ms[i] = ctx->breakdown[i];
    return k;
}
extern "C" int qq_dev_alloc(qq_ctx* ctx, void** dptr, size_t bytes) {
    if (!ctx || !dptr) return QQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMalloc(dptr, bytes ? bytes : 16));
    return QQ_OK;
}
extern "C" int qq_dev_free(qq_ctx* ctx, void* dptr) {
    if (!ctx) return QQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaFree(dptr));
    return QQ_OK;
}
extern "C" int qq_dev_upload(qq_ctx* ctx, void* dptr, const void* host, size_t bytes) {
    if (!ctx) return QQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return QQ_OK;
}
extern "C" int qq_dev_download(qq_ctx* ctx, void* host, const void* dptr, size_t bytes) {
    if (!ctx) return QQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return QQ_OK;
}
extern "C" int qq_event_record(qq_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot >= 8) return QQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->user_ev[slot]) CK(cudaEventCreate(&ctx->user_ev[slot]));
    CK(cudaEventRecord(ctx->user_ev[slot], ctx->stream));
    return QQ_OK;
}
extern "C" int qq_event_elapsed_ms(qq_ctx* ctx, int slot_a, int slot_b, float* ms) {
    if (!ctx || !ms || slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8) return QQ_ERR_ARG;
    if (!ctx->user_ev[slot_a] || !ctx->user_ev[slot_b]) return QQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->user_ev[slot_b]));
    CK(cudaEventElapsedTime(ms, ctx->user_ev[slot_a], ctx->user_ev[slot_b]));
    return QQ_OK;
}
extern "C" int qq_measure_imad_peak(qq_ctx* ctx, double* wide_ops_per_s, double* lo_ops_per_s) {
    if (!ctx) return QQ_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int grid = ctx->sms * 4, outer = 4000;
    u32* out = nullptr;
    CK(cudaMalloc((void**)&out, (size_t)grid * 256 * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double ops = (double)grid * 256 * (double)outer * 32 * 8;
    double best[2] = {0, 0};
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 4; rep++) {
            CK(cudaEventRecord(e0, ctx->stream));
            if (mode == 0) k_imad_peak<0><<<grid, 256, 0, ctx->stream>>>(out, 1234u + rep, outer);
            else k_imad_peak<1><<<grid, 256, 0, ctx->stream>>>(out, 1234u + rep, outer);
            CK(cudaEventRecord(e1, ctx->stream));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            double r = ops / (ms * 1e-3);
            if (rep > 0 && r > best[mode]) best[mode] = r;
            ctx->launches++;
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CK(cudaFree(out));
    if (lo_ops_per_s) *lo_ops_per_s = best[0];
    if (wide_ops_per_s) *wide_ops_per_s = best[1];
    return QQ_OK;
}

// =================================================================================================================
// device-pointer cores.  Each processes [0, n) in chunks so the workspace stays bounded.
// =================================================================================================================
#define QQ_CHUNK ((size_t)1 << 20)
static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
#define REQUIRE(c) \
    do {           \
        if (!(c)) { ctx->err = "bad argument: " #c; return QQ_ERR_ARG; } \
    } while (0)

// out = s * (both points of each 64-byte pair)      [update_public_key, Mul for ElGamalCommitment]
static int core_scale_pairs(qq_ctx* ctx, const uint8_t* pk, const uint8_t* s, uint8_t* out, uint8_t* status, size_t n) {
    for (size_t base = 0; base < n; base += QQ_CHUNK) {
        size_t m = n - base < QQ_CHUNK ? n - base : QQ_CHUNK;
        CKQ(ws_begin(ctx, ws_need({2 * m * QQ_PT_BYTES, 2 * m * QQ_PT_BYTES, 2 * m, vb_scratch_bytes(ctx, 1)}) + dc_scratch_bytes(2 * m)));
        u32x4* P = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);
        u32x4* R = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);
        uint8_t* ok = ws_take<uint8_t>(ctx, 2 * m);
        u32x4* scratch = ws_take<u32x4>(ctx, vb_scratch_bytes(ctx, 1));
        dc_ws dc = dc_take(ctx, 2 * m);
        const uint8_t* pk_c = pk + base * 64;
        const uint8_t* s_c = s + base * 32;
        CKQ(launch_decompress(ctx, pk_c, IDENT, P, ok, 2 * m));
        CKQ(launch_status(ctx, s_c, nullptr, nullptr, ok, 2, status + base, m));
        // R = (s/2) * P, encoded as enc(2 R): no square root on the output side (compress_batch.cuh)
        CKQ(launch_varbase(ctx, 1, P, IDENT, s_c, nullptr, 2, R, nullptr, scratch, 2 * m, 1));
        CKQ(launch_finish_dbl(ctx, dc, fsrc(R, IDENT), FNONE, FNONE, out + base * 64, IDENT, status + base, 2, 2 * m));
    }
    return QQ_OK;
}

// run a launch helper on another stream of the context (the helpers launch on ctx->stream)
template <class F>
static int on_stream(qq_ctx* ctx, cudaStream_t s, F&& f) {
    cudaStream_t keep = ctx->stream;
    ctx->stream = s;
    int rc = f();
    ctx->stream = keep;
    return rc;
}
// small batches: independent launches of one call fan out over the copy streams (QQ_SMALL_FANOUT=0 keeps one stream)
static bool small_fanout(qq_ctx* ctx, size_t m) {
    return ctx->small_fanout && ctx->copy_in != nullptr && ctx->copy_out != nullptr && m <= (size_t)ctx->sms * 16;
}
static int core_generate_commitment(qq_ctx* ctx, const uint8_t* pk, const uint8_t* r, const uint8_t* v, uint8_t* out,
                                    uint8_t* status, size_t n) {
    for (size_t base = 0; base < n; base += QQ_CHUNK) {
        size_t m = n - base < QQ_CHUNK ? n - base : QQ_CHUNK;
        CKQ(ws_begin(ctx, ws_need({2 * m * QQ_PT_BYTES, 2 * m * QQ_PT_BYTES, m * QQ_PT_BYTES, 2 * m, vb_scratch_bytes(ctx, 1)}) + 2 * dc_scratch_bytes(m)));
        u32x4* P = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);
        u32x4* R = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);
        u32x4* F = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);
        uint8_t* ok = ws_take<uint8_t>(ctx, 2 * m);
        u32x4* scratch = ws_take<u32x4>(ctx, vb_scratch_bytes(ctx, 1));
        dc_ws dc = dc_take(ctx, m);
        dc_ws dc2 = dc_take(ctx, m);      // the second encoder's own inversion tree: the two run side by side
        const uint8_t* pk_c = pk + base * 64;
        const uint8_t* r_c = r + base * 32;
        const uint8_t* v_c = v + base * 32;
        // small batches: the fixed-base walk and the second encoder run on a copy stream (see core_update_account)
        // small batches: the fixed-base walk and the second encoder run on a copy stream (see core_update_account); large ones: the
        // second encoder only (its batch inversion is a chain of one-block launches, 0.4 ms during which the GPU is idle)
        const bool fan = small_fanout(ctx, m);
        const bool side = ctx->small_fanout && ctx->copy_in != nullptr;
        if (fan) {
            CK(cudaEventRecord(ctx->msm_ev[8], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_in, ctx->msm_ev[8], 0));
            CKQ(on_stream(ctx, ctx->copy_in, [&] { return launch_fixedbase(ctx, QQ_BASE_B, v_c, F, m, 1); }));
        }
        CKQ(launch_decompress(ctx, pk_c, IDENT, P, ok, 2 * m));
        CKQ(launch_status(ctx, r_c, v_c, nullptr, ok, 2, status + base, m));
        // every term carries half its scalar; the outputs are enc(2 * sum)
        CKQ(launch_varbase(ctx, 1, P, IDENT, r_c, nullptr, 2, R, nullptr, scratch, 2 * m, 1));
        if (!fan) CKQ(launch_fixedbase(ctx, QQ_BASE_B, v_c, F, m, 1));
        if (side) {
            CK(cudaEventRecord(ctx->msm_ev[9], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_in, ctx->msm_ev[9], 0));
        }
        CKQ(launch_finish_dbl(ctx, dc, fsrc(R, imap(1, 2, 0)), FNONE, FNONE, out + base * 64, imap(1, 2, 0), status + base, 1, m));
        CKQ(on_stream(ctx, side ? ctx->copy_in : ctx->stream, [&] {
            return launch_finish_dbl(ctx, dc2, fsrc(R, imap(1, 2, 1)), fsrc(F, IDENT), FNONE, out + base * 64, imap(1, 2, 1), status + base, 1, m);
        }));
        if (side) {
            CK(cudaEventRecord(ctx->msm_ev[10], ctx->copy_in));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->msm_ev[10], 0));
        }
    }
    return QQ_OK;
}

static int core_add_commitments(qq_ctx* ctx, const uint8_t* a, const uint8_t* b, int negate_b, uint8_t* out,
                                uint8_t* status, size_t n) {
    for (size_t base = 0; base < n; base += QQ_CHUNK) {
        size_t m = n - base < QQ_CHUNK ? n - base : QQ_CHUNK;
        CKQ(ws_begin(ctx, ws_need({2 * m * QQ_PT_BYTES, 2 * m * QQ_PT_BYTES, 2 * m, 2 * m, m, m})));
        u32x4* A = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);
        u32x4* B = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);
        uint8_t* oka = ws_take<uint8_t>(ctx, 2 * m);
        uint8_t* okb = ws_take<uint8_t>(ctx, 2 * m);
        uint8_t* sa = ws_take<uint8_t>(ctx, m);
        uint8_t* sb = ws_take<uint8_t>(ctx, m);
        CKQ(launch_decompress(ctx, a + base * 64, IDENT, A, oka, 2 * m));
        CKQ(launch_decompress(ctx, b + base * 64, IDENT, B, okb, 2 * m));
        CKQ(launch_status(ctx, nullptr, nullptr, nullptr, oka, 2, sa, m));
        CKQ(launch_status(ctx, nullptr, nullptr, nullptr, okb, 2, sb, m));
        k_or_status<<<grid_for(m, 256, ctx->sms * 16), 256, 0, ctx->stream>>>(status + base, sa, sb, m);
        ctx->launches++;
        for (int j = 0; j < 2; j++)
            CKQ(launch_finish(ctx, fsrc(A, imap(1, 2, j)), fsrc(B, imap(1, 2, j), negate_b), FNONE, out + base * 64,
                              imap(1, 2, j), status + base, m));
    }
    return QQ_OK;
}

static int core_update_account(qq_ctx* ctx, const uint8_t* acc, const uint8_t* bl, const uint8_t* u, const uint8_t* c,
                               uint8_t* out, uint8_t* status, size_t n) {
    for (size_t base = 0; base < n; base += QQ_CHUNK) {
        size_t m = n - base < QQ_CHUNK ? n - base : QQ_CHUNK;
        CKQ(ws_begin(ctx, ws_need({4 * m * QQ_PT_BYTES, 2 * m * QQ_PT_BYTES, 2 * m * QQ_PT_BYTES, m * QQ_PT_BYTES, 4 * m, vb_scratch_bytes(ctx, 2)}) + dc_scratch_bytes(2 * m)));
        u32x4* P = ws_take<u32x4>(ctx, 4 * m * QQ_PT_BYTES);   // gr, grsk, c, d of every account
        u32x4* Ru = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);  // u*gr, u*grsk
        u32x4* Rc = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);  // c*gr, c*grsk
        u32x4* F = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);       // bl*B
        uint8_t* ok = ws_take<uint8_t>(ctx, 4 * m);
        u32x4* scratch = ws_take<u32x4>(ctx, vb_scratch_bytes(ctx, 2));
        dc_ws dc = dc_take(ctx, 2 * m);
        const uint8_t* acc_c = acc + base * 128;
        const uint8_t *bl_c = bl + base * 32, *u_c = u + base * 32, *c_c = c + base * 32;
        uint8_t* out_c = out + base * 128;
        // Small batches (a 9-account anonymity set, a block's worth of transactions) are chains of latency-bound launches: the
        // fixed-base walk runs beside the decompression and the scalar multiplications, and the three encoders (each ends in
        // one inverse-square-root or inversion chain, ~60 us) run side by side on the two copy streams.
        const bool fan = small_fanout(ctx, m);
        if (fan) {
            CK(cudaEventRecord(ctx->msm_ev[8], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_in, ctx->msm_ev[8], 0));
            CK(cudaStreamWaitEvent(ctx->copy_out, ctx->msm_ev[8], 0));
            CKQ(on_stream(ctx, ctx->copy_in, [&] { return launch_fixedbase(ctx, QQ_BASE_B, bl_c, F, m); }));
        }
        CKQ(launch_decompress(ctx, acc_c, IDENT, P, ok, 4 * m));
        CKQ(launch_status(ctx, bl_c, u_c, c_c, ok, 4, status + base, m));
        // items: (account i, point j in {gr, grsk}); two scalars (u_i, c_i) share one window table
        // u is halved: Ru = (u/2) * (gr, grsk) feeds the double-and-compress encoder; c stays whole because the
        // commitment adds the account's own points, which cannot be halved
        CKQ(launch_varbase(ctx, 2, P, imap(2, 4, 0, 1), u_c, c_c, 2, Ru, Rc, scratch, 2 * m, 1, 0));
        if (!fan) CKQ(launch_fixedbase(ctx, QQ_BASE_B, bl_c, F, m));
        if (fan) {
            CK(cudaEventRecord(ctx->msm_ev[9], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_in, ctx->msm_ev[9], 0));
            CK(cudaStreamWaitEvent(ctx->copy_out, ctx->msm_ev[9], 0));
        }
        // pk' = (u*gr, u*grsk).  Large batches: the batch encoder's inversion tree is a chain of one-block launches (0.4 ms per
        // 2^18 accounts with the GPU idle); it runs on the high-priority stream while the two classic encoders below fill the SMs.
        const bool side_big = !fan && ctx->small_fanout && ctx->msm_hi != nullptr && m >= 4096;
        if (side_big) {
            CK(cudaEventRecord(ctx->msm_ev[9], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->msm_hi, ctx->msm_ev[9], 0));
        }
        CKQ(on_stream(ctx, side_big ? ctx->msm_hi : ctx->stream, [&] {
            return launch_finish_dbl(ctx, dc, fsrc(Ru, IDENT), FNONE, FNONE, out_c, imap(2, 4, 0, 1), status + base, 2, 2 * m);
        }));
        if (side_big) CK(cudaEventRecord(ctx->msm_ev[10], ctx->msm_hi));
        // comm' = (c*gr + a.c, bl*B + c*grsk + a.d)        -- OLD pk, reference src/accounts/accounts.rs:149-152
        CKQ(on_stream(ctx, fan ? ctx->copy_out : ctx->stream, [&] {
            return launch_finish(ctx, fsrc(Rc, imap(1, 2, 0)), fsrc(P, imap(1, 4, 2)), FNONE, out_c, imap(1, 4, 2), status + base, m);
        }));
        CKQ(on_stream(ctx, fan ? ctx->copy_in : ctx->stream, [&] {
            return launch_finish(ctx, fsrc(Rc, imap(1, 2, 1)), fsrc(F, IDENT), fsrc(P, imap(1, 4, 3)), out_c, imap(1, 4, 3), status + base, m);
        }));
        if (fan) {
            CK(cudaEventRecord(ctx->msm_ev[10], ctx->copy_in));
            CK(cudaEventRecord(ctx->msm_ev[11], ctx->copy_out));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->msm_ev[10], 0));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->msm_ev[11], 0));
        }
        if (side_big) CK(cudaStreamWaitEvent(ctx->stream, ctx->msm_ev[10], 0));
    }
    return QQ_OK;
}

static int core_verify_account(qq_ctx* ctx, const uint8_t* acc, const uint8_t* sk, const uint8_t* bl, uint8_t* status,
                               size_t n) {
    for (size_t base = 0; base < n; base += QQ_CHUNK) {
        size_t m = n - base < QQ_CHUNK ? n - base : QQ_CHUNK;
        CKQ(ws_begin(ctx, ws_need({2 * m * QQ_PT_BYTES, 2 * m * QQ_PT_BYTES, m * QQ_PT_BYTES, 2 * m, 2 * m, m, vb_scratch_bytes(ctx, 1)}) + 2 * dc_scratch_bytes(m)));
        u32x4* P = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);  // gr, c
        u32x4* R = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);  // sk*gr, sk*c
        u32x4* F = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);      // bl*B
        uint8_t* ok = ws_take<uint8_t>(ctx, 2 * m);
        uint8_t* eq = ws_take<uint8_t>(ctx, 2 * m);
        uint8_t* pre = ws_take<uint8_t>(ctx, m);
        u32x4* scratch = ws_take<u32x4>(ctx, vb_scratch_bytes(ctx, 1));
        dc_ws dc = dc_take(ctx, m);
        dc_ws dc2 = dc_take(ctx, m);      // the second encoder's own inversion tree (they run side by side)
        const uint8_t* acc_c = acc + base * 128;
        const uint8_t *sk_c = sk + base * 32, *bl_c = bl + base * 32;
        const bool fan = small_fanout(ctx, m);
        const bool side = ctx->small_fanout && ctx->copy_in != nullptr;      // large batches too: the second encoder beside the first
        if (fan) {
            CK(cudaEventRecord(ctx->msm_ev[8], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_in, ctx->msm_ev[8], 0));
            CKQ(on_stream(ctx, ctx->copy_in, [&] { return launch_fixedbase(ctx, QQ_BASE_B, bl_c, F, m, 1); }));
        }
        CKQ(launch_decompress(ctx, acc_c, imap(2, 4, 0, 2), P, ok, 2 * m));
        CKQ(launch_status(ctx, sk_c, bl_c, nullptr, ok, 0, pre, m));
        CKQ(launch_varbase(ctx, 1, P, IDENT, sk_c, nullptr, 2, R, nullptr, scratch, 2 * m, 1));
        if (!fan) CKQ(launch_fixedbase(ctx, QQ_BASE_B, bl_c, F, m, 1));
        if (side) {
            CK(cudaEventRecord(ctx->msm_ev[9], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_in, ctx->msm_ev[9], 0));
        }
        // grsk == enc(sk*gr)                       reference src/ristretto/keys.rs:187-195
        CKQ(launch_finish_dbl(ctx, dc, fsrc(R, imap(1, 2, 0)), FNONE, FNONE, nullptr, IDENT, nullptr, 1, m, acc_c, imap(1, 4, 1), eq));
        // d == enc(bl*B + sk*c)                    reference src/elgamal/elgamal.rs:81-95
        CKQ(on_stream(ctx, side ? ctx->copy_in : ctx->stream, [&] {
            return launch_finish_dbl(ctx, dc2, fsrc(R, imap(1, 2, 1)), fsrc(F, IDENT), FNONE, nullptr, IDENT, nullptr, 1, m, acc_c, imap(1, 4, 3), eq + m);
        }));
        if (side) {
            CK(cudaEventRecord(ctx->msm_ev[10], ctx->copy_in));
            CK(cudaStreamWaitEvent(ctx->stream, ctx->msm_ev[10], 0));
        }
        k_verify_account_status<<<grid_for(m, 256, ctx->sms * 16), 256, 0, ctx->stream>>>(pre, ok, eq, status + base, m);
        ctx->launches++;
    }
    return QQ_OK;
}

static int core_verify_pk_update(qq_ctx* ctx, const uint8_t* upd, const uint8_t* pk, const uint8_t* r, uint8_t* status,
                                 size_t n) {
    for (size_t base = 0; base < n; base += QQ_CHUNK) {
        size_t m = n - base < QQ_CHUNK ? n - base : QQ_CHUNK;
        CKQ(ws_begin(ctx, ws_need({2 * m * QQ_PT_BYTES, 2 * m * QQ_PT_BYTES, 2 * m * QQ_PT_BYTES, 4 * m, 2 * m, m, m, vb_scratch_bytes(ctx, 1)})));
        u32x4* P = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);
        u32x4* U = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);
        u32x4* R = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);
        uint8_t* ok = ws_take<uint8_t>(ctx, 4 * m);
        uint8_t* eq = ws_take<uint8_t>(ctx, 2 * m);
        uint8_t* s1 = ws_take<uint8_t>(ctx, m);
        uint8_t* s2 = ws_take<uint8_t>(ctx, m);
        u32x4* scratch = ws_take<u32x4>(ctx, vb_scratch_bytes(ctx, 1));
        CKQ(launch_decompress(ctx, pk + base * 64, IDENT, P, ok, 2 * m));
        CKQ(launch_decompress(ctx, upd + base * 64, IDENT, U, ok + 2 * m, 2 * m));
        CKQ(launch_status(ctx, r + base * 32, nullptr, nullptr, ok, 2, s1, m));
        CKQ(launch_status(ctx, nullptr, nullptr, nullptr, ok + 2 * m, 2, s2, m));
        k_or_status<<<grid_for(m, 256, ctx->sms * 16), 256, 0, ctx->stream>>>(s1, s1, s2, m);
        CKQ(launch_varbase(ctx, 1, P, IDENT, r + base * 32, nullptr, 2, R, nullptr, scratch, 2 * m));
        span_begin(ctx, FAM_FIN);
        k_points_equal<<<grid_for(2 * m, 256, ctx->sms * 64), 256, 0, ctx->stream>>>(R, U, eq, 2 * m);
        span_end(ctx);
        k_pair_status<<<grid_for(m, 256, ctx->sms * 16), 256, 0, ctx->stream>>>(s1, eq, QQ_ST_KEYPAIR, status + base, m);
        ctx->launches += 3;
    }
    return QQ_OK;
}

static int core_delta_epsilon(qq_ctx* ctx, const uint8_t* acc, const uint8_t* bl, const uint8_t* r, uint8_t* delta,
                              uint8_t* eps, uint8_t* status, size_t n) {
    for (size_t base = 0; base < n; base += QQ_CHUNK) {
        size_t m = n - base < QQ_CHUNK ? n - base : QQ_CHUNK;
        CKQ(ws_begin(ctx, ws_need({2 * m * QQ_PT_BYTES, 2 * m * QQ_PT_BYTES, m * QQ_PT_BYTES, m * QQ_PT_BYTES, m * QQ_PT_BYTES, 2 * m, vb_scratch_bytes(ctx, 1)}) + dc_scratch_bytes(m)));
        u32x4* P = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);   // gr, grsk
        u32x4* R = ws_take<u32x4>(ctx, 2 * m * QQ_PT_BYTES);   // r*gr, r*grsk
        u32x4* FB = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);      // bl*B
        u32x4* RB = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);      // r*B
        u32x4* RH = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);      // r*H
        uint8_t* ok = ws_take<uint8_t>(ctx, 2 * m);
        u32x4* scratch = ws_take<u32x4>(ctx, vb_scratch_bytes(ctx, 1));
        dc_ws dc = dc_take(ctx, m);
        const uint8_t* acc_c = acc + base * 128;
        const uint8_t *bl_c = bl + base * 32, *r_c = r + base * 32;
        uint8_t *d_c = delta + base * 128, *e_c = eps + base * 128;
        CKQ(launch_decompress(ctx, acc_c, imap(2, 4, 0, 1), P, ok, 2 * m));
        CKQ(launch_status(ctx, r_c, bl_c, nullptr, ok, 2, status + base, m));
        // all four commitment points are sums of scalar multiples: halve every scalar, encode 2 * sum
        CKQ(launch_varbase(ctx, 1, P, IDENT, r_c, nullptr, 2, R, nullptr, scratch, 2 * m, 1));
        CKQ(launch_fixedbase(ctx, QQ_BASE_B, bl_c, FB, m, 1));
        CKQ(launch_fixedbase(ctx, QQ_BASE_B, r_c, RB, m, 1));
        CKQ(launch_fixedbase(ctx, QQ_BASE_H, r_c, RH, m, 1));
        // pk halves: delta keeps the account pk, epsilon carries base_pk (zeroed when the element is bad)
        pk_bytes bpk;
        memcpy(&bpk, ctx->base_pk, 64);
        k_copy_pk<<<grid_for(m, 256, ctx->sms * 16), 256, 0, ctx->stream>>>((const u32x4*)acc_c, (u32x4*)d_c, (u32x4*)e_c,
                                                                            bpk, status + base, m);
        ctx->launches++;
        CKQ(launch_finish_dbl(ctx, dc, fsrc(R, imap(1, 2, 0)), FNONE, FNONE, d_c, imap(1, 4, 2), status + base, 1, m));
        CKQ(launch_finish_dbl(ctx, dc, fsrc(R, imap(1, 2, 1)), fsrc(FB, IDENT), FNONE, d_c, imap(1, 4, 3), status + base, 1, m));
        CKQ(launch_finish_dbl(ctx, dc, fsrc(RB, IDENT), FNONE, FNONE, e_c, imap(1, 4, 2), status + base, 1, m));
        CKQ(launch_finish_dbl(ctx, dc, fsrc(RH, IDENT), fsrc(FB, IDENT), FNONE, e_c, imap(1, 4, 3), status + base, 1, m));
    }
    return QQ_OK;
}

static int core_fixed_base(qq_ctx* ctx, int which, const uint8_t* s, uint8_t* out, uint8_t* status, size_t n) {
    // larger chunks than the account paths: a fixed-base mult is ~100x cheaper than an account update, so the
    // latency-bound tail of the batch inversion is amortised over 4x more elements (workspace: 400 B per element)
    const size_t CH = (size_t)1 << 22;
    for (size_t base = 0; base < n; base += CH) {
        size_t m = n - base < CH ? n - base : CH;
        CKQ(ws_begin(ctx, ws_need({m * QQ_PT_BYTES}) + dc_scratch_bytes(m)));
        u32x4* F = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);
        dc_ws dc = dc_take(ctx, m);
        CKQ(launch_status(ctx, s + base * 32, nullptr, nullptr, nullptr, 0, status + base, m));
        if (!ctx->secret_mode && ctx->fbt[which] != nullptr && m >= QQ_FBT_MIN_BATCH) {
            // table walk fused with the first stage of the batch encoder: the extended point never goes to HBM
            span_begin(ctx, FAM_FB);
            k_fixedbase_big_dc<<<grid_for(m, QQ_FBT_BLOCK, ctx->sms * 4), QQ_FBT_BLOCK, 0, ctx->stream>>>(
                ctx->fbt[which], ctx->fbt_g[which], (const u32x4*)(s + base * 32), dc.state, dc.w, dc.zflag, m);
            span_end(ctx);
            ctx->launches++;
            span_begin(ctx, FAM_FIN);
            CKQ(launch_batch_invert(ctx, dc, m));
            k_dc_finish<<<grid_for(m, 256, ctx->sms * 64), 256, 0, ctx->stream>>>(dc.state, dc.prefix, dc.zflag, status + base, 1,
                                                                                (u32x4*)(out + base * 32), IDENT, nullptr, IDENT,
                                                                                nullptr, m);
            span_end(ctx);
            ctx->launches++;
            CK(cudaGetLastError());
            continue;
        }
        CKQ(launch_fixedbase(ctx, which, s + base * 32, F, m, 1));
        CKQ(launch_finish_dbl(ctx, dc, fsrc(F, IDENT), FNONE, FNONE, out + base * 32, IDENT, status + base, 1, m));
    }
    return QQ_OK;
}

// ---- MSM ----------------------------------------------------------------------------------------------------------
// result: extended point (device, 128 B) = sum s_i * P_i ; *dstatus (device byte) = 0 / 1 / 2
struct qq_prepared;
static int core_msm_to_point(qq_ctx* ctx, const uint8_t* scalars, const uint8_t* points, size_t n, u32x4* result,
                             uint8_t* dstatus, const qq_prepared* pre = nullptr);

// group > 1: every `group` consecutive instances are the parts of ONE MSM (its terms split so that more threads share the
// work: an MSM of 9 terms as 5 + 4 costs 252 doublings twice but halves the chain of additions); out / status then hold
// m / group entries, the parts being summed by the batch encoder (it adds up to three sources).  group <= 3.
// max_terms: the largest instance when the caller knows it (0: unknown) - sizes the cooperative kernel's shared memory
static int core_segmented(qq_ctx* ctx, const uint8_t* scalars, const uint8_t* points, const uint32_t* offsets, size_t m,
                          size_t nterms, uint8_t* out, uint8_t* status, int group = 1, int max_terms = 0) {
    const size_t mo = m / (size_t)group;      // outputs
    auto finish = [&](const dc_ws& dc, u32x4* half, uint8_t* part_status) -> int {
        if (group == 1) return launch_finish_dbl(ctx, dc, fsrc(half, IDENT), FNONE, FNONE, out, IDENT, status, 1, m);
        k_group_status<<<grid_for(mo, 256, ctx->sms * 8), 256, 0, ctx->stream>>>(part_status, group, status, mo);
        ctx->launches++;
        return launch_finish_dbl(ctx, dc, fsrc(half, imap(1, group, 0)), fsrc(half, imap(1, group, 1)),
                                 group == 3 ? fsrc(half, imap(1, group, 2)) : FNONE, out, IDENT, status, 1, mo);
    };
    // Few instances (one proof's worth): four lanes per instance, no ordering pass, direct encoder -- 4 launches
    if (ctx->vbc_max_jobs != 0 && m <= (size_t)ctx->sms * ctx->stc_per_sm) {
        if (!ctx->stc_ready) {
            CK(cudaFuncSetAttribute(k_straus_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, QQ_STC_SMEM_BYTES));
            ctx->stc_ready = true;
        }
        CKQ(ws_begin(ctx, ws_need({nterms * QQ_PT_BYTES, nterms, nterms, m * QQ_PT_BYTES, m}) + dc_scratch_bytes(m)));
        u32x4* P = ws_take<u32x4>(ctx, nterms * QQ_PT_BYTES);
        uint8_t* ok = ws_take<uint8_t>(ctx, nterms);
        uint8_t* tst = ws_take<uint8_t>(ctx, nterms);
        u32x4* half = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);
        uint8_t* part_status = ws_take<uint8_t>(ctx, m);
        dc_ws dc = dc_take(ctx, m);
        CKQ(launch_decompress(ctx, points, IDENT, P, ok, nterms));
        CKQ(launch_status(ctx, scalars, nullptr, nullptr, ok, 1, tst, nterms));
        straus_args a;
        a.pts = P; a.scalars = (const u32x4*)scalars; a.offsets = offsets; a.term_status = tst;
        a.out = (u32x4*)out; a.half_out = half; a.status = group == 1 ? status : part_status; a.scratch = nullptr; a.order = nullptr; a.m = m;
        a.kc = max_terms > 0 && max_terms < QQ_STC_KC ? max_terms : QQ_STC_KC;
        span_begin(ctx, FAM_VB);
        k_straus_coop<<<(unsigned)((m + 7) / 8), 32, (size_t)8 * a.kc * QQ_STC_TERM_Q * 16, ctx->stream>>>(a);
        span_end(ctx);
        ctx->launches++;
        CK(cudaGetLastError());
        CKQ(finish(dc, half, part_status));
        return QQ_OK;
    }
    // 128-thread blocks, 2 per SM, no barrier: a batch holds only a few instances per thread, the lockstep forms of
    // k_varbase (256 x 1, 512 x 1 with a barrier per instance) measured 0-5 % slower here
    const int sblock = 128;
    int occ = 0;
    // more than one wave of instances at 2 blocks per SM: the 4-blocks-per-SM build (128 registers) keeps twice as many
    // instances resident (QQ_STRAUS_MINB = 2 forces the old choice)
    const bool dense = ctx->straus_minb >= 4 && m > (size_t)ctx->sms * 2 * sblock;
    if (dense) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_straus<128, 4, false>, 128, 0));
    else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_straus<128, 2, false>, 128, 0));
    if (occ < 1) occ = 1;
    size_t grid = (m + sblock - 1) / sblock;
    if (grid > (size_t)ctx->sms * occ) grid = (size_t)ctx->sms * occ;
    size_t scratch_bytes = grid * sblock * (size_t)QQ_STRAUS_KMAX * QQ_STRAUS_TERM_Q * 16;
    CKQ(ws_begin(ctx, ws_need({nterms * QQ_PT_BYTES, nterms, nterms, scratch_bytes, m * 4, m * 4, QQ_ORDER_BINS * 4, QQ_ORDER_BINS * 4,
                               QQ_ORDER_BINS * 4, 4096 * 4, m * QQ_PT_BYTES, m}) + dc_scratch_bytes(m)));
    unsigned int* counts = ws_take<unsigned int>(ctx, m * 4);
    unsigned int* order = ws_take<unsigned int>(ctx, m * 4);
    unsigned int* ohist = ws_take<unsigned int>(ctx, QQ_ORDER_BINS * 4);
    unsigned int* ooff = ws_take<unsigned int>(ctx, QQ_ORDER_BINS * 4);
    unsigned int* ocur = ws_take<unsigned int>(ctx, QQ_ORDER_BINS * 4);
    unsigned int* scan_tmp = ws_take<unsigned int>(ctx, 4096 * 4);
    u32x4* P = ws_take<u32x4>(ctx, nterms * QQ_PT_BYTES);
    uint8_t* ok = ws_take<uint8_t>(ctx, nterms);
    uint8_t* tst = ws_take<uint8_t>(ctx, nterms);
    u32x4* scratch = ws_take<u32x4>(ctx, scratch_bytes);
    u32x4* half = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);
    uint8_t* part_status = ws_take<uint8_t>(ctx, m);
    dc_ws dc = dc_take(ctx, m);
    CKQ(launch_decompress(ctx, points, IDENT, P, ok, nterms));
    CKQ(launch_status(ctx, scalars, nullptr, nullptr, ok, 1, tst, nterms));
    // order the instances by term count so that the 32 instances of a warp have (nearly) equal length
    CK(cudaMemsetAsync(ohist, 0, QQ_ORDER_BINS * 4, ctx->stream));
    CK(cudaMemsetAsync(ocur, 0, QQ_ORDER_BINS * 4, ctx->stream));
    k_seg_counts<<<grid_for(m, 256, ctx->sms * 8), 256, 0, ctx->stream>>>(offsets, m, counts);
    k_msm_order_hist<<<grid_for(m, 256, ctx->sms * 8), 256, 0, ctx->stream>>>(counts, m, ohist);
    launch_scan_exclusive(ohist, ooff, QQ_ORDER_BINS, scan_tmp, ctx->stream);
    k_msm_order_scatter<<<grid_for(m, 256, ctx->sms * 8), 256, 0, ctx->stream>>>(counts, m, ooff, ocur, order);
    ctx->launches += 6;
    straus_args a;
    a.pts = P; a.scalars = (const u32x4*)scalars; a.offsets = offsets; a.term_status = tst;
    // every scalar of an instance is halved, the instance sum is encoded as enc(2 * sum) by the batch encoder
    a.out = (u32x4*)out; a.half_out = half; a.status = group == 1 ? status : part_status; a.scratch = scratch; a.order = order; a.m = m;
    a.kc = QQ_STC_KC;
    span_begin(ctx, FAM_VB);
    if (dense) k_straus<128, 4, false><<<(unsigned)grid, 128, 0, ctx->stream>>>(a);
    else k_straus<128, 2, false><<<(unsigned)grid, 128, 0, ctx->stream>>>(a);
    span_end(ctx);
    ctx->launches++;
    CK(cudaGetLastError());
    CKQ(finish(dc, half, part_status));
    return QQ_OK;
}

// =================================================================================================================
// exported entry points: _dev variants call the cores directly, host variants stage through the workspace tail
// =================================================================================================================
struct stage {
    // Device staging for the host-pointer entry points, carved from a grow-only slab owned by the ctx.
    // plan() must be called once with every buffer size before in()/outbuf().
    qq_ctx* ctx;
    size_t off = 0;
    explicit stage(qq_ctx* c) : ctx(c) {}
    int plan(std::initializer_list<size_t> sizes) {
        size_t total = 0;
        for (size_t b : sizes) total += align_up(b ? b : 16, 256);
        if (total > ctx->io_cap) {
            CK(cudaStreamSynchronize(ctx->stream));
            if (ctx->io) CK(cudaFree(ctx->io));
            ctx->io = nullptr;
            ctx->io_cap = 0;
            CK(cudaMalloc((void**)&ctx->io, total));
            ctx->io_cap = total;
        }
        off = 0;
        return QQ_OK;
    }
    uint8_t* take(size_t bytes) {
        uint8_t* p = (uint8_t*)ctx->io + off;
        off += align_up(bytes ? bytes : 16, 256);
        return p;
    }
    int in(const void* host, size_t bytes, uint8_t** out) {
        if (off + align_up(bytes ? bytes : 16, 256) > ctx->io_cap) CKQ(grow(bytes));
        uint8_t* d = take(bytes);
        if (bytes) CK(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
        *out = d;
        return QQ_OK;
    }
    int outbuf(size_t bytes, uint8_t** out) {
        if (off + align_up(bytes ? bytes : 16, 256) > ctx->io_cap) CKQ(grow(bytes));
        *out = take(bytes);
        return QQ_OK;
    }
    int grow(size_t) {
        ctx->err = "internal: staging plan too small";
        return QQ_ERR_ARG;
    }
    int back(void* host, const void* dev, size_t bytes) {
        if (bytes) CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        return QQ_OK;
    }
};

// No exception may cross the C ABI (the callers are ctypes and Rust): the entry points that allocate with std::vector / new or
// start threads run their body through this guard.
template <class F>
static int qq_guarded(qq_ctx* ctx, F&& body) {
    try {
        return body();
    } catch (const std::bad_alloc&) {
        if (ctx) ctx->err = "out of host memory";
        return QQ_ERR_NOMEM;
    } catch (const std::exception& e) {
        if (ctx) ctx->err = std::string("internal error: ") + e.what();
        return QQ_ERR_INTERNAL;
    } catch (...) {
        if (ctx) ctx->err = "internal error";
        return QQ_ERR_INTERNAL;
    }
}

#define ENTER()                                          \
    if (!ctx) return QQ_ERR_ARG;                         \
    ctx->capture_live = ctx->transcript_capture;         \
    ctx->capture_cap = ctx->transcript_capture_cap;      \
    ctx->transcript_capture = nullptr;                   \
    ctx->transcript_capture_cap = 0;                     \
    CK(cudaSetDevice(ctx->device));                      \
    call_begin(ctx)

extern "C" int qq_update_public_key_batch_dev(qq_ctx* ctx, const uint8_t* pk, const uint8_t* r, uint8_t* out_pk,
                                              uint8_t* status, size_t n) {
    ENTER();
    REQUIRE(aligned16(pk) && aligned16(r) && aligned16(out_pk) && status);
    CKQ(core_scale_pairs(ctx, pk, r, out_pk, status, n));
    return call_end(ctx);
}
extern "C" int qq_update_public_key_batch(qq_ctx* ctx, const uint8_t* pk, const uint8_t* r, uint8_t* out_pk,
                                          uint8_t* status, size_t n) {
    ENTER();
    REQUIRE(pk && r && out_pk && status);
    stage st(ctx);
    CKQ(st.plan({(size_t)(n * 64), (size_t)(n * 32), (size_t)(n * 64), (size_t)(n)}));
    uint8_t *dpk, *dr, *dout, *dst;
    CKQ(st.in(pk, n * 64, &dpk));
    CKQ(st.in(r, n * 32, &dr));
    CKQ(st.outbuf(n * 64, &dout));
    CKQ(st.outbuf(n, &dst));
    CKQ(core_scale_pairs(ctx, dpk, dr, dout, dst, n));
    CKQ(st.back(out_pk, dout, n * 64));
    CKQ(st.back(status, dst, n));
    return call_end(ctx);
}
extern "C" int qq_mul_commitment_batch(qq_ctx* ctx, const uint8_t* comm, const uint8_t* s, uint8_t* out_comm,
                                       uint8_t* status, size_t n) {
    return qq_update_public_key_batch(ctx, comm, s, out_comm, status, n);
}
extern "C" int qq_verify_public_key_update_batch(qq_ctx* ctx, const uint8_t* updated_pk, const uint8_t* pk,
                                                 const uint8_t* r, uint8_t* status, size_t n) {
    ENTER();
    REQUIRE(updated_pk && pk && r && status);
    stage st(ctx);
    CKQ(st.plan({(size_t)(n * 64), (size_t)(n * 64), (size_t)(n * 32), (size_t)(n)}));
    uint8_t *du, *dpk, *dr, *dst;
    CKQ(st.in(updated_pk, n * 64, &du));
    CKQ(st.in(pk, n * 64, &dpk));
    CKQ(st.in(r, n * 32, &dr));
    CKQ(st.outbuf(n, &dst));
    CKQ(core_verify_pk_update(ctx, du, dpk, dr, dst, n));
    CKQ(st.back(status, dst, n));
    return call_end(ctx);
}
extern "C" int qq_generate_commitment_batch_dev(qq_ctx* ctx, const uint8_t* pk, const uint8_t* r, const uint8_t* v,
                                                uint8_t* out_comm, uint8_t* status, size_t n) {
    ENTER();
    REQUIRE(aligned16(pk) && aligned16(r) && aligned16(v) && aligned16(out_comm) && status);
    CKQ(core_generate_commitment(ctx, pk, r, v, out_comm, status, n));
    return call_end(ctx);
}
extern "C" int qq_generate_commitment_batch(qq_ctx* ctx, const uint8_t* pk, const uint8_t* r, const uint8_t* v,
                                            uint8_t* out_comm, uint8_t* status, size_t n) {
    ENTER();
    REQUIRE(pk && r && v && out_comm && status);
    stage st(ctx);
    CKQ(st.plan({(size_t)(n * 64), (size_t)(n * 32), (size_t)(n * 32), (size_t)(n * 64), (size_t)(n)}));
    uint8_t *dpk, *dr, *dv, *dout, *dst;
    CKQ(st.in(pk, n * 64, &dpk));
    CKQ(st.in(r, n * 32, &dr));
    CKQ(st.in(v, n * 32, &dv));
    CKQ(st.outbuf(n * 64, &dout));
    CKQ(st.outbuf(n, &dst));
    CKQ(core_generate_commitment(ctx, dpk, dr, dv, dout, dst, n));
    CKQ(st.back(out_comm, dout, n * 64));
    CKQ(st.back(status, dst, n));
    return call_end(ctx);
}
extern "C" int qq_add_commitments_batch(qq_ctx* ctx, const uint8_t* a, const uint8_t* b, int negate_b,
                                        uint8_t* out_comm, uint8_t* status, size_t n) {
    ENTER();
    REQUIRE(a && b && out_comm && status);
    stage st(ctx);
    CKQ(st.plan({(size_t)(n * 64), (size_t)(n * 64), (size_t)(n * 64), (size_t)(n)}));
    uint8_t *da, *db, *dout, *dst;
    CKQ(st.in(a, n * 64, &da));
    CKQ(st.in(b, n * 64, &db));
    CKQ(st.outbuf(n * 64, &dout));
    CKQ(st.outbuf(n, &dst));
    CKQ(core_add_commitments(ctx, da, db, negate_b ? 1 : 0, dout, dst, n));
    CKQ(st.back(out_comm, dout, n * 64));
    CKQ(st.back(status, dst, n));
    return call_end(ctx);
}
extern "C" int qq_update_account_batch_dev(qq_ctx* ctx, const uint8_t* acc, const uint8_t* bl, const uint8_t* u,
                                           const uint8_t* c, uint8_t* out_acc, uint8_t* status, size_t n) {
    ENTER();
    REQUIRE(aligned16(acc) && aligned16(bl) && aligned16(u) && aligned16(c) && aligned16(out_acc) && status);
    CKQ(core_update_account(ctx, acc, bl, u, c, out_acc, status, n));
    return call_end(ctx);
}
// Host-pointer entry of the headline path.  The batch is cut into slices; the upload of slice i + 1 (copy stream) and the
// download of slice i - 1 (second copy stream) run under the kernels of slice i, so that with pinned host memory only the
// first upload and the last download are exposed (352 bytes per account cross PCIe).  Those two are kept short - the first
// slice is two waves of the dominant kernel (k_varbase_split: one 512-thread block per SM, two scalar-mult jobs per account
// = 256 accounts per SM and wave; 6 ms of kernels, enough to cover the upload of everything behind it up to 2^20 accounts),
// the last slice one wave - and the slices between them are as large as the chunking allows: every slice pays the latency
// of its own batch inversions (k_binv_*, ~0.3 ms), so few slices, each a whole number of waves.
static void pipe_slices(const qq_ctx* ctx, size_t n, std::vector<size_t>& cuts) {
    const size_t wave = (size_t)ctx->sms * 256;
    cuts.clear();
    cuts.push_back(0);
    if (n <= 4 * wave) {
        cuts.push_back(n);
        return;
    }
    size_t lo = 2 * wave;
    cuts.push_back(lo);
    const size_t big = (((size_t)1 << 20) / wave) * wave;
    while (n - lo > big + wave) {
        lo += big;
        cuts.push_back(lo);
    }
    cuts.push_back(n - wave);
    cuts.push_back(n);
}
extern "C" int qq_update_account_batch(qq_ctx* ctx, const uint8_t* acc, const uint8_t* bl, const uint8_t* u,
                                       const uint8_t* c, uint8_t* out_acc, uint8_t* status, size_t n) {
    return qq_guarded(ctx, [&]() -> int {
    ENTER();
    REQUIRE(acc && bl && u && c && out_acc && status);
    stage st(ctx);
    CKQ(st.plan({(size_t)(n * 128), (size_t)(n * 32), (size_t)(n * 32), (size_t)(n * 32), (size_t)(n * 128), (size_t)(n)}));
    uint8_t* dacc = st.take(n * 128);
    uint8_t* dbl = st.take(n * 32);
    uint8_t* du = st.take(n * 32);
    uint8_t* dc = st.take(n * 32);
    uint8_t* dout = st.take(n * 128);
    uint8_t* dst = st.take(n);
    std::vector<size_t> cuts;
    pipe_slices(ctx, n, cuts);
    const size_t nsl = cuts.size() - 1;
    while (ctx->pipe_ev.size() < 2 * nsl) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->pipe_ev.push_back(e);
    }
    // the staging slab may still be read by the previous call's downloads: order the copy streams after the main stream
    CK(cudaEventRecord(ctx->msm_ev[7], ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_in, ctx->msm_ev[7], 0));
    for (size_t i = 0; i < nsl; i++) {
        const size_t lo = cuts[i], m = cuts[i + 1] - lo;
        CK(cudaMemcpyAsync(dacc + lo * 128, acc + lo * 128, m * 128, cudaMemcpyHostToDevice, ctx->copy_in));
        CK(cudaMemcpyAsync(dbl + lo * 32, bl + lo * 32, m * 32, cudaMemcpyHostToDevice, ctx->copy_in));
        CK(cudaMemcpyAsync(du + lo * 32, u + lo * 32, m * 32, cudaMemcpyHostToDevice, ctx->copy_in));
        CK(cudaMemcpyAsync(dc + lo * 32, c + lo * 32, m * 32, cudaMemcpyHostToDevice, ctx->copy_in));
        CK(cudaEventRecord(ctx->pipe_ev[2 * i], ctx->copy_in));
    }
    for (size_t i = 0; i < nsl; i++) {
        const size_t lo = cuts[i], m = cuts[i + 1] - lo;
        CK(cudaStreamWaitEvent(ctx->stream, ctx->pipe_ev[2 * i], 0));
        CKQ(core_update_account(ctx, dacc + lo * 128, dbl + lo * 32, du + lo * 32, dc + lo * 32, dout + lo * 128, dst + lo, m));
        CK(cudaEventRecord(ctx->pipe_ev[2 * i + 1], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->copy_out, ctx->pipe_ev[2 * i + 1], 0));
        CK(cudaMemcpyAsync(out_acc + lo * 128, dout + lo * 128, m * 128, cudaMemcpyDeviceToHost, ctx->copy_out));
        CK(cudaMemcpyAsync(status + lo, dst + lo, m, cudaMemcpyDeviceToHost, ctx->copy_out));
    }
    CK(cudaEventRecord(ctx->msm_ev[6], ctx->copy_out));
    CK(cudaStreamWaitEvent(ctx->stream, ctx->msm_ev[6], 0));   // call_end synchronises the main stream
    return call_end(ctx);
    });
}
extern "C" int qq_verify_account_batch_dev(qq_ctx* ctx, const uint8_t* acc, const uint8_t* sk, const uint8_t* bl,
                                           uint8_t* status, size_t n) {
    ENTER();
    REQUIRE(aligned16(acc) && aligned16(sk) && aligned16(bl) && status);
    CKQ(core_verify_account(ctx, acc, sk, bl, status, n));
    return call_end(ctx);
}
extern "C" int qq_verify_account_batch(qq_ctx* ctx, const uint8_t* acc, const uint8_t* sk, const uint8_t* bl,
                                       uint8_t* status, size_t n) {
    ENTER();
    REQUIRE(acc && sk && bl && status);
    stage st(ctx);
    CKQ(st.plan({(size_t)(n * 128), (size_t)(n * 32), (size_t)(n * 32), (size_t)(n)}));
    uint8_t *dacc, *dsk, *dbl, *dst;
    CKQ(st.in(acc, n * 128, &dacc));
    CKQ(st.in(sk, n * 32, &dsk));
    CKQ(st.in(bl, n * 32, &dbl));
    CKQ(st.outbuf(n, &dst));
    CKQ(core_verify_account(ctx, dacc, dsk, dbl, dst, n));
    CKQ(st.back(status, dst, n));
    return call_end(ctx);
}
extern "C" int qq_delta_epsilon_batch(qq_ctx* ctx, const uint8_t* acc, const uint8_t* bl, const uint8_t* r,
                                      const uint8_t* base_pk, uint8_t* out_delta, uint8_t* out_epsilon,
                                      uint8_t* status, size_t n) {
    ENTER();
    REQUIRE(acc && bl && r && base_pk && out_delta && out_epsilon && status);
    if (memcmp(base_pk, ctx->base_pk, 64) != 0) {
        ctx->err = "qq_delta_epsilon_batch: base_pk must be BASE_PK_BTC_COMPRESSED (B, H)";
        return QQ_ERR_ARG;
    }
    stage st(ctx);
    CKQ(st.plan({(size_t)(n * 128), (size_t)(n * 32), (size_t)(n * 32), (size_t)(n * 128), (size_t)(n * 128), (size_t)(n)}));
    uint8_t *dacc, *dbl, *dr, *dd, *de, *dst;
    CKQ(st.in(acc, n * 128, &dacc));
    CKQ(st.in(bl, n * 32, &dbl));
    CKQ(st.in(r, n * 32, &dr));
    CKQ(st.outbuf(n * 128, &dd));
    CKQ(st.outbuf(n * 128, &de));
    CKQ(st.outbuf(n, &dst));
    CKQ(core_delta_epsilon(ctx, dacc, dbl, dr, dd, de, dst, n));
    CKQ(st.back(out_delta, dd, n * 128));
    CKQ(st.back(out_epsilon, de, n * 128));
    CKQ(st.back(status, dst, n));
    return call_end(ctx);
}
extern "C" int qq_fixed_base_batch_dev(qq_ctx* ctx, int which, const uint8_t* s, uint8_t* out_points, uint8_t* status,
                                       size_t n) {
    ENTER();
    REQUIRE((which == 0 || which == 1) && aligned16(s) && aligned16(out_points) && status);
    CKQ(core_fixed_base(ctx, which, s, out_points, status, n));
    return call_end(ctx);
}
extern "C" int qq_fixed_base_batch(qq_ctx* ctx, int which, const uint8_t* s, uint8_t* out_points, uint8_t* status,
                                   size_t n) {
    ENTER();
    REQUIRE((which == 0 || which == 1) && s && out_points && status);
    stage st(ctx);
    CKQ(st.plan({(size_t)(n * 32), (size_t)(n * 32), (size_t)(n)}));
    uint8_t *ds, *dout, *dst;
    CKQ(st.in(s, n * 32, &ds));
    CKQ(st.outbuf(n * 32, &dout));
    CKQ(st.outbuf(n, &dst));
    CKQ(core_fixed_base(ctx, which, ds, dout, dst, n));
    CKQ(st.back(out_points, dout, n * 32));
    CKQ(st.back(status, dst, n));
    return call_end(ctx);
}

// 64-bit signed values (balances): out_i = enc(v_i * Base).  Needs a large-window table (qq_fixed_base_set_window != 0).
static int core_fixed_base_i64(qq_ctx* ctx, int which, const int64_t* v, uint8_t* out, size_t n) {
    if (ctx->secret_mode) {
        // balances are secrets on the wallet side: expanded to canonical scalars and sent through the masked-scan table walk
        const size_t CH = (size_t)1 << 22;
        for (size_t base = 0; base < n; base += CH) {
            size_t m = n - base < CH ? n - base : CH;
            CKQ(ws_begin(ctx, ws_need({m * QQ_PT_BYTES, m * 32, m}) + dc_scratch_bytes(m)));
            u32x4* F = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);
            uint8_t* sc = ws_take<uint8_t>(ctx, m * 32);
            uint8_t* st = ws_take<uint8_t>(ctx, m);
            dc_ws dc = dc_take(ctx, m);
            k_i64_to_scalars<<<grid_for(m, 256, ctx->sms * 8), 256, 0, ctx->stream>>>((const long long*)(v + base), (u32x4*)sc, m);
            ctx->launches++;
            CK(cudaMemsetAsync(st, 0, m, ctx->stream));
            CKQ(launch_fixedbase(ctx, which, sc, F, m, 1));
            CKQ(launch_finish_dbl(ctx, dc, fsrc(F, IDENT), FNONE, FNONE, out + base * 32, IDENT, st, 1, m));
        }
        return QQ_OK;
    }
    if (ctx->fbt[which] == nullptr) {
        ctx->err = "qq_fixed_base_i64_batch needs a large-window table (qq_fixed_base_set_window)";
        return QQ_ERR_ARG;
    }
    fbt_geom g = ctx->fbt_g[which];
    g.NW = (66 + g.W - 1) / g.W;                      // windows covering a 63-bit half magnitude + recoding bias
    const size_t CH = (size_t)1 << 22;
    for (size_t base = 0; base < n; base += CH) {
        size_t m = n - base < CH ? n - base : CH;
        CKQ(ws_begin(ctx, dc_scratch_bytes(m)));
        dc_ws dc = dc_take(ctx, m);
        span_begin(ctx, FAM_FB);
        k_fixedbase_big_i64_dc<<<grid_for(m, QQ_FBT_BLOCK, ctx->sms * 4), QQ_FBT_BLOCK, 0, ctx->stream>>>(
            ctx->fbt[which], g, (const long long*)(v + base), ctx->half_base[which], dc.state, dc.w, dc.zflag, m);
        span_end(ctx);
        ctx->launches++;
        span_begin(ctx, FAM_FIN);
        CKQ(launch_batch_invert(ctx, dc, m));
        k_dc_finish<<<grid_for(m, 256, ctx->sms * 64), 256, 0, ctx->stream>>>(dc.state, dc.prefix, dc.zflag, nullptr, 1,
                                                                            (u32x4*)(out + base * 32), IDENT, nullptr, IDENT, nullptr, m);
        span_end(ctx);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return QQ_OK;
}
extern "C" int qq_fixed_base_i64_batch_dev(qq_ctx* ctx, int which, const int64_t* v, uint8_t* out_points, size_t n) {
    ENTER();
    REQUIRE((which == 0 || which == 1) && (((uintptr_t)v & 7) == 0) && aligned16(out_points));
    CKQ(core_fixed_base_i64(ctx, which, v, out_points, n));
    return call_end(ctx);
}
extern "C" int qq_fixed_base_i64_batch(qq_ctx* ctx, int which, const int64_t* v, uint8_t* out_points, size_t n) {
    ENTER();
    REQUIRE((which == 0 || which == 1) && (n == 0 || (v && out_points)));
    stage st(ctx);
    CKQ(st.plan({(size_t)(n * 8), (size_t)(n * 32)}));
    uint8_t *dv, *dout;
    CKQ(st.in(v, n * 8, &dv));
    CKQ(st.outbuf(n * 32, &dout));
    CKQ(core_fixed_base_i64(ctx, which, (const int64_t*)dv, dout, n));
    CKQ(st.back(out_points, dout, n * 32));
    return call_end(ctx);
}

// Prover-side sigma-protocol commitments (reference src/accounts/prover.rs:164-207, 415-454, 629-640,
// 742-753, 885-921): every e / f the provers compute is  r_i * P_i  (P_i = a key or commitment component of an account) or
// v_i * B + r_i * P_i.  out_i = enc(r_i * dec(P_i) [+ v_i * B]); v == nullptr: no fixed-base term.  The blindings are
// secrets: callers switch qq_set_secret_mode on.
static int core_sigma_commit(qq_ctx* ctx, const uint8_t* points, const uint8_t* r, const uint8_t* v, uint8_t* out, uint8_t* status,
                             size_t n) {
    for (size_t base = 0; base < n; base += QQ_CHUNK) {
        size_t m = n - base < QQ_CHUNK ? n - base : QQ_CHUNK;
        CKQ(ws_begin(ctx, ws_need({m * QQ_PT_BYTES, m * QQ_PT_BYTES, m * QQ_PT_BYTES, m, vb_scratch_bytes(ctx, 1)}) + dc_scratch_bytes(m)));
        u32x4* P = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);
        u32x4* R = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);
        u32x4* F = ws_take<u32x4>(ctx, m * QQ_PT_BYTES);
        uint8_t* ok = ws_take<uint8_t>(ctx, m);
        u32x4* scratch = ws_take<u32x4>(ctx, vb_scratch_bytes(ctx, 1));
        dc_ws dc = dc_take(ctx, m);
        CKQ(launch_decompress(ctx, points + base * 32, IDENT, P, ok, m));
        CKQ(launch_status(ctx, r + base * 32, v ? v + base * 32 : nullptr, nullptr, ok, 1, status + base, m));
        CKQ(launch_varbase(ctx, 1, P, IDENT, r + base * 32, nullptr, 1, R, nullptr, scratch, m, 1));
        if (v) CKQ(launch_fixedbase(ctx, QQ_BASE_B, v + base * 32, F, m, 1));
        CKQ(launch_finish_dbl(ctx, dc, fsrc(R, IDENT), v ? fsrc(F, IDENT) : FNONE, FNONE, out + base * 32, IDENT, status + base, 1, m));
    }
    return QQ_OK;
}
extern "C" int qq_sigma_commit_batch_dev(qq_ctx* ctx, const uint8_t* points, const uint8_t* r, const uint8_t* v, uint8_t* out_points,
                                         uint8_t* status, size_t n) {
    ENTER();
    REQUIRE(aligned16(points) && aligned16(r) && (v == nullptr || aligned16(v)) && aligned16(out_points) && status);
    CKQ(core_sigma_commit(ctx, points, r, v, out_points, status, n));
    return call_end(ctx);
}
extern "C" int qq_sigma_commit_batch(qq_ctx* ctx, const uint8_t* points, const uint8_t* r, const uint8_t* v, uint8_t* out_points,
                                     uint8_t* status, size_t n) {
    ENTER();
    REQUIRE(points && r && out_points && status);
    stage st(ctx);
    CKQ(st.plan({(size_t)(n * 32), (size_t)(n * 32), (size_t)(n * 32), (size_t)(n * 32), (size_t)n}));
    uint8_t *dp, *dr, *dv = nullptr, *dout, *dst;
    CKQ(st.in(points, n * 32, &dp));
    CKQ(st.in(r, n * 32, &dr));
    if (v) CKQ(st.in(v, n * 32, &dv));
    CKQ(st.outbuf(n * 32, &dout));
    CKQ(st.outbuf(n, &dst));
    CKQ(core_sigma_commit(ctx, dp, dr, dv, dout, dst, n));
    CKQ(st.back(out_points, dout, n * 32));
    CKQ(st.back(status, dst, n));
    return call_end(ctx);
}

#include "qq_api_msm.inc"
#include "qq_api_sigma.inc"
#include "qq_api_shuffle.inc"
#include "qq_api_rangeproof.inc"
#include "qq_api_wire.inc"
#include "qq_api_multi.inc"
