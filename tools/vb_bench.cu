// Tuning harness for the variable-base kernels: block size x register cap x table scheme x per-item barrier.
// Arithmetic is data-independent, so random limbs are fine for timing.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "kernels.cuh"
#ifndef QQ_BUILD_TAG
#define QQ_BUILD_TAG "default"
#endif
using namespace qq;

// MODE 0: one 9-entry table, 252 doublings per scalar.  MODE 1: split tables (vbs_*), 192 + 60 per scalar.
// MODE 2: MODE 1 + one block barrier per item (all threads run the same number of rounds; out-of-range threads redo the
// last item without storing).  MODE 3: MODE 0 + the barrier.  Build flags compared in profiles/: -DQQ_FE_SINGLE (one
// product per out-of-line call), -DQQ_INLINE_FIELD_OPS (everything inlined), -DQQ_FE_MUL_KARATSUBA.
template <int BLOCK, int MINB, int MODE>
__global__ void __launch_bounds__(BLOCK, MINB) k_vb(vb_args a) {
    size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    u32x4* tbl = a.scratch + gtid * (MODE ? QQ_VBS_TABLE_Q : QQ_VB_ENTRIES * QQ_PT_Q);
    size_t rounds = (a.n + stride - 1) / stride;
    for (size_t it = 0; it < rounds; it++) {
        size_t t = gtid + it * stride;
        bool live = t < a.n;
        if (MODE >= 2) {
            __syncthreads();
            if (!live) t = a.n - 1;
        } else if (!live) {
            continue;
        }
        ge_p3 p, r;
        ge_p3_load(p, a.pts + QQ_PT_Q * map_index(a.map, t));
        if (MODE == 1 || MODE == 2) vbs_build_tables(tbl, p);
        else vb_build_table(tbl, p);
        u32 s[8];
        load_words32(s, a.s0, t / (size_t)a.sdiv);
        if (MODE == 1 || MODE == 2) vbs_scalarmult(r, tbl, s);
        else vb_scalarmult(r, tbl, s);
        if (live) ge_p3_store(a.out0 + QQ_PT_Q * t, r);
        load_words32(s, a.s1, t / (size_t)a.sdiv);
        if (MODE == 1 || MODE == 2) vbs_scalarmult(r, tbl, s);
        else vb_scalarmult(r, tbl, s);
        if (live) ge_p3_store(a.out1 + QQ_PT_Q * t, r);
    }
}

template <int BLOCK, int MINB, int MODE>
static void run(size_t n, int sms) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_vb<BLOCK, MINB, MODE>, BLOCK, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_vb<BLOCK, MINB, MODE>);
    int grid = sms * occ;
    u32x4 *pts, *s0, *s1, *o0, *o1, *scratch;
    cudaMalloc(&pts, n * QQ_PT_BYTES); cudaMalloc(&s0, n * 32); cudaMalloc(&s1, n * 32);
    cudaMalloc(&o0, n * QQ_PT_BYTES); cudaMalloc(&o1, n * QQ_PT_BYTES);
    cudaMalloc(&scratch, (size_t)grid * BLOCK * QQ_VBS_TABLE_WORDS * 4);
    size_t words = n * 32;
    u32* h = (u32*)malloc(words * 4);
    for (size_t i = 0; i < words; i++) h[i] = (u32)rand() * 2654435761u;
    cudaMemcpy(pts, h, words * 4, cudaMemcpyHostToDevice);
    for (size_t i = 0; i < n * 8; i++) h[i] = (u32)rand() * 2654435761u;
    for (size_t i = 7; i < n * 8; i += 8) h[i] &= 0x0fffffff;
    cudaMemcpy(s0, h, n * 32, cudaMemcpyHostToDevice);
    cudaMemcpy(s1, h + 8, n * 32 - 32, cudaMemcpyHostToDevice);
    vb_args a;
    a.pts = pts; a.map = idx_map{1, 1, {0, 0, 0, 0}}; a.s0 = s0; a.s1 = s1; a.sdiv = 1; a.halve0 = 0; a.halve1 = 0;
    a.out0 = o0; a.out1 = o1; a.scratch = scratch; a.n = n;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_vb<BLOCK, MINB, MODE><<<grid, BLOCK>>>(a);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 2; r++) {
        cudaEventRecord(e0);
        k_vb<BLOCK, MINB, MODE><<<grid, BLOCK>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("{\"build\": \"" QQ_BUILD_TAG "\", \"variant\": \"block%d_minb%d_%s\", \"regs\": %d, \"local_bytes\": %zu, \"blocks_per_sm\": %d, \"n_items\": %zu, \"ms\": %.3f, \"scalar_mults_per_s\": %.4e, \"err\": \"%s\"}\n",
           BLOCK, MINB, MODE == 3 ? "plain_sync" : (MODE == 2 ? "split4_sync" : (MODE ? "split4" : "plain")), fa.numRegs, (size_t)fa.localSizeBytes, occ, n, best,
           2.0 * n / (best * 1e-3), cudaGetErrorString(cudaGetLastError()));
    cudaFree(pts); cudaFree(s0); cudaFree(s1); cudaFree(o0); cudaFree(o1); cudaFree(scratch); free(h);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    size_t n = (size_t)1 << (getenv("VB_LOG2N") ? atoi(getenv("VB_LOG2N")) : 18);
    int sms = p.multiProcessorCount;
    if (!getenv("VB_SPLIT_ONLY")) {
        run<128, 2, 0>(n, sms);
        run<128, 4, 0>(n, sms);
        run<128, 2, 1>(n, sms);
        run<128, 4, 1>(n, sms);
    }
    run<384, 1, 2>(n, sms);
    run<512, 1, 2>(n, sms);
    run<640, 1, 2>(n, sms);
    run<128, 4, 3>(n, sms);
    run<256, 2, 3>(n, sms);
    run<512, 1, 3>(n, sms);
    return 0;
}
