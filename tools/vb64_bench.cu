// Bake-off for the two-pipe variable-base kernel (k_varbase_split_hybrid): integer warps alone, FP64 warps alone, both.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I quisquis-rust_b200/csrc -I tools -o tools/bin/vb64_bench tools/vb64_bench.cu
// Points are random limbs (the formulas are polynomial identities: both paths must agree mod p on any input), which
// also makes the comparison below a parity check of the FP64 path against the integer path on 2^n items.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "kernels.cuh"
#include "fe64/fe64.cuh"
using namespace qq;

namespace qq {
// Both multiplier pipes at once: NI integer warps run vbs_* (IMAD.WIDE on the FMA-heavy pipe, carries on the ALU pipe)
// and NF warps run vbs64_* (fe64.cuh: nothing but DFMA / DADD / DMUL on the FP64 pipe, which the integer warps leave
// idle).  The two groups draw runs of items from one global counter, so the split follows their measured speeds; every
// group keeps the lockstep barrier of k_varbase_split among its own warps (named barriers 1 and 2).  Outputs are extended
// points in the saturated integer form either way; the encodings downstream are byte-identical.
template <int NI, int NF>
__global__ void __launch_bounds__((NI + NF) * 32, 1) k_varbase_split_hybrid(vb_args a, double* __restrict__ scratch64,
                                                                            unsigned long long* __restrict__ counter) {
    __shared__ unsigned long long next[2][2];
    const int lt = threadIdx.x;
    if (lt < NI * 32) {
        if constexpr (NI > 0) {
            u32x4* tbl = a.scratch + ((size_t)blockIdx.x * (NI * 32) + lt) * QQ_VBS_TABLE_Q;
            for (int round = 0;; round++) {
                if (lt == 0) next[0][round & 1] = atomicAdd(counter, (unsigned long long)(NI * 32));
                asm volatile("bar.sync 1, %0;" ::"n"(NI * 32) : "memory");
                size_t t = (size_t)next[0][round & 1];
                if (t >= a.n) break;
                if (lt == 0) atomicAdd(counter + 1, 1ull);      // runs taken by the integer group (reported by the bench)
                t += lt;
                if (t >= a.n) continue;
                ge_p3 p, r;
                ge_p3_load(p, a.pts + QQ_PT_Q * map_index(a.map, t));
                vbs_build_tables(tbl, p);
                u32 s[8];
                load_words32(s, a.s0, t / (size_t)a.sdiv);
                if (a.halve0) sc_halve(s, s);
                vbs_scalarmult<false>(r, tbl, s);
                ge_p3_store(a.out0 + QQ_PT_Q * t, r);
                load_words32(s, a.s1, t / (size_t)a.sdiv);
                if (a.halve1) sc_halve(s, s);
                vbs_scalarmult<false>(r, tbl, s);
                ge_p3_store(a.out1 + QQ_PT_Q * t, r);
            }
        }
    } else {
        if constexpr (NF > 0) {
            const int lf = lt - NI * 32;
            double* tbl = scratch64 + ((size_t)blockIdx.x * (NF * 32) + lf) * QQ_VBS64_TABLE_D;
            fe64 d2;
            fe64_from_fe(d2, fe_2d());
            for (int round = 0;; round++) {
                if (lf == 0) next[1][round & 1] = atomicAdd(counter, (unsigned long long)(NF * 32));
                asm volatile("bar.sync 2, %0;" ::"n"(NF * 32) : "memory");
                size_t t = (size_t)next[1][round & 1];
                if (t >= a.n) break;
                if (lf == 0) atomicAdd(counter + 2, 1ull);      // runs taken by the FP64 group
                t += lf;
                if (t >= a.n) continue;
                ge_p3 p, r;
                ge_p3_load(p, a.pts + QQ_PT_Q * map_index(a.map, t));
                ge64_p3 p64, r64;
                ge64_from_p3(p64, p);
                vbs64_build_tables(tbl, p64, d2);
                u32 s[8];
                load_words32(s, a.s0, t / (size_t)a.sdiv);
                if (a.halve0) sc_halve(s, s);
                vbs64_scalarmult(r64, tbl, s);
                ge64_to_p3(r, r64);
                ge_p3_store(a.out0 + QQ_PT_Q * t, r);
                load_words32(s, a.s1, t / (size_t)a.sdiv);
                if (a.halve1) sc_halve(s, s);
                vbs64_scalarmult(r64, tbl, s);
                ge64_to_p3(r, r64);
                ge_p3_store(a.out1 + QQ_PT_Q * t, r);
            }
        }
    }
}

}  // namespace qq

static std::vector<u32> g_ref0, g_ref1;

template <int NI, int NF>
static void run(size_t n, int sms, bool is_ref) {
    auto kern = k_varbase_split_hybrid<NI, NF>;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    int grid = sms;
    const int BLOCK = (NI + NF) * 32;
    u32x4 *pts, *s0, *s1, *o0, *o1, *scratch;
    double* scratch64;
    unsigned long long* counter;
    cudaMalloc(&pts, n * QQ_PT_BYTES); cudaMalloc(&s0, n * 32); cudaMalloc(&s1, n * 32);
    cudaMalloc(&o0, n * QQ_PT_BYTES); cudaMalloc(&o1, n * QQ_PT_BYTES);
    cudaMalloc(&scratch, (size_t)grid * (NI ? NI : 1) * 32 * QQ_VBS_TABLE_WORDS * 4);
    cudaMalloc(&scratch64, (size_t)grid * (NF ? NF : 1) * 32 * QQ_VBS64_TABLE_D * 8);
    cudaMalloc(&counter, 64);
    size_t words = n * 32;
    std::vector<u32> h(words);
    srand(12345);
    for (size_t i = 0; i < words; i++) h[i] = (u32)rand() * 2654435761u + (u32)rand();
    cudaMemcpy(pts, h.data(), n * QQ_PT_BYTES, cudaMemcpyHostToDevice);
    for (size_t i = 0; i < n * 8; i++) { h[i] = (u32)rand() * 2246822519u + (u32)rand(); if ((i & 7) == 7) h[i] &= 0x0fffffffu; }
    cudaMemcpy(s0, h.data(), n * 32, cudaMemcpyHostToDevice);
    for (size_t i = 0; i < n * 8; i++) { h[i] = (u32)rand() * 3266489917u + (u32)rand(); if ((i & 7) == 7) h[i] &= 0x0fffffffu; }
    cudaMemcpy(s1, h.data(), n * 32, cudaMemcpyHostToDevice);
    vb_args a;
    a.pts = pts; a.map = {1, 1, {0, 0, 0, 0}}; a.s0 = s0; a.s1 = s1; a.sdiv = 1; a.halve0 = 0; a.halve1 = 1;
    a.out0 = o0; a.out1 = o1; a.scratch = scratch; a.n = n;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    unsigned long long st[3] = {0, 0, 0};
    for (int rep = 0; rep < 3; rep++) {
        cudaMemsetAsync(counter, 0, 64);
        cudaEventRecord(e0);
        kern<<<grid, BLOCK>>>(a, scratch64, counter);
        cudaEventRecord(e1);
        cudaError_t err = cudaEventSynchronize(e1);
        if (err != cudaSuccess) { printf("{\"error\": \"%s\", \"ni\": %d, \"nf\": %d}\n", cudaGetErrorString(err), NI, NF); exit(1); }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        cudaMemcpy(st, counter, 24, cudaMemcpyDeviceToHost);
    }
    // parity against the integer-only run: canonical X, Y, Z, T of both outputs
    std::vector<u32> r0(n * 32), r1(n * 32);
    cudaMemcpy(r0.data(), o0, n * QQ_PT_BYTES, cudaMemcpyDeviceToHost);
    cudaMemcpy(r1.data(), o1, n * QQ_PT_BYTES, cudaMemcpyDeviceToHost);
    for (size_t i = 0; i < n * 4; i++) {
        fe f;
        u32 w[8];
        memcpy(f.v, &r0[i * 8], 32); fe_towords(w, f); memcpy(&r0[i * 8], w, 32);
        memcpy(f.v, &r1[i * 8], 32); fe_towords(w, f); memcpy(&r1[i * 8], w, 32);
    }
    size_t bad = 0;
    if (is_ref) { g_ref0 = r0; g_ref1 = r1; }
    else {
        for (size_t i = 0; i < n * 4; i++)
            if (memcmp(&r0[i * 8], &g_ref0[i * 8], 32) || memcmp(&r1[i * 8], &g_ref1[i * 8], 32)) bad++;
    }
    double items_i = (double)st[1] * NI * 32, items_f = (double)st[2] * NF * 32;
    printf("{\"int_warps\": %d, \"fp64_warps\": %d, \"regs\": %d, \"local_bytes\": %zu, \"n\": %zu, \"ms\": %.3f, "
           "\"scalar_mults_per_s\": %.4e, \"share_fp64\": %.3f, \"mismatching_coordinates\": %zu}\n",
           NI, NF, fa.numRegs, (size_t)fa.localSizeBytes, n, best, 2.0 * n / (best * 1e-3),
           items_f / (items_i + items_f > 0 ? items_i + items_f : 1), bad);
    fflush(stdout);
    cudaFree(pts); cudaFree(s0); cudaFree(s1); cudaFree(o0); cudaFree(o1); cudaFree(scratch); cudaFree(scratch64); cudaFree(counter);
}


// Pipe-interference probe: NI integer warps do the real split scalar multiplications; NF warps spin on a register-only
// chain of FP64 field products (fe64_mul_ool / fe64_sq_ool: 3.5 + 2.5 KB of code, no memory traffic, no spills) until the
// integer group has finished, and report how many products they got through.  Separates what the two PIPES cost each
// other from what the real FP64 scalar multiplication adds on top (register spills, table traffic, instruction cache).
template <int NI, int NF>
__global__ void __launch_bounds__((NI + NF) * 32, 1) k_probe(vb_args a, unsigned long long* __restrict__ counter, double* sink) {
    __shared__ unsigned long long next[2];
    __shared__ volatile int done;
    const int lt = threadIdx.x;
    if (lt == 0) done = 0;
    __syncthreads();
    if (lt < NI * 32) {
        u32x4* tbl = a.scratch + ((size_t)blockIdx.x * (NI * 32) + lt) * QQ_VBS_TABLE_Q;
        for (int round = 0;; round++) {
            if (lt == 0) next[round & 1] = atomicAdd(counter, (unsigned long long)(NI * 32));
            asm volatile("bar.sync 1, %0;" ::"n"(NI * 32) : "memory");
            size_t t = (size_t)next[round & 1];
            if (t >= a.n) break;
            t += lt;
            if (t >= a.n) continue;
            ge_p3 p, r;
            ge_p3_load(p, a.pts + QQ_PT_Q * map_index(a.map, t));
            vbs_build_tables(tbl, p);
            u32 s[8];
            load_words32(s, a.s0, t / (size_t)a.sdiv);
            vbs_scalarmult<false>(r, tbl, s);
            ge_p3_store(a.out0 + QQ_PT_Q * t, r);
            load_words32(s, a.s1, t / (size_t)a.sdiv);
            vbs_scalarmult<false>(r, tbl, s);
            ge_p3_store(a.out1 + QQ_PT_Q * t, r);
        }
        if (lt == 0) done = 1;
    } else {
        fe64 x, y;
        fe64_from_fe(x, fe_2d());
        fe64_from_fe(y, fe_sqrt_m1());
        unsigned long long it = 0;
        while (!done) {
#pragma unroll 1
            for (int k = 0; k < 8; k++) {
                x = fe64_mul_ool(x, y);
                y = fe64_sq_ool(x);
            }
            it += 16;
        }
        if ((lt & 31) == 0) atomicAdd(counter + 2, it);
        if (x.v[0] + y.v[3] == 1.2345) sink[0] = x.v[1];
    }
}
template <int NI, int NF>
static void probe(size_t n, int sms) {
    auto kern = k_probe<NI, NF>;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    u32x4 *pts, *s0, *s1, *o0, *o1, *scratch;
    unsigned long long* counter;
    double* sink;
    cudaMalloc(&pts, n * QQ_PT_BYTES); cudaMalloc(&s0, n * 32); cudaMalloc(&s1, n * 32);
    cudaMalloc(&o0, n * QQ_PT_BYTES); cudaMalloc(&o1, n * QQ_PT_BYTES);
    cudaMalloc(&scratch, (size_t)sms * NI * 32 * QQ_VBS_TABLE_WORDS * 4);
    cudaMalloc(&counter, 64); cudaMalloc(&sink, 64);
    cudaMemset(pts, 0x5a, n * QQ_PT_BYTES); cudaMemset(s0, 0x07, n * 32); cudaMemset(s1, 0x03, n * 32);
    vb_args a;
    a.pts = pts; a.map = {1, 1, {0, 0, 0, 0}}; a.s0 = s0; a.s1 = s1; a.sdiv = 1; a.halve0 = 0; a.halve1 = 0;
    a.out0 = o0; a.out1 = o1; a.scratch = scratch; a.n = n;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    unsigned long long st[3] = {0, 0, 0}, fits = 0;
    for (int rep = 0; rep < 3; rep++) {
        cudaMemsetAsync(counter, 0, 64);
        cudaEventRecord(e0);
        kern<<<sms, (NI + NF) * 32>>>(a, counter, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(st, counter, 24, cudaMemcpyDeviceToHost);
        if (ms < best) { best = ms; fits = st[2]; }
    }
    // one FP64 scalar multiplication of the split form = 156 dbl + 78 add = 624 sq + 507 mul + 624 mul = 1755 products
    double fp64_products = (double)fits * 32;
    printf("{\"probe\": \"int warps + register-only FP64 product chain\", \"int_warps\": %d, \"fp64_warps\": %d, \"regs\": %d, \"ms\": %.3f, "
           "\"int_scalar_mults_per_s\": %.4e, \"fp64_products_per_s\": %.4e, \"fp64_equiv_scalar_mults_per_s\": %.4e}\n",
           NI, NF, fa.numRegs, best, 2.0 * n / (best * 1e-3), fp64_products / (best * 1e-3), fp64_products / 1755.0 / (best * 1e-3));
    fflush(stdout);
    cudaFree(pts); cudaFree(s0); cudaFree(s1); cudaFree(o0); cudaFree(o1); cudaFree(scratch); cudaFree(counter); cudaFree(sink);
}

int main(int argc, char** argv) {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    size_t n = argc > 1 ? (size_t)atol(argv[1]) : (size_t)1 << 19;
    printf("{\"device\": \"%s\", \"sms\": %d}\n", prop.name, sms);
    if (argc > 2 && !strcmp(argv[2], "probe")) {
        probe<8, 1>(n, sms);
        probe<8, 4>(n, sms);
        probe<8, 8>(n, sms);
        probe<12, 4>(n, sms);
        probe<16, 4>(n, sms);
        probe<16, 8>(n, sms);
        return 0;
    }
    run<16, 0>(n, sms, true);
    run<0, 4>(n, sms, false);
    run<0, 8>(n, sms, false);
    run<12, 4>(n, sms, false);
    run<8, 4>(n, sms, false);
    run<8, 8>(n, sms, false);
    run<12, 0>(n, sms, false);
    run<8, 0>(n, sms, false);
    return 0;
}
