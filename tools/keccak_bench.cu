// Latency of Keccak-f[1600] as the transcript kernels run it: one thread per state (keccak_host.hpp f1600) against the
// warp-cooperative form (25 lanes hold one 64-bit word each).  One warp per block, one block per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/keccak_bench tools/keccak_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../quisquis-rust_b200/csrc/keccak_host.hpp"

__global__ void __launch_bounds__(32) k_single(uint64_t* st, int reps) {
    uint64_t a[25];
    uint64_t* mine = st + 25 * (size_t)(blockIdx.x * 32 + threadIdx.x);
    for (int i = 0; i < 25; i++) a[i] = mine[i];
    for (int r = 0; r < reps; r++) qq_keccak::f1600(a);
    for (int i = 0; i < 25; i++) mine[i] = a[i];
}
__global__ void __launch_bounds__(32) k_coop(uint64_t* st, int reps) {
    uint64_t a[25];
    uint64_t* mine = st + 25 * (size_t)blockIdx.x;     // one state per warp, every lane holds a copy
    for (int i = 0; i < 25; i++) a[i] = mine[i];
    for (int r = 0; r < reps; r++) qq_keccak::f1600_warp(a);
    if (threadIdx.x == 7)
        for (int i = 0; i < 25; i++) mine[i] = a[i];
}
int main(int argc, char** argv) {
    int reps = argc > 1 ? atoi(argv[1]) : 400;
    int sms = 148;
    size_t n = (size_t)sms * 32 * 25;
    uint64_t *d, *h = (uint64_t*)malloc(n * 8), *h2 = (uint64_t*)malloc(n * 8);
    for (size_t i = 0; i < n; i++) h[i] = 0x9e3779b97f4a7c15ULL * (i % (25 * 3) + 1);
    cudaMalloc(&d, n * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms;
    // host reference
    uint64_t ref[25];
    for (int i = 0; i < 25; i++) ref[i] = h[i];
    for (int r = 0; r < reps; r++) qq_keccak::f1600(ref);
    for (int pass = 0; pass < 2; pass++) {
        cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        k_single<<<sms, 32>>>(d, reps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaMemcpy(h2, d, n * 8, cudaMemcpyDeviceToHost);
    bool ok = true;
    for (int i = 0; i < 25; i++) ok &= h2[i] == ref[i];
    printf("{\"kernel\": \"f1600, one thread per state\", \"reps\": %d, \"us_per_permutation\": %.3f, \"matches_host\": %s}\n", reps, ms * 1e3 / reps, ok ? "true" : "false");
    for (int pass = 0; pass < 2; pass++) {
        cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
        cudaEventRecord(e0);
        k_coop<<<sms, 32>>>(d, reps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    cudaMemcpy(h2, d, n * 8, cudaMemcpyDeviceToHost);
    ok = true;
    for (int i = 0; i < 25; i++) ok &= h2[i] == ref[i];
    printf("{\"kernel\": \"f1600_warp, 25 lanes per state\", \"reps\": %d, \"us_per_permutation\": %.3f, \"matches_host\": %s}\n", reps, ms * 1e3 / reps, ok ? "true" : "false");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("cuda error: %s\n", cudaGetErrorString(e));
    return 0;
}
