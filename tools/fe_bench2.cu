// Field-multiplication bake-off, round 2: the saturated 8 x 32-bit representation (fe25519.cuh).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -DQQ_INLINE_FIELD_OPS -I quisquis-rust_b200/csrc -o tools/bin/fe_bench2 tools/fe_bench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fe25519.cuh"
using namespace qq;

// MODE 0: Karatsuba mul  1: schoolbook mul  2: sq  3: add+sub pair
template <int MODE>
__global__ void __launch_bounds__(128) k_bench(u32* out, const u32* in, int iters) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    fe x, y;
    for (int i = 0; i < 8; i++) { x.v[i] = in[(tid * 16 + i) % 4096]; y.v[i] = in[(tid * 16 + 8 + i) % 4096]; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { fe z; fe_mul_inl(z, x, y); x = y; y = z; }
        if (MODE == 1) { fe z; fe_mul_school(z, x, y); x = y; y = z; }
        if (MODE == 2) { fe_sq_inl(x, x); }
        if (MODE == 3) { fe z; fe_add(z, x, y); fe_sub(x, y, z); y = z; }
    }
    u32 a = 0;
    for (int i = 0; i < 8; i++) a ^= x.v[i] ^ y.v[i];
    out[tid] = a;
}

template <int MODE>
static void run(const char* name, int nsm, int bps, int iters) {
    int grid = nsm * bps;
    u32 *out, *in;
    cudaMalloc(&out, (size_t)grid * 128 * 4);
    cudaMalloc(&in, 4096 * 4);
    u32 h[4096];
    for (int i = 0; i < 4096; i++) h[i] = 2654435761u * (i + 1);
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_bench<MODE><<<grid, 128>>>(out, in, 10);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        k_bench<MODE><<<grid, 128>>>(out, in, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double ops = (double)grid * 128 * iters;
    printf("{\"test\": \"%s\", \"blocks_per_sm\": %d, \"warps_per_sm\": %d, \"ms\": %.3f, \"ops_per_s\": %.4e}\n", name, bps, bps * 4, best, ops / (best * 1e-3));
    cudaFree(out); cudaFree(in);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    for (int bps : {2, 4, 8}) {
        run<0>("mul_sat8_karatsuba", nsm, bps, 4000);
        run<1>("mul_sat8_school", nsm, bps, 4000);
        run<2>("sq_sat8", nsm, bps, 4000);
        run<3>("addsub_pair_sat8", nsm, bps, 4000);
    }
    return 0;
}
