// Field-multiplication bake-off on the GPU: 10-limb radix-2^25.5 (the shipping fe25519.cuh) versus a saturated
// 8 x 32-bit representation with 96-bit column accumulators.  Reports multiplications / squarings per second.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I quisquis-rust_b200/csrc -o tools/fe_bench tools/fe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "fe25519.cuh"
using namespace qq;
typedef unsigned __int128 u128;
struct fe8 { u32 v[8]; };

__device__ __forceinline__ void sat_reduce(fe8& h, const u32 t[16]) {
    u64 c = 0;
    u32 r[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { u64 w = mul_wide(t[8 + k], 38u); u128 s = (u128)w + t[k] + c; r[k] = (u32)s; c = (u64)(s >> 32); }
    u64 d = (u64)r[0] + (u64)(u32)c * 38u; r[0] = (u32)d; d >>= 32;
#pragma unroll
    for (int k = 1; k < 8; k++) { d += r[k]; r[k] = (u32)d; d >>= 32; }
    r[0] += (u32)d * 38u;
#pragma unroll
    for (int k = 0; k < 8; k++) h.v[k] = r[k];
}
__device__ __forceinline__ void sat_mul(fe8& h, const fe8& f, const fe8& g) {
    u32 t[16];
    u128 acc = 0;
#pragma unroll
    for (int k = 0; k < 15; k++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { int j = k - i; if (j < 0 || j > 7) continue; acc += mul_wide(f.v[i], g.v[j]); }
        t[k] = (u32)acc; acc >>= 32;
    }
    t[15] = (u32)acc;
    sat_reduce(h, t);
}
__device__ __forceinline__ void sat_sq(fe8& h, const fe8& f) {
    u32 t[16];
    u128 acc = 0;
    t[0] = 0;
#pragma unroll
    for (int k = 1; k < 15; k++) {
#pragma unroll
        for (int i = 0; i < 8; i++) { int j = k - i; if (j <= i || j > 7) continue; acc += mul_wide(f.v[i], f.v[j]); }
        t[k] = (u32)acc; acc >>= 32;
    }
    t[15] = (u32)acc;
    // double
#pragma unroll
    for (int k = 15; k > 0; k--) t[k] = (t[k] << 1) | (t[k - 1] >> 31);
    t[0] = 0;
    // add diagonal squares
    u64 c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u64 s = mul_wide(f.v[i], f.v[i]);
        u128 w = (u128)s + t[2 * i] + ((u64)t[2 * i + 1] << 32) + c;
        t[2 * i] = (u32)w; t[2 * i + 1] = (u32)(w >> 32); c = (u64)(w >> 64);
    }
    sat_reduce(h, t);
}

template <int MODE>
__global__ void __launch_bounds__(128) k_bench(u32* out, const u32* in, int iters) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 0 || MODE == 1) {
        fe x, y;
        for (int i = 0; i < 10; i++) { x.v[i] = in[(tid * 20 + i) % 4096] & 0x1ffffff; y.v[i] = in[(tid * 20 + 10 + i) % 4096] & 0x1ffffff; }
        for (int it = 0; it < iters; it++) {
            if (MODE == 0) { fe z; fe_mul(z, x, y); x = y; y = z; }
            else { fe_sq(x, x); }
        }
        u32 a = 0;
        for (int i = 0; i < 10; i++) a ^= x.v[i] ^ y.v[i];
        out[tid] = a;
    } else {
        fe8 x, y;
        for (int i = 0; i < 8; i++) { x.v[i] = in[(tid * 16 + i) % 4096]; y.v[i] = in[(tid * 16 + 8 + i) % 4096]; }
        for (int it = 0; it < iters; it++) {
            if (MODE == 2) { fe8 z; sat_mul(z, x, y); x = y; y = z; }
            else { sat_sq(x, x); }
        }
        u32 a = 0;
        for (int i = 0; i < 8; i++) a ^= x.v[i] ^ y.v[i];
        out[tid] = a;
    }
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
        run<0>("mul_10limb", nsm, bps, 4000);
        run<1>("sq_10limb", nsm, bps, 4000);
        run<2>("mul_sat8", nsm, bps, 4000);
        run<3>("sq_sat8", nsm, bps, 4000);
    }
    return 0;
}
