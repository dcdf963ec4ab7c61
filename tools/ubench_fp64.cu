// FP64-pipe throughput micro-benchmark for sm_100a (B200): is DFMA a viable multiplier for GF(2^255-19)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ubench_fp64 tools/ubench_fp64.cu
// 8 independent chains per thread; operands change every step so nothing is hoisted.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define INNER 64
#define OUTER 256

// MODE 0: 8x DFMA.RN   1: 8x DFMA.RZ   2: 8x DADD   3: 8x DFMA + 8x IADD3 (u64 add = 2 ALU)
// MODE 4: 8x DFMA + 8x IMAD.WIDE   5: 8x DFMA + 4x IMAD.WIDE   6: 8x DFMA + 16 ALU   7: 8x DMUL
template <int MODE>
__global__ void __launch_bounds__(256) k_ubench(double* out, double seed, long long* cycles) {
    double x[2], b[4], c[8];
    uint64_t w[8];
    uint32_t u[8], v[8];
    for (int j = 0; j < 4; j++) b[j] = seed * 1.0000001 + j * 0.5 + threadIdx.x * 1e-3;
    for (int i = 0; i < 8; i++) { c[i] = i + seed; w[i] = (uint64_t)(i * 77 + threadIdx.x) * 2654435761u; u[i] = i * 3 + threadIdx.x; v[i] = i * 5 + 1; }
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = 1.0 + i * 1e-9 + seed * 1e-12;
    __syncthreads();
    long long t0 = clock64();
    for (int o = 0; o < OUTER; o++) {
#pragma unroll
        for (int k = 0; k < INNER; k++) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                x[i] = sm[(threadIdx.x + (o * INNER + k) * 2 + i) & 1023];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int q = i * 4 + j;
                    if (MODE == 0 || MODE >= 3 && MODE != 7) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(c[q]) : "d"(x[i]), "d"(b[j]));
                    if (MODE == 1) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(c[q]) : "d"(x[i]), "d"(b[j]));
                    if (MODE == 2) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(c[q]) : "d"(x[i]));
                    if (MODE == 7) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(c[q]) : "d"(x[i]));
                    if (MODE == 3) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(u[q]), "+r"(v[q]) : "r"(u[(q + 1) & 7]), "r"(v[(q + 3) & 7]));
                    if (MODE == 4 || (MODE == 5 && (j & 1))) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[q]) : "r"(u[q]), "r"(v[q]));
                    if (MODE == 6) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3; xor.b32 %0, %0, %3; shf.r.wrap.b32 %1, %1, %2, 7;" : "+r"(u[q]), "+r"(v[q]) : "r"(u[(q + 1) & 7]), "r"(v[(q + 3) & 7]));
                }
            }
        }
    }
    long long t1 = clock64();
    double acc = x[0] + x[1];
#pragma unroll
    for (int i = 0; i < 8; i++) acc += c[i] + (double)(w[i] ^ u[i] ^ v[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int nsm, int blocks_per_sm, int threads) {
    int grid = nsm * blocks_per_sm;
    double* out;
    long long* cyc;
    cudaMalloc(&out, (size_t)grid * threads * 8);
    cudaMalloc(&cyc, grid * sizeof(long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_ubench<MODE><<<grid, threads>>>(out, 1.5, cyc);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k_ubench<MODE><<<grid, threads>>>(out, 1.5 + r, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    long long* h = (long long*)malloc(grid * sizeof(long long));
    cudaMemcpy(h, cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double avgc = 0;
    for (int i = 0; i < grid; i++) avgc += (double)h[i];
    avgc /= grid;
    free(h);
    double per_thread = (double)OUTER * INNER * 8;
    double per_s = (double)grid * threads * per_thread / (best * 1e-3);
    printf("{\"test\": \"%s\", \"blocks_per_sm\": %d, \"threads\": %d, \"ms\": %.4f, \"primary_ops_per_s\": %.4e, "
           "\"lanes_per_clk_per_sm_at_1965\": %.2f, \"lanes_per_clk_per_sm_clock64\": %.2f, \"implied_mhz\": %.0f}\n", name, blocks_per_sm, threads, best, per_s, per_s / nsm / 1.965e9,
           (double)blocks_per_sm * threads * per_thread / avgc, avgc / (best * 1e-3) / 1e6);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, nsm, p.clockRate);
    for (int bps = 1; bps <= 2; bps *= 2) {
        run<0>("dfma_rn", nsm, bps, 256);
        run<1>("dfma_rz", nsm, bps, 256);
        run<2>("dadd", nsm, bps, 256);
        run<7>("dmul", nsm, bps, 256);
        run<3>("dfma+u64add", nsm, bps, 256);
        run<4>("dfma+imad_wide", nsm, bps, 256);
        run<5>("dfma+0.5imad_wide", nsm, bps, 256);
        run<6>("dfma+4alu", nsm, bps, 256);
    }
    return 0;
}
