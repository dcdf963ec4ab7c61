// Instruction-throughput micro-benchmark for the sm_100a integer pipes.
// Establishes the roofline denominator (thread-level IMAD-class ops / s) used by bench.py and DESIGN.md.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_imad tools/ubench_imad.cu
// Each test keeps 8 independent accumulator chains per thread; multipliers change every step (x += a) so
// ptxas cannot hoist or CSE products.  The SASS of every loop body was inspected with cuobjdump.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define INNER 64
#define OUTER 256

// MODE 0: 8x IMAD (lo)            per step, +2 IADD
// MODE 1: 8x IMAD.HI.U32          per step, +2 IADD
// MODE 2: 8x IMAD.WIDE.U32 (64-bit accumulate) per step, +2 IADD
// MODE 3: 8x IADD3 only
// MODE 4: 8x IMAD.WIDE + 8 ALU (IADD3)  -> 1:1 mix
// MODE 5: 8x IMAD.WIDE + 16 ALU         -> 1:2 mix
// MODE 6: 8x IMAD.WIDE + 8 SHF          -> 1:1 mix with funnel shifts
// MODE 7: 8x IMAD.WIDE + 4 ALU          -> 2:1 mix
template <int MODE>
__global__ void __launch_bounds__(256) k_ubench(uint32_t* out, uint32_t seed, long long* cycles) {
    uint32_t x[2], a[2], b[4], c[8], e[8];
    uint64_t w[8];
    for (int i = 0; i < 2; i++) { x[i] = seed * (threadIdx.x + 1) + i * 77u; a[i] = (seed ^ (blockIdx.x * 131u + i)) | 1u; }
    for (int j = 0; j < 4; j++) b[j] = seed * 2654435761u + j * 40503u + threadIdx.x;
    for (int i = 0; i < 8; i++) { c[i] = i + seed; e[i] = i * seed + 3; w[i] = ((uint64_t)c[i] << 32) | b[i & 3]; }
    __shared__ uint32_t sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 2654435761u + seed;
    __syncthreads();
    long long t0 = clock64();
    for (int o = 0; o < OUTER; o++) {
#pragma unroll
        for (int k = 0; k < INNER; k++) {
#pragma unroll
            for (int i = 0; i < 2; i++) {
                x[i] = sm[(threadIdx.x + (o * INNER + k) * 2 + i + a[i]) & 1023];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    int q = i * 4 + j;
                    if (MODE == 0) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c[q]) : "r"(x[i]), "r"(b[j]));
                    if (MODE == 1) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c[q]) : "r"(x[i]), "r"(b[j]));
                    if (MODE == 2 || MODE >= 4) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[q]) : "r"(x[i]), "r"(b[j]));
                    if (MODE == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(c[q]) : "r"(b[j]));
                    if (MODE == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(c[q]) : "r"(b[j]));
                    if (MODE == 5) {
                        asm volatile("add.u32 %0, %0, %1;" : "+r"(c[q]) : "r"(b[j]));
                        asm volatile("xor.b32 %0, %0, %1;" : "+r"(e[q]) : "r"(c[q]));
                    }
                    if (MODE == 6) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(c[q]) : "r"(e[q]));
                    if (MODE == 7 && (j & 1)) asm volatile("add.u32 %0, %0, %1;" : "+r"(c[q]) : "r"(b[j]));
                }
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = x[0] ^ x[1];
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= c[i] ^ e[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int nsm, int blocks_per_sm, int threads) {
    int grid = nsm * blocks_per_sm;
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, (size_t)grid * threads * 4);
    cudaMalloc(&cyc, grid * sizeof(long long));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_ubench<MODE><<<grid, threads>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        cudaEventRecord(e0);
        k_ubench<MODE><<<grid, threads>>>(out, 12345u + r, cyc);
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
    double per_thread = (double)OUTER * INNER * 8;  // primary ops per thread
    double per_s = (double)grid * threads * per_thread / (best * 1e-3);
    double per_clk_sm = (double)blocks_per_sm * threads * per_thread / avgc;
    printf("{\"test\": \"%s\", \"blocks_per_sm\": %d, \"threads\": %d, \"ms\": %.4f, \"primary_ops_per_s\": %.4e, "
           "\"primary_ops_per_clk_per_sm\": %.2f, \"implied_clock_mhz\": %.0f}\n",
           name, blocks_per_sm, threads, best, per_s, per_clk_sm, per_s / (per_clk_sm * nsm) / 1e6);
    cudaFree(out);
    cudaFree(cyc);
    free(h);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int nsm = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, nsm, p.clockRate);
    for (int bps = 1; bps <= 4; bps *= 2) {
        run<0>("imad_lo", nsm, bps, 256);
        run<1>("imad_hi", nsm, bps, 256);
        run<2>("imad_wide_acc64", nsm, bps, 256);
        run<3>("iadd3", nsm, bps, 256);
        run<4>("imad_wide+1alu", nsm, bps, 256);
        run<5>("imad_wide+2alu", nsm, bps, 256);
        run<6>("imad_wide+1shf", nsm, bps, 256);
        run<7>("imad_wide+0.5alu", nsm, bps, 256);
    }
    return 0;
}
