// Keccak-f[1600] sponge (FIPS 202), callable from host AND device code: SHA3-512 and the SHAKE256 XOF for generator
// derivation (reference src/pedersen/vectorpedersen.rs:45-75 uses sha3::Sha3_512; the Bulletproofs generator chains use
// Shake256: a few kilobytes per call, hashed on the host, the field and curve work of hash-to-group runs in k_from_uniform),
// and the permutation under the Merlin transcripts (merlin_host.hpp) that the batched verifiers run one-thread-per-proof in
// their transcript kernels (shuffle_verify.cuh).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

#ifdef __CUDACC__
#define QQ_HOSTDEV __host__ __device__
// One copy of the big bodies (the Keccak round function, a Z/l product, one transcript operation) in device code: the
// transcript kernels run one warp per SM, and with everything inlined pass A alone was 152 000 SASS instructions (2.4 MB) -
// 37 % of its issue slots waited for instruction fetch (ncu: stall_no_instruction, profiles/ncu_shuffle_pass_a_r02.json).
#define QQ_NOINLINE __noinline__
#else
#define QQ_HOSTDEV
#define QQ_NOINLINE
#endif
#define QQ_KECCAK_RC_WORDS {                                                                                              \
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, \
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, \
        0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, \
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL, \
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL}

namespace qq_keccak {

static const uint64_t RC_HOST[24] = QQ_KECCAK_RC_WORDS;
#ifdef __CUDACC__
static __constant__ uint64_t RC_DEV[24] = QQ_KECCAK_RC_WORDS;     // the same round constants through the constant bank
#endif

// 0 < n < 64.  Device: two funnel shifts on the 32-bit halves (n is a constant at every call site); written as shifts and an OR
// ptxas spends 4-5 instructions per rotation (279 instructions per round instead of ~210).
QQ_HOSTDEV static inline uint64_t rotl(uint64_t x, int n) {
#ifdef __CUDA_ARCH__
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if (n & 32) {
        const uint32_t t = lo;
        lo = hi;
        hi = t;
    }
    if ((n & 31) == 0) return ((uint64_t)hi << 32) | lo;
    return ((uint64_t)__funnelshift_l(lo, hi, n & 31) << 32) | __funnelshift_l(hi, lo, n & 31);
#else
    return (x << n) | (x >> (64 - n));
#endif
}

QQ_HOSTDEV QQ_NOINLINE static void f1600(uint64_t a[25]) {
#ifdef __CUDA_ARCH__
    const uint64_t* RC = RC_DEV;
#else
    const uint64_t* RC = RC_HOST;
#endif
    // one round written out on 25 locals (theta, rho + pi, chi, iota): the Fiat-Shamir transcripts of the batched verifiers
    // spend most of their time here
    uint64_t a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3], a4 = a[4];
    uint64_t a5 = a[5], a6 = a[6], a7 = a[7], a8 = a[8], a9 = a[9];
    uint64_t a10 = a[10], a11 = a[11], a12 = a[12], a13 = a[13], a14 = a[14];
    uint64_t a15 = a[15], a16 = a[16], a17 = a[17], a18 = a[18], a19 = a[19];
    uint64_t a20 = a[20], a21 = a[21], a22 = a[22], a23 = a[23], a24 = a[24];
#pragma unroll 2
    for (int round = 0; round < 24; round++) {
        const uint64_t c0 = a0 ^ a5 ^ a10 ^ a15 ^ a20, c1 = a1 ^ a6 ^ a11 ^ a16 ^ a21, c2 = a2 ^ a7 ^ a12 ^ a17 ^ a22,
                       c3 = a3 ^ a8 ^ a13 ^ a18 ^ a23, c4 = a4 ^ a9 ^ a14 ^ a19 ^ a24;
        const uint64_t d0 = c4 ^ rotl(c1, 1), d1 = c0 ^ rotl(c2, 1), d2 = c1 ^ rotl(c3, 1), d3 = c2 ^ rotl(c4, 1), d4 = c3 ^ rotl(c0, 1);
        const uint64_t b0 = (a0 ^ d0), b1 = rotl(a6 ^ d1, 44);
        const uint64_t b2 = rotl(a12 ^ d2, 43), b3 = rotl(a18 ^ d3, 21);
        const uint64_t b4 = rotl(a24 ^ d4, 14), b5 = rotl(a3 ^ d3, 28);
        const uint64_t b6 = rotl(a9 ^ d4, 20), b7 = rotl(a10 ^ d0, 3);
        const uint64_t b8 = rotl(a16 ^ d1, 45), b9 = rotl(a22 ^ d2, 61);
        const uint64_t b10 = rotl(a1 ^ d1, 1), b11 = rotl(a7 ^ d2, 6);
        const uint64_t b12 = rotl(a13 ^ d3, 25), b13 = rotl(a19 ^ d4, 8);
        const uint64_t b14 = rotl(a20 ^ d0, 18), b15 = rotl(a4 ^ d4, 27);
        const uint64_t b16 = rotl(a5 ^ d0, 36), b17 = rotl(a11 ^ d1, 10);
        const uint64_t b18 = rotl(a17 ^ d2, 15), b19 = rotl(a23 ^ d3, 56);
        const uint64_t b20 = rotl(a2 ^ d2, 62), b21 = rotl(a8 ^ d3, 55);
        const uint64_t b22 = rotl(a14 ^ d4, 39), b23 = rotl(a15 ^ d0, 41);
        const uint64_t b24 = rotl(a21 ^ d1, 2);
        a0 = b0 ^ (~b1 & b2); a1 = b1 ^ (~b2 & b3); a2 = b2 ^ (~b3 & b4); a3 = b3 ^ (~b4 & b0); a4 = b4 ^ (~b0 & b1);
        a5 = b5 ^ (~b6 & b7); a6 = b6 ^ (~b7 & b8); a7 = b7 ^ (~b8 & b9); a8 = b8 ^ (~b9 & b5); a9 = b9 ^ (~b5 & b6);
        a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14); a13 = b13 ^ (~b14 & b10); a14 = b14 ^ (~b10 & b11);
        a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19); a18 = b18 ^ (~b19 & b15); a19 = b19 ^ (~b15 & b16);
        a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24); a23 = b23 ^ (~b24 & b20); a24 = b24 ^ (~b20 & b21);
        a0 ^= RC[round];
    }
    a[0] = a0; a[1] = a1; a[2] = a2; a[3] = a3; a[4] = a4;
    a[5] = a5; a[6] = a6; a[7] = a7; a[8] = a8; a[9] = a9;
    a[10] = a10; a[11] = a11; a[12] = a12; a[13] = a13; a[14] = a14;
    a[15] = a15; a[16] = a16; a[17] = a17; a[18] = a18; a[19] = a19;
    a[20] = a20; a[21] = a21; a[22] = a22; a[23] = a23; a[24] = a24;
}

#ifdef __CUDACC__
// ---- warp-cooperative form -------------------------------------------------------------------------------------------------
// Every lane of a fully converged warp holds the SAME state a[25] (the transcript kernels in their warp-per-proof form: 32
// lanes run one proof's script redundantly).  Lane i < 25 carries word i through the 24 rounds - theta's column parities, the
// pi permutation and chi's row neighbours are warp shuffles (8 64-bit shuffles per round) - and all lanes receive all 25
// words at the end.  Lanes 25..31 shadow words 0..6; nothing reads them.  One thread alone needs ~4 600 dependent-ish
// instructions per permutation; here a round is ~45.
static __constant__ uint8_t RHO_DEV[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
    const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t rotl_var(uint64_t v, unsigned n) {      // 0 <= n < 64
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    if (n & 32) {
        const uint32_t t = lo;
        lo = hi;
        hi = t;
    }
    const uint32_t nlo = __funnelshift_l(hi, lo, n), nhi = __funnelshift_l(lo, hi, n);      // shift taken mod 32
    return ((uint64_t)nhi << 32) | nlo;
}
__device__ __noinline__ static void f1600_warp(uint64_t a[25]) {
    const int lane = threadIdx.x & 31;
    const int i = lane < 25 ? lane : lane - 25;
    const int x = i % 5, y = i / 5;
    const unsigned rho = RHO_DEV[i];
    const int up1 = i + 5 < 25 ? i + 5 : i - 20, up2 = i + 10 < 25 ? i + 10 : i - 15, up4 = i + 20 < 25 ? i + 20 : i - 5;
    const int col_m1 = (x + 4) % 5, col_p1 = (x + 1) % 5;                   // any lane of a column holds its parity: row 0
    const int src_pi = ((x + 3 * y) % 5) + 5 * x;                           // B[x'][y'] = rot(A[x][y]), x' = y, y' = 2x + 3y, as a gather
    const int row_p1 = 5 * y + (x + 1) % 5, row_p2 = 5 * y + (x + 2) % 5;
    uint64_t w = a[i];
#pragma unroll 1
    for (int round = 0; round < 24; round++) {
        const uint64_t s1 = w ^ shfl64(w, up1);
        const uint64_t c = s1 ^ shfl64(s1, up2) ^ shfl64(w, up4);
        const uint64_t d = shfl64(c, col_m1) ^ rotl_var(shfl64(c, col_p1), 1);
        const uint64_t b = shfl64(rotl_var(w ^ d, rho), src_pi);
        const uint64_t b1 = shfl64(b, row_p1), b2 = shfl64(b, row_p2);
        w = b ^ (~b1 & b2);
        if (i == 0) w ^= RC_DEV[round];
    }
#pragma unroll
    for (int k = 0; k < 25; k++) a[k] = shfl64(w, k);
}
#endif

struct sponge {
    uint64_t st[25];
    size_t rate, pos;
    uint8_t suffix;
    bool squeezing;
    sponge(size_t rate_bytes, uint8_t domain_suffix) : rate(rate_bytes), pos(0), suffix(domain_suffix), squeezing(false) {
        memset(st, 0, sizeof st);
    }
    void xor_byte(size_t i, uint8_t v) { st[i / 8] ^= (uint64_t)v << (8 * (i % 8)); }
    uint8_t get_byte(size_t i) const { return (uint8_t)(st[i / 8] >> (8 * (i % 8))); }
    void absorb(const void* data, size_t len) {
        const uint8_t* p = (const uint8_t*)data;
        for (size_t i = 0; i < len; i++) {
            xor_byte(pos++, p[i]);
            if (pos == rate) {
                f1600(st);
                pos = 0;
            }
        }
    }
    void squeeze(void* out, size_t len) {
        if (!squeezing) {
            xor_byte(pos, suffix);
            xor_byte(rate - 1, 0x80);
            f1600(st);
            pos = 0;
            squeezing = true;
        }
        uint8_t* o = (uint8_t*)out;
        for (size_t i = 0; i < len; i++) {
            if (pos == rate) {
                f1600(st);
                pos = 0;
            }
            o[i] = get_byte(pos++);
        }
    }
};

static inline void sha3_512(const void* data, size_t len, uint8_t out[64]) {
    sponge s(72, 0x06);
    s.absorb(data, len);
    s.squeeze(out, 64);
}
struct shake256 : sponge {
    shake256() : sponge(136, 0x1f) {}
};

}  // namespace qq_keccak
