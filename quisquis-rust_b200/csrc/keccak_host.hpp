// Keccak-f[1600] sponge on the HOST (FIPS 202): SHA3-512 and the SHAKE256 XOF.
// Only the hashing of a few labels / 32-byte encodings that precedes the hash-to-group map of generator derivation
// (reference src/pedersen/vectorpedersen.rs:45-75 uses sha3::Sha3_512; the Bulletproofs generator chains use Shake256);
// the field and curve work of hash-to-group runs on the GPU (k_from_uniform).  A few kilobytes per call: host code.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

namespace qq_keccak {

static inline uint64_t rotl(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

static inline void f1600(uint64_t a[25]) {
    static const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
        0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
        0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
        0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int round = 0; round < 24; round++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        // rho + pi: B[y][2x + 3y] = rot(A[x][y]);  index = x + 5 y
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl(a[x + 5 * y], ROT[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= RC[round];
    }
}

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
