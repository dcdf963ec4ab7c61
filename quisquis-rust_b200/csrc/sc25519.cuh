// Scalar (mod l) helpers needed on the device: canonical check and signed fixed-window recoding.
// Replaces curve25519-dalek 3.x scalar.rs::{is_canonical, to_radix_16, to_radix_2w} for the hot path.
// All scalar *algebra* (sums, products, challenges) stays on the host in the reference and here.
#pragma once
#include "fe25519.cuh"

namespace qq {

// l = 2^252 + 27742317777372353535851937790883648493, little-endian words
QQ_HD u32 sc_l_word(int i) {
    const u32 L[8] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0u, 0u, 0u, 0x10000000u};
    return L[i];
}
// 1 if s < l
QQ_HD u32 sc_is_canonical(const u32 s[8]) {
    // compute borrow of s - l
    u32 borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u64 d = (u64)s[i] - sc_l_word(i) - borrow;
        borrow = (u32)(d >> 63);
    }
    return borrow;
}

// h = s / 2 mod l for s < l:  s >> 1 when s is even, (s + l) >> 1 when odd.  Used with the double-and-compress encoder
// (compress_batch.cuh): enc(s P) = enc(2 (h P)).  Non-canonical input gives an unspecified (but bounded) result; the
// callers zero such outputs through the status byte.
QQ_HD void sc_halve(u32 h[8], const u32 s[8]) {
    u32 m = 0u - (s[0] & 1u);
    u32 t[8];
    u32 carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u64 a = (u64)s[i] + (sc_l_word(i) & m) + carry;
        t[i] = (u32)a;
        carry = (u32)(a >> 32);
    }
#pragma unroll
    for (int i = 0; i < 7; i++) h[i] = (t[i] >> 1) | (t[i + 1] << 31);
    h[7] = (t[7] >> 1) | (carry << 31);
}

// Signed radix-2^W recoding without digit storage: r = s + sum_i 2^(W-1) 2^(W i)  (NW windows, NW*W >= 255), then
// digit_i = ((r >> W i) & (2^W - 1)) - 2^(W-1)  in [-2^(W-1), 2^(W-1)).   s < 2^253.
// r has 9 words (NW*W may exceed 256).
template <int W, int NW>
QQ_HD void sc_recode_bias(u32 r[9], const u32 s[8]) {
    u32 c[9];
#pragma unroll
    for (int i = 0; i < 9; i++) c[i] = 0;
#pragma unroll
    for (int k = 0; k < NW; k++) {
        int bit = W * k + (W - 1);
        if (bit < 288) c[bit >> 5] |= 1u << (bit & 31);
    }
    u32 carry = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        u64 t = (u64)(i < 8 ? s[i] : 0u) + c[i] + carry;
        r[i] = (u32)t;
        carry = (u32)(t >> 32);
    }
}
// digit k (compile-time or runtime k; with runtime k the array must live in memory, use sc_digit_top for registers)
template <int W>
QQ_HD int sc_digit(const u32 r[9], int k) {
    int bit = W * k;
    int wi = bit >> 5, sh = bit & 31;
    u64 two = (u64)r[wi] | ((u64)(wi + 1 < 9 ? r[wi + 1] : 0u) << 32);
    u32 raw = (u32)(two >> sh) & ((1u << W) - 1u);
    return (int)raw - (1 << (W - 1));
}

}  // namespace qq
