// Scalar multiplication building blocks shared by the kernels (and by the host-compiled unit tests).
//
//  * variable base: signed radix-16 fixed window, 9-entry projective-Niels table {0,1..8}P kept in memory
//    (global scratch on the GPU), 63 x (3 dbl-noT + 1 dbl) + 64 table additions, uniform control flow.
//    Same algorithm family as curve25519-dalek 3.x backend/serial/scalar_mul/variable_base.rs, which the reference
//    reaches through `&Scalar * &RistrettoPoint` (src/ristretto/keys.rs:279-280, src/elgamal/elgamal.rs:47,50).
//  * fixed base: signed radix-2^W windows over a precomputed affine-Niels table
//    tbl[k][j] = j * 2^(W k) * Base, j = 0..2^(W-1); NW madds, no doublings.  Replaces
//    `&Scalar * &RISTRETTO_BASEPOINT_TABLE` (src/elgamal/elgamal.rs:49,85) and the reference's variable-base use of
//    the constants B, H in create_delta_and_epsilon_accounts (src/accounts/accounts.rs:214).
#pragma once
#include "ge25519.cuh"
#include "sc25519.cuh"

namespace qq {

struct alignas(16) u32x4 {
    u32 x, y, z, w;
};

// Memory layouts (all in 16-byte units, u32x4): an extended or cached point is 4 field elements = 8 x 16 B = 128 B,
// an affine Niels point is 3 field elements = 96 B.
#define QQ_PT_Q 8
#define QQ_PT_BYTES (QQ_PT_Q * 16)
#define QQ_VB_ENTRIES 9
#define QQ_VB_TABLE_WORDS (QQ_VB_ENTRIES * QQ_PT_Q * 4)

QQ_HD void fe_store(u32x4* dst, int& o, const fe& a) {
    u32x4 q;
    q.x = a.v[0]; q.y = a.v[1]; q.z = a.v[2]; q.w = a.v[3]; dst[o++] = q;
    q.x = a.v[4]; q.y = a.v[5]; q.z = a.v[6]; q.w = a.v[7]; dst[o++] = q;
}
QQ_HD void fe_load(const u32x4* src, int& o, fe& a) {
    u32x4 q;
    q = src[o++]; a.v[0] = q.x; a.v[1] = q.y; a.v[2] = q.z; a.v[3] = q.w;
    q = src[o++]; a.v[4] = q.x; a.v[5] = q.y; a.v[6] = q.z; a.v[7] = q.w;
}
QQ_HD void fe_store4(u32x4* dst, int& o, const fe& a, const fe& b) {
    fe_store(dst, o, a);
    fe_store(dst, o, b);
}
QQ_HD void fe_load4(const u32x4* src, int& o, fe& a, fe& b) {
    fe_load(src, o, a);
    fe_load(src, o, b);
}
QQ_HD void ge_cached_store(u32x4* dst, const ge_cached& c) {
    int o = 0;
    fe_store4(dst, o, c.YpX, c.YmX);
    fe_store4(dst, o, c.Z2, c.T2d);
}
QQ_HD void ge_cached_load(ge_cached& c, const u32x4* src) {
    int o = 0;
    fe_load4(src, o, c.YpX, c.YmX);
    fe_load4(src, o, c.Z2, c.T2d);
}
QQ_HD void ge_p3_store(u32x4* dst, const ge_p3& p) {
    int o = 0;
    fe_store4(dst, o, p.X, p.Y);
    fe_store4(dst, o, p.Z, p.T);
}
QQ_HD void ge_p3_load(ge_p3& p, const u32x4* src) {
    int o = 0;
    fe_load4(src, o, p.X, p.Y);
    fe_load4(src, o, p.Z, p.T);
}

// Build the 9-entry table {0P, 1P, ..., 8P} (cached form) into tbl (9 x 8 x u32x4 = 1152 B).
QQ_HD void vb_build_table(u32x4* tbl, const ge_p3& p) {
    ge_cached c0, c;
    ge_p3 id, q;
    ge_identity(id);
    ge_to_cached(c, id);
    ge_cached_store(tbl, c);
    ge_to_cached(c0, p);
    ge_cached_store(tbl + QQ_PT_Q, c0);
    q = p;
    for (int i = 2; i <= 8; i++) {
        ge_add(q, q, c0);
        ge_to_cached(c, q);
        ge_cached_store(tbl + QQ_PT_Q * i, c);
    }
}

// Table lookup of entry idx (0 .. nent - 1) in cached form.
// SECRET = false: the entry is read directly - uniform control flow, but the ADDRESS depends on the digit (verifier side,
// public scalars).  SECRET = true (qq_set_secret_mode: the wallet's u, c, bl, r, sk and the provers' blindings): every entry is
// read and the wanted one kept with masks, so the address stream is the same for every scalar - the constant-time table
// access curve25519-dalek's LookupTable::select gives the reference (window.rs; reached through variable_base::mul).
template <bool SECRET>
QQ_HD void vb_lookup(ge_cached& c, const u32x4* tbl, u32 idx, int nent = QQ_VB_ENTRIES) {
    if constexpr (!SECRET) {
        ge_cached_load(c, tbl + QQ_PT_Q * idx);
    } else {
    u32 acc[32];
#pragma unroll
    for (int w = 0; w < 32; w++) acc[w] = 0;
    for (int i = 0; i < nent; i++) {
        const u32 m = 0u - (u32)((((u32)i ^ idx) - 1u) >> 31);      // all ones when i == idx
#pragma unroll
        for (int q = 0; q < QQ_PT_Q; q++) {
            const u32x4 v = tbl[QQ_PT_Q * i + q];
            acc[4 * q + 0] |= v.x & m;
            acc[4 * q + 1] |= v.y & m;
            acc[4 * q + 2] |= v.z & m;
            acc[4 * q + 3] |= v.w & m;
        }
    }
#pragma unroll
    for (int w = 0; w < 8; w++) {
        c.YpX.v[w] = acc[w];
        c.YmX.v[w] = acc[8 + w];
        c.Z2.v[w] = acc[16 + w];
        c.T2d.v[w] = acc[24 + w];
    }
    }
}

// r = s * P using a table built by vb_build_table.  s: 8 little-endian words, s < 2^253.
// ROLLED keeps one copy of the doubling body (smaller instruction footprint; T is computed by every doubling).
template <bool ROLLED, bool SECRET = false>
QQ_HD void vb_scalarmult_t(ge_p3& r, const u32x4* tbl, const u32 s[8]) {
    u32 rr[9];
    sc_recode_bias<4, 64>(rr, s);        // rr[8] == 0 for s < 2^253
    u32 w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = rr[i];
    ge_identity(r);
#pragma unroll 1
    for (int k = 63; k >= 0; k--) {
        if (k != 63) {
            if (ROLLED) {
#pragma unroll 1
                for (int d = 0; d < 4; d++) ge_dbl<true>(r, r);
            } else {
                ge_dbl<false>(r, r);
                ge_dbl<false>(r, r);
                ge_dbl<false>(r, r);
                ge_dbl<true>(r, r);
            }
        }
        int d = (int)(w[7] >> 28) - 8;    // signed digit in [-8, 8)
        // shift the 256-bit register left by one nibble
#pragma unroll
        for (int i = 7; i > 0; i--) w[i] = (w[i] << 4) | (w[i - 1] >> 28);
        w[0] <<= 4;
        u32 neg = (u32)d >> 31;
        u32 idx = (u32)((d ^ (d >> 31)) - (d >> 31));      // |d| without a branch
        ge_cached c;
        vb_lookup<SECRET>(c, tbl, idx);
        ge_cached_cneg(c, neg);
        ge_add(r, r, c);
    }
}
QQ_HD void vb_scalarmult(ge_p3& r, const u32x4* tbl, const u32 s[8]) { vb_scalarmult_t<false>(r, tbl, s); }
QQ_HD void vb_scalarmult_secret(ge_p3& r, const u32x4* tbl, const u32 s[8]) { vb_scalarmult_t<false, true>(r, tbl, s); }

// ---------------------------------------------------------------------------------------------------------
// Split variable base, for a point that is multiplied by SEVERAL scalars (update_account multiplies each of gr, grsk
// by both u and c, reference src/accounts/accounts.rs:146-152).  The 64 signed radix-16 digits are cut into
// QQ_VBS_PARTS quarters of 16 digits; quarter q works on P_q = 2^(64 q) P:
//     s P = sum_q s_q P_q,   s_q = digits 16q .. 16q+15,
// evaluated Straus-style: 15 x 4 doublings + 64 table additions per scalar.  The 192 doublings that produce P_1..P_3
// and the four 9-entry tables are paid once per point, so two scalars cost 192 + 2 * 60 = 312 doublings instead of
// 2 * 252 = 504 (additions unchanged).  Same digits, same tables, same uniform control flow as vb_scalarmult.
// ---------------------------------------------------------------------------------------------------------
#define QQ_VBS_PARTS 4
#define QQ_VBS_TABLE_Q (QQ_VBS_PARTS * QQ_VB_ENTRIES * QQ_PT_Q)
#define QQ_VBS_TABLE_WORDS (QQ_VBS_TABLE_Q * 4)

QQ_HD void vbs_build_tables(u32x4* tbl, const ge_p3& p) {
    ge_p3 q = p;
#pragma unroll 1
    for (int part = 0; part < QQ_VBS_PARTS; part++) {
        vb_build_table(tbl + part * (QQ_VB_ENTRIES * QQ_PT_Q), q);
        if (part + 1 < QQ_VBS_PARTS) {
#pragma unroll 1
            for (int i = 0; i < 63; i++) ge_dbl<false>(q, q);
            ge_dbl<true>(q, q);
        }
    }
}
template <bool SECRET = false>
QQ_HD void vbs_scalarmult(ge_p3& r, const u32x4* tbl, const u32 s[8]) {
    u32 rr[9];
    sc_recode_bias<4, 64>(rr, s);        // rr[8] == 0 for s < 2^253
    ge_identity(r);
#pragma unroll 1
    for (int half = 1; half >= 0; half--) {
        // digits 8..15 (half = 1) then 0..7 (half = 0) of every quarter: word 2q + half of the biased scalar
        u32 w0 = half ? rr[1] : rr[0], w1 = half ? rr[3] : rr[2], w2 = half ? rr[5] : rr[4], w3 = half ? rr[7] : rr[6];
#pragma unroll 1
        for (int j = 7; j >= 0; j--) {
            if (!(half == 1 && j == 7)) {
                ge_dbl<false>(r, r);
                ge_dbl<false>(r, r);
                ge_dbl<false>(r, r);
                ge_dbl<true>(r, r);
            }
#pragma unroll 1
            for (int part = 0; part < QQ_VBS_PARTS; part++) {
                int d = (int)(w0 >> 28) - 8;    // signed digit in [-8, 8)
                w0 = (w0 << 4);
                // rotate the four quarter registers so that the loop body stays one copy
                u32 tw = w0; w0 = w1; w1 = w2; w2 = w3; w3 = tw;
                u32 neg = (u32)d >> 31;
                u32 idx = (u32)((d ^ (d >> 31)) - (d >> 31));
                ge_cached c;
                vb_lookup<SECRET>(c, tbl + QQ_PT_Q * (part * QQ_VB_ENTRIES), idx);
                ge_cached_cneg(c, neg);
                ge_add(r, r, c);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fixed-base tables.  Layout: entry (k, j) at tbl[(k * (2^(W-1) + 1) + j) * 24 .. +24] words:
// ypx[8], ymx[8], xy2d[8].  j = 0 is the identity (1, 1, 0).
// ---------------------------------------------------------------------------------------------------------
#define QQ_NIELS_WORDS 24

QQ_HD int fb_num_windows(int w) { return (256 + w - 1) / w; }   // W*NW >= 256 keeps the recoding carry-free
QQ_HD int fb_entries(int w) { return (1 << (w - 1)) + 1; }

QQ_HD void ge_niels_load(ge_niels& n, const u32* src) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        n.ypx.v[i] = src[i];
        n.ymx.v[i] = src[8 + i];
        n.xy2d.v[i] = src[16 + i];
    }
}

// r = s * Base from table (W-bit signed windows).  tbl may point to shared or global memory.  SECRET: see vb_lookup (all
// 2^(W-1) + 1 entries of the window are read and masked; meant for the shared-memory table).
template <int W, bool SECRET = false>
QQ_HD void fb_scalarmult(ge_p3& r, const u32* tbl, const u32 s[8]) {
    const int NW = (256 + W - 1) / W;
    const int ENT = (1 << (W - 1)) + 1;
    u32 rr[9];
    sc_recode_bias<W, NW>(rr, s);
    ge_identity(r);
#pragma unroll 1
    for (int k = 0; k < NW; k++) {
        // extract digit k: shift-free approach needs dynamic word index -> read via small switchless funnel
        int bit = W * k;
        int wi = bit >> 5, sh = bit & 31;
        u32 lo = 0, hi = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            lo = (i == wi) ? rr[i] : lo;
            hi = (i == wi + 1) ? rr[i] : hi;
        }
        u64 two = (u64)lo | ((u64)hi << 32);
        int d = (int)((u32)(two >> sh) & ((1u << W) - 1u)) - (1 << (W - 1));
        u32 neg = (u32)d >> 31;
        u32 idx = (u32)((d ^ (d >> 31)) - (d >> 31));
        ge_niels n;
        if (SECRET) {
            u32 acc[QQ_NIELS_WORDS];
#pragma unroll
            for (int w = 0; w < QQ_NIELS_WORDS; w++) acc[w] = 0;
            for (int i = 0; i < ENT; i++) {
                const u32 m = 0u - (u32)((((u32)i ^ idx) - 1u) >> 31);
                const u32* e = tbl + (size_t)(k * ENT + i) * QQ_NIELS_WORDS;
#pragma unroll
                for (int w = 0; w < QQ_NIELS_WORDS; w++) acc[w] |= e[w] & m;
            }
#pragma unroll
            for (int w = 0; w < 8; w++) {
                n.ypx.v[w] = acc[w];
                n.ymx.v[w] = acc[8 + w];
                n.xy2d.v[w] = acc[16 + w];
            }
        } else {
            ge_niels_load(n, tbl + (size_t)(k * ENT + idx) * QQ_NIELS_WORDS);
        }
        ge_niels_cneg(n, neg);
        ge_madd(r, r, n);
    }
}

// Field inversion z^(p-2) (used only when normalising precomputed tables; not on the hot path).
QQ_HD void fe_invert(fe& out, const fe& z) {
    // z^(p-2) = z^(2^255 - 21) = (z^(2^252-3))^8 * z^3
    fe t, z2, z3;
    fe_pow22523(t, z);      // z^(2^252-3)
    fe_sqn(t, t, 3);        // z^(2^255-24)
    fe_sq(z2, z);
    fe_mul(z3, z2, z);
    fe_mul(out, t, z3);     // z^(2^255-21)
}

// Table entry (k, j) = (j << (W k)) * base by plain double-and-add (table construction only, one-off).
QQ_HD void fb_build_entry(u32* dst, const ge_p3& base, int W, int k, int j);

// Convert an extended point to affine Niels (one inversion).  For table construction only.
QQ_HD void ge_to_niels_affine(u32* dst, const ge_p3& p) {
    fe zi, x, y, t;
    fe_invert(zi, p.Z);
    fe_mul(x, p.X, zi);
    fe_mul(y, p.Y, zi);
    fe ypx, ymx, xy2d;
    fe_add(ypx, y, x);
    fe_sub(ymx, y, x);
    fe_mul(t, x, y);
    fe_mul(xy2d, t, fe_2d());
#pragma unroll
    for (int i = 0; i < 8; i++) {
        dst[i] = ypx.v[i];
        dst[8 + i] = ymx.v[i];
        dst[16 + i] = xy2d.v[i];
    }
}

QQ_HD void fb_build_entry(u32* dst, const ge_p3& base, int W, int k, int j) {
    ge_p3 r;
    ge_identity(r);
    ge_cached cb;
    ge_to_cached(cb, base);
    for (int b = W - 1; b >= 0; b--) {
        ge_dbl<true>(r, r);
        if ((j >> b) & 1) ge_add(r, r, cb);
    }
    for (int i = 0; i < W * k; i++) ge_dbl<true>(r, r);
    ge_to_niels_affine(dst, r);
}

}  // namespace qq
