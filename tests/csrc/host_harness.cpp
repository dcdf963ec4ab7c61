// TEST INFRASTRUCTURE: compiles the device arithmetic headers (fe25519/ge25519/ristretto/scalarmult .cuh) for the
// host with g++, so that tests/test_host_arith.py can check the exact limb arithmetic the kernels run against
// the oracle without a GPU.  This object is never linked into libqq_b200.so and is not a CPU fallback.
#include <cstring>
#include <string>
#include <vector>
#include "../../quisquis-rust_b200/csrc/ristretto.cuh"
#include "../../quisquis-rust_b200/csrc/scalarmult.cuh"
#include "../../quisquis-rust_b200/csrc/compress_batch.cuh"
#include "../../quisquis-rust_b200/csrc/sc_host.hpp"
#include "../../quisquis-rust_b200/csrc/keccak_host.hpp"
#include "../../quisquis-rust_b200/csrc/merlin_host.hpp"
#include "../../quisquis-rust_b200/csrc/shuffle_verify.cuh"
#include "../../tools/fe64/fe64.cuh"

using namespace qq;

static void load_words(u32 w[8], const uint8_t* b) { memcpy(w, b, 32); }
static void store_words(uint8_t* b, const u32 w[8]) { memcpy(b, w, 32); }

extern "C" {

void hh_fe_mul_limbs(u32* out, const u32* f, const u32* g) {
    fe a, b, c;
    memcpy(a.v, f, 32);
    memcpy(b.v, g, 32);
    fe_mul(c, a, b);
    memcpy(out, c.v, 32);
}
void hh_fe_sq_limbs(u32* out, const u32* f) {
    fe a, c;
    memcpy(a.v, f, 32);
    fe_sq(c, a);
    memcpy(out, c.v, 32);
}
void hh_fe_mul_school_limbs(u32* out, const u32* f, const u32* g) {
    fe a, b, c;
    memcpy(a.v, f, 32);
    memcpy(b.v, g, 32);
    fe_mul_school(c, a, b);
    memcpy(out, c.v, 32);
}
void hh_fe_addsub_limbs(u32* sum, u32* diff, u32* neg, const u32* f, const u32* g) {
    fe a, b, c;
    memcpy(a.v, f, 32);
    memcpy(b.v, g, 32);
    fe_add(c, a, b);
    memcpy(sum, c.v, 32);
    fe_sub(c, a, b);
    memcpy(diff, c.v, 32);
    fe_neg(c, a);
    memcpy(neg, c.v, 32);
}
void hh_fe_tobytes_limbs(uint8_t* out, const u32* f) {
    fe a;
    memcpy(a.v, f, 32);
    u32 w[8];
    fe_towords(w, a);
    store_words(out, w);
}
void hh_fe_frombytes(u32* out, const uint8_t* in) {
    u32 w[8];
    load_words(w, in);
    fe a;
    fe_fromwords(a, w);
    memcpy(out, a.v, 32);
}
void hh_fe_const(int which, uint8_t* out) {
    fe c = which == 0 ? fe_d() : which == 1 ? fe_2d() : which == 2 ? fe_sqrt_m1() : which == 3 ? fe_invsqrt_a_minus_d()
         : which == 4 ? fe_sqrt_ad_minus_one() : which == 5 ? fe_one_minus_d_sq() : fe_d_minus_one_sq();
    u32 w[8];
    fe_towords(w, c);
    store_words(out, w);
}
int hh_sqrt_ratio_i(uint8_t* out, const uint8_t* u, const uint8_t* v) {
    u32 w[8];
    fe fu, fv, r;
    load_words(w, u);
    fe_fromwords(fu, w);
    load_words(w, v);
    fe_fromwords(fv, w);
    int ok = (int)fe_sqrt_ratio_i(r, fu, fv);
    fe_towords(w, r);
    store_words(out, w);
    return ok;
}
void hh_fe_invert(uint8_t* out, const uint8_t* z) {
    u32 w[8];
    fe a, r;
    load_words(w, z);
    fe_fromwords(a, w);
    fe_invert(r, a);
    fe_towords(w, r);
    store_words(out, w);
}
// decompress: returns ok; xyzt = 4 x 32 canonical bytes
int hh_decompress(uint8_t* xyzt, const uint8_t* in) {
    u32 w[8];
    load_words(w, in);
    ge_p3 p;
    int ok = (int)ristretto_decompress(p, w);
    fe_towords(w, p.X); store_words(xyzt, w);
    fe_towords(w, p.Y); store_words(xyzt + 32, w);
    fe_towords(w, p.Z); store_words(xyzt + 64, w);
    fe_towords(w, p.T); store_words(xyzt + 96, w);
    return ok;
}
static void p3_from_bytes(ge_p3& p, const uint8_t* xyzt) {
    u32 w[8];
    load_words(w, xyzt); fe_fromwords(p.X, w);
    load_words(w, xyzt + 32); fe_fromwords(p.Y, w);
    load_words(w, xyzt + 64); fe_fromwords(p.Z, w);
    load_words(w, xyzt + 96); fe_fromwords(p.T, w);
}
void hh_compress_xyzt(uint8_t* out, const uint8_t* xyzt) {
    ge_p3 p;
    p3_from_bytes(p, xyzt);
    u32 w[8];
    ristretto_compress(w, p);
    store_words(out, w);
}
int hh_add(uint8_t* out, const uint8_t* a, const uint8_t* b) {
    u32 w[8];
    ge_p3 p, q, r;
    load_words(w, a);
    int ok = (int)ristretto_decompress(p, w);
    load_words(w, b);
    ok &= (int)ristretto_decompress(q, w);
    ge_cached c;
    ge_to_cached(c, q);
    ge_add(r, p, c);
    ristretto_compress(w, r);
    store_words(out, w);
    return ok;
}
int hh_sub_madd(uint8_t* out, const uint8_t* a, const uint8_t* b) {  // a - b via negated affine Niels
    u32 w[8];
    ge_p3 p, q, r;
    load_words(w, a);
    int ok = (int)ristretto_decompress(p, w);
    load_words(w, b);
    ok &= (int)ristretto_decompress(q, w);
    ge_niels n;
    ge_to_niels_z1(n, q);
    ge_niels_cneg(n, 1);
    ge_madd(r, p, n);
    ristretto_compress(w, r);
    store_words(out, w);
    return ok;
}
int hh_dbl(uint8_t* out, const uint8_t* a, int n) {
    u32 w[8];
    ge_p3 p;
    load_words(w, a);
    int ok = (int)ristretto_decompress(p, w);
    for (int i = 0; i < n; i++) {
        if (i == n - 1) ge_dbl<true>(p, p);
        else ge_dbl<false>(p, p);
    }
    ristretto_compress(w, p);
    store_words(out, w);
    return ok;
}
int hh_eq(const uint8_t* a, const uint8_t* b) {
    u32 w[8];
    ge_p3 p, q;
    load_words(w, a);
    ristretto_decompress(p, w);
    load_words(w, b);
    ristretto_decompress(q, w);
    return (int)ge_ristretto_eq(p, q) | ((int)ge_ristretto_is_identity(p) << 1);
}
int hh_scalarmult(uint8_t* out, const uint8_t* scalar, const uint8_t* point) {
    u32 w[8], s[8];
    ge_p3 p, r;
    load_words(w, point);
    int ok = (int)ristretto_decompress(p, w);
    load_words(s, scalar);
    std::vector<u32x4> tbl(QQ_VB_ENTRIES * QQ_PT_Q);
    vb_build_table(tbl.data(), p);
    vb_scalarmult(r, tbl.data(), s);
    ristretto_compress(w, r);
    store_words(out, w);
    return ok | ((int)sc_is_canonical(s) << 1);
}
// split variable base (two scalars, one point): out0 = s0 * P, out1 = s1 * P
int hh_scalarmult_split(uint8_t* out0, uint8_t* out1, const uint8_t* s0, const uint8_t* s1, const uint8_t* point) {
    u32 w[8], s[8];
    ge_p3 p, r;
    load_words(w, point);
    int ok = (int)ristretto_decompress(p, w);
    std::vector<u32x4> tbl(QQ_VBS_TABLE_Q);
    vbs_build_tables(tbl.data(), p);
    load_words(s, s0);
    vbs_scalarmult(r, tbl.data(), s);
    ristretto_compress(w, r);
    store_words(out0, w);
    load_words(s, s1);
    vbs_scalarmult(r, tbl.data(), s);
    ristretto_compress(w, r);
    store_words(out1, w);
    return ok;
}
// out = enc(s * P) computed as enc(2 * ((s / 2 mod l) * P)) through the batch encoder's per-item functions
// (the inversion that the GPU shares across a batch is a plain fe_invert here); zscale rescales the projective
// representative first.  Returns the zero flag (1 when the result is the identity class).
int hh_halve_dblcompress(uint8_t* out, const uint8_t* scalar, const uint8_t* point, const uint8_t* zscale) {
    u32 w[8], s[8], h[8];
    ge_p3 p, r;
    load_words(w, point);
    ristretto_decompress(p, w);
    load_words(s, scalar);
    sc_halve(h, s);
    std::vector<u32x4> tbl(QQ_VB_ENTRIES * QQ_PT_Q);
    vb_build_table(tbl.data(), p);
    vb_scalarmult(r, tbl.data(), h);
    fe z;
    load_words(w, zscale);
    fe_fromwords(z, w);
    fe_mul(r.X, r.X, z); fe_mul(r.Y, r.Y, z); fe_mul(r.Z, r.Z, z); fe_mul(r.T, r.T, z);
    dc_state st;
    fe wv, inv;
    dc_prepare(st, wv, r);
    int zero = (int)fe_iszero(wv);
    fe_invert(inv, wv);
    dc_finish(w, st, inv);
    if (zero) memset(w, 0, 32);
    store_words(out, w);
    return zero;
}
void hh_sc_halve(uint8_t* out, const uint8_t* scalar) {
    u32 s[8], h[8];
    load_words(s, scalar);
    sc_halve(h, s);
    store_words(out, h);
}
void hh_from_uniform(uint8_t* out, const uint8_t* in64) {
    u32 w[16], o[8];
    memcpy(w, in64, 64);
    ge_p3 p;
    ristretto_from_uniform(p, w);
    ristretto_compress(o, p);
    store_words(out, o);
}
// fixed base: W in {4,5,6,8}; builds the table on each call into caller-provided buffer (words)
size_t hh_fb_table_words(int W) { return (size_t)fb_num_windows(W) * fb_entries(W) * QQ_NIELS_WORDS; }
int hh_fb_build(u32* tbl, int W, const uint8_t* base) {
    u32 w[8];
    ge_p3 p;
    load_words(w, base);
    int ok = (int)ristretto_decompress(p, w);
    for (int k = 0; k < fb_num_windows(W); k++)
        for (int j = 0; j < fb_entries(W); j++)
            fb_build_entry(tbl + ((size_t)k * fb_entries(W) + j) * QQ_NIELS_WORDS, p, W, k, j);
    return ok;
}
void hh_fb_mult(uint8_t* out, const u32* tbl, int W, const uint8_t* scalar) {
    u32 s[8], w[8];
    load_words(s, scalar);
    ge_p3 r;
    if (W == 4) fb_scalarmult<4>(r, tbl, s);
    else if (W == 5) fb_scalarmult<5>(r, tbl, s);
    else if (W == 6) fb_scalarmult<6>(r, tbl, s);
    else fb_scalarmult<8>(r, tbl, s);
    ristretto_compress(w, r);
    store_words(out, w);
}
// host scalar field (sc_host.hpp): op 0 add, 1 sub, 2 mul, 3 invert(a), 4 from_wide(a || b), 5 invert_vartime, 6 mul in the
// 32-bit limb form of the device code, 7 its wide reduction; returns 0 when an input is not canonical
int hh_sc_op(uint8_t* out, int op, const uint8_t* a, const uint8_t* b) {
    qq_sc::sc x, y, r;
    if (op == 4 || op == 7) {
        uint8_t w[64];
        memcpy(w, a, 32);
        memcpy(w + 32, b, 32);
        if (op == 4) {
            r = qq_sc::from_wide(w);
        } else {                      // the 32-bit limb reduction the device code runs
            uint32_t x32[16];
            memcpy(x32, w, 64);
            r = qq_sc::reduce512_w32(x32);
        }
    } else {
        if (!qq_sc::from_bytes(x, a) || !qq_sc::from_bytes(y, b)) return 0;
        r = op == 0 ? qq_sc::add(x, y) : op == 1 ? qq_sc::sub(x, y) : op == 2 ? qq_sc::mul(x, y) : op == 3 ? qq_sc::invert(x)
          : op == 6 ? qq_sc::mul_w32(x, y) : op == 8 ? qq_sc::invert_fixed(x) : qq_sc::invert_vartime(x);
    }
    qq_sc::to_bytes(out, r);
    return 1;
}
// host Keccak (keccak_host.hpp): SHA3-512 and SHAKE256 of one message
void hh_sha3_512(uint8_t* out64, const uint8_t* data, size_t len) { qq_keccak::sha3_512(data, len, out64); }
void hh_shake256(uint8_t* out, size_t outlen, const uint8_t* data, size_t len) {
    qq_keccak::shake256 x;
    x.absorb(data, len);
    x.squeeze(out, outlen);
}
// host Merlin (merlin_host.hpp): Transcript::new(label); a script of operations; the last challenge goes to out.
// script = records of: kind (0 append_message, 1 challenge_bytes) | label length (1 B) | label | data length (2 B LE) | data
// (for kind 1 the "data length" is the number of challenge bytes and no data follows); out must hold the largest challenge.
void hh_merlin_script(uint8_t* out, const uint8_t* label, size_t label_len, const uint8_t* script, size_t script_len) {
    qq_merlin::transcript tr(label, label_len);
    size_t i = 0;
    while (i < script_len) {
        int kind = script[i++];
        size_t ll = script[i++];
        std::string lab((const char*)script + i, ll);
        i += ll;
        size_t dl = script[i] | ((size_t)script[i + 1] << 8);
        i += 2;
        if (kind == 0) {
            tr.append_message(lab.c_str(), script + i, dl);
            i += dl;
        } else {
            tr.challenge_bytes(lab.c_str(), out, dl);
        }
    }
}
// One segmented MSM batch evaluated on the host with the device arithmetic (what decompress + k_straus + the encoder compute):
// out[m] = enc(sum s_i dec(P_i)), status 2 for a non-canonical scalar, 1 for an undecodable point (output zeroed then).
static void hh_segmented(const uint8_t* sc, const uint8_t* pt, const uint32_t* first, size_t msms, uint8_t* e, uint8_t* st) {
    std::vector<u32x4> tbl(QQ_VB_ENTRIES * QQ_PT_Q);
    for (size_t m = 0; m < msms; m++) {
        ge_p3 acc;
        ge_identity(acc);
        uint8_t s = 0;
        for (uint32_t t = first[m]; t < first[m + 1]; t++) {
            u32 w[8], k[8];
            load_words(w, pt + 32 * t);
            load_words(k, sc + 32 * t);
            ge_p3 P, R;
            u32 ok = ristretto_decompress(P, w);
            uint8_t ts = !sc_is_canonical(k) ? 2 : (ok ? 0 : 1);
            s = (ts == 2 || s == 2) ? 2 : (s | ts);
            if (ts) continue;
            vb_build_table(tbl.data(), P);
            vb_scalarmult(R, tbl.data(), k);
            ge_cached c;
            ge_to_cached(c, R);
            ge_add(acc, acc, c);
        }
        u32 w[8];
        ristretto_compress(w, acc);
        store_words(e + 32 * m, w);
        if (s) memset(e + 32 * m, 0, 32);
        st[m] = s;
    }
}
// ShuffleProof::verify through the per-proof phases of shuffle_verify.cuh (the code the transcript kernels run), MSM batches by
// hh_segmented: checks the shared phase logic without a GPU.  xpc: H | G[0..3) of VectorPedersenGens::new(4); base_pk: B | H_p.
void hh_shuffle_verify(const char* transcript_label, const char* verifier_label, const uint8_t* in, const uint8_t* out,
                       const uint8_t* stm, const uint8_t* proof, size_t n, const uint8_t* base_pk, const uint8_t* xpc,
                       uint8_t* status, uint8_t* stage, uint8_t* detail) {
    using namespace qq_shuffle;
    qq_merlin::transcript tr0((const uint8_t*)transcript_label, strlen(transcript_label));
    tr0.domain_sep(verifier_label);
    gens g{base_pk, base_pk + 32, xpc, xpc + 32};
    const uint32_t sz1[QQ_SHUFFLE_MSMS_1] = {8, 8, 8, 8, 5, 5, 5, 3, 1, 7, 8, 9, 6, 5, 9, 9, 10, 10};
    const uint32_t sz2[QQ_SHUFFLE_MSMS_2] = {8, 8, 8, 8, 8, 8, 9, 9, 8, 8, 8, 8, 8, 9};
    for (size_t p = 0; p < n; p++) {
        std::vector<uint8_t> sc1(QQ_SHUFFLE_TERMS_1 * 32, 0), pt1(QQ_SHUFFLE_TERMS_1 * 32), sc2(QQ_SHUFFLE_TERMS_2 * 32, 0), pt2(QQ_SHUFFLE_TERMS_2 * 32);
        for (size_t t = 0; t < QQ_SHUFFLE_TERMS_1; t++) memcpy(&pt1[32 * t], base_pk, 32);
        for (size_t t = 0; t < QQ_SHUFFLE_TERMS_2; t++) memcpy(&pt2[32 * t], base_pk, 32);
        job_sink j1, j2;
        memset((void*)&j1, 0, sizeof j1);
        memset((void*)&j2, 0, sizeof j2);
        for (int m = 0; m < QQ_JOB_MAX_MSMS; m++) j1.exact_slot[m] = j2.exact_slot[m] = -1;
        j1.sc = sc1.data(); j1.pt = pt1.data(); j1.msms_pp = QQ_SHUFFLE_MSMS_1; j1.terms_pp = QQ_SHUFFLE_TERMS_1; j1.base = p;
        j2.sc = sc2.data(); j2.pt = pt2.data(); j2.msms_pp = QQ_SHUFFLE_MSMS_2; j2.terms_pp = QQ_SHUFFLE_TERMS_2; j2.base = p;
        uint32_t t = 0;
        for (int m = 0; m < QQ_SHUFFLE_MSMS_1; m++) { j1.first[m] = t; t += sz1[m]; }
        j1.first[QQ_SHUFFLE_MSMS_1] = t;
        t = 0;
        for (int m = 0; m < QQ_SHUFFLE_MSMS_2; m++) { j2.first[m] = t; t += sz2[m]; }
        j2.first[QQ_SHUFFLE_MSMS_2] = t;
        const uint8_t *pr = proof + QQ_SHUFFLE_PROOF_BYTES * p, *sm = stm + QQ_SHUFFLE_STATEMENT_BYTES * p;
        proof_state S(tr0);
        pass_a(S, j1, p, pr, sm, in + 1152 * p, g);
        uint8_t e1[QQ_SHUFFLE_MSMS_1 * 32], s1[QQ_SHUFFLE_MSMS_1], e2[QQ_SHUFFLE_MSMS_2 * 32], s2[QQ_SHUFFLE_MSMS_2];
        hh_segmented(sc1.data(), pt1.data(), j1.first, QQ_SHUFFLE_MSMS_1, e1, s1);
        pass_b(S, j2, p, pr, sm, in + 1152 * p, out + 1152 * p, e1, s1, e1 + 32 * 14, s1 + 14, g);
        hh_segmented(sc2.data(), pt2.data(), j2.first, QQ_SHUFFLE_MSMS_2, e2, s2);
        pass_final(S, pr, e2, s2);
        status[p] = S.st;
        stage[p] = S.sg;
        detail[p] = S.dt;
    }
}
// Aggregate form of the same verification (the fast path of qq_verify_shuffle_batch): per proof the exact MSMs g_r, h_r,
// every other group equation weighted into ONE sum over all proofs, which must be the identity.  clean[p]: the proof passed
// every scalar check and is inside the aggregate; *agg_identity: the aggregated MSM (clean proofs only) is the identity;
// counts[2 p], counts[2 p + 1]: aggregated (non-fixed) terms the proof emitted in batch 1 / 2.
static uint8_t* g_apt_dump = nullptr;      // test hook: the aggregated MSM's point list of the next hh_shuffle_verify_aggregate call
void hh_set_apt_dump(uint8_t* p) { g_apt_dump = p; }
void hh_shuffle_verify_aggregate(const char* transcript_label, const char* verifier_label, const uint8_t* in, const uint8_t* out,
                                 const uint8_t* stm, const uint8_t* proof, size_t n, const uint8_t* base_pk, const uint8_t* xpc,
                                 const uint8_t* entropy, uint8_t* clean, uint8_t* agg_identity, uint32_t* counts) {
    using namespace qq_shuffle;
    qq_merlin::transcript tr0((const uint8_t*)transcript_label, strlen(transcript_label));
    tr0.domain_sep(verifier_label);
    gens g{base_pk, base_pk + 32, xpc, xpc + 32};
    const size_t C1 = QQ_SHUFFLE_AGG_CAP_1, C2 = QQ_SHUFFLE_AGG_CAP_2, N = n * (C1 + C2) + 6;
    std::vector<uint8_t> asc(N * 32, 0), apt(N * 32);
    for (size_t t = 0; t < N; t++) memcpy(&apt[32 * t], base_pk, 32);
    qq_sc::sc fixed[6];
    for (int i = 0; i < 6; i++) fixed[i] = qq_sc::zero();
    for (size_t p = 0; p < n; p++) {
        std::vector<uint8_t> xsc(4 * 32, 0), xpt(4 * 32);
        for (size_t t = 0; t < 4; t++) memcpy(&xpt[32 * t], base_pk, 32);
        job_sink j1, j2;
        memset((void*)&j1, 0, sizeof j1);
        memset((void*)&j2, 0, sizeof j2);
        for (int m = 0; m < QQ_JOB_MAX_MSMS; m++) j1.exact_slot[m] = j2.exact_slot[m] = -1;
        const uint32_t xf[3] = {0, 2, 4};
        for (int m = 0; m < 3; m++) j1.first[m] = xf[m];
        j1.sc = xsc.data(); j1.pt = xpt.data(); j1.msms_pp = 2; j1.terms_pp = 4; j1.base = p;
        j1.exact_slot[16] = 0;
        j1.exact_slot[17] = 1;
        j1.asc = asc.data() + 32 * C1 * p; j1.apt = apt.data() + 32 * C1 * p; j1.cap = (uint32_t)C1;
        j2.base = p;
        j2.asc = asc.data() + 32 * (C1 * n + C2 * p); j2.apt = apt.data() + 32 * (C1 * n + C2 * p); j2.cap = (uint32_t)C2;
        j1.fB = j2.fB = g.B; j1.fHp = j2.fHp = g.Hp; j1.fH = j2.fH = g.H; j1.fG = j2.fG = g.G;
        const uint8_t *pr = proof + QQ_SHUFFLE_PROOF_BYTES * p, *sm = stm + QQ_SHUFFLE_STATEMENT_BYTES * p;
        agg_ctx A;
        agg_begin_a(A, entropy, p);
        j1.agg = &A;
        proof_state S(tr0);
        pass_a(S, j1, p, pr, sm, in + 1152 * p, g);
        counts[2 * p] = A.k;
        qq_sc::sc fa[6];
        for (int i = 0; i < 6; i++) fa[i] = A.fixed[i];
        bool ok = !A.overflow;
        uint8_t eG[128], sG[4] = {0, 0, 0, 0};
        memcpy(eG, pr + 2048 + 224 + 96, 32);          // G, H: the proof's own encodings (checked inside the aggregate)
        memcpy(eG + 32, pr + 2048 + 416 + 96, 32);
        hh_segmented(xsc.data(), xpt.data(), j1.first, 2, eG + 64, sG + 2);
        agg_begin_b(A, entropy, p);
        j2.agg = &A;
        if (ok) ok = pass_b(S, j2, p, pr, sm, in + 1152 * p, out + 1152 * p, nullptr, nullptr, eG, sG, g) && !A.overflow;
        counts[2 * p + 1] = A.k;
        clean[p] = ok ? 1 : 0;
        if (ok) {
            for (int i = 0; i < 6; i++) fixed[i] = qq_sc::add(fixed[i], qq_sc::add(fa[i], A.fixed[i]));
        } else {
            memset(j1.asc, 0, 32 * C1);
            memset(j2.asc, 0, 32 * C2);
        }
    }
    for (int i = 0; i < 6; i++) {
        const uint8_t* pt = i == 0 ? g.B : i == 1 ? g.Hp : i == 2 ? g.H : g.G + 32 * (i - 3);
        job_sink::put(asc.data(), apt.data(), n * (C1 + C2) + i, fixed[i], pt);
    }
    if (g_apt_dump) memcpy(g_apt_dump, apt.data(), N * 32);
    uint32_t first[2] = {0, (uint32_t)N};
    uint8_t e[32], st;
    hh_segmented(asc.data(), apt.data(), first, 1, e, &st);
    *agg_identity = (st == 0 && is_zero32(e)) ? 1 : 0;
}
size_t hh_transcript_state_bytes() { return qq_merlin::transcript::STATE_BYTES; }
// export -> import round trip of a transcript state, then one more challenge from both; returns 1 when they agree and the
// import rejects a corrupted tag / position
int hh_transcript_state_roundtrip(const uint8_t* label, size_t label_len, uint8_t* state_out) {
    qq_merlin::transcript a(label, label_len);
    a.domain_sep("state-test");
    uint8_t st[qq_merlin::transcript::STATE_BYTES];
    a.export_state(st);
    memcpy(state_out, st, sizeof st);
    qq_merlin::transcript b((const uint8_t*)"other", 5);
    if (!b.import_state(st)) return 0;
    uint8_t ca[32], cb[32];
    a.get_challenge("c", ca);
    b.get_challenge("c", cb);
    if (memcmp(ca, cb, 32)) return 0;
    uint8_t bad[qq_merlin::transcript::STATE_BYTES];
    memcpy(bad, st, sizeof st);
    bad[203] = 0;
    if (b.import_state(bad)) return 0;
    memcpy(bad, st, sizeof st);
    bad[200] = 166;
    if (b.import_state(bad)) return 0;
    memset(bad, 0, sizeof bad);
    if (b.import_state(bad)) return 0;
    return 1;
}

// ---- FP64-pipe field / group arithmetic (fe64.cuh) ---------------------------------------------------------------------
// raw limbs: x_i = q_i * 2^o_i (signed integers q_i chosen by the test, so worst-case magnitudes can be driven);
// op 0: a * b, 1: a^2, 2: 2 a^2.  The result comes back in the saturated integer form.
void hh_fe64_op_raw(u32* out, const long long* qa, const long long* qb, int op) {
    fe64 a, b, h;
    for (int i = 0; i < QQ_F64_LIMBS; i++) {
        a.v[i] = (double)qa[i] * f64_p2(f64_off(i));
        b.v[i] = (double)qb[i] * f64_p2(f64_off(i));
    }
    if (op == 0) fe64_mul(h, a, b);
    else if (op == 1) fe64_sq<false>(h, a);
    else fe64_sq<true>(h, a);
    // the product must come out carried: |x_i| <= 2^(o_(i+1) - 1) (+ the tiny second-lap excess on limb 2)
    for (int i = 0; i < QQ_F64_LIMBS; i++) {
        double lim = f64_p2(f64_off(i + 1) - 1) * (i == 2 ? 1.001 : 1.0);
        if (!(h.v[i] <= lim && h.v[i] >= -lim)) { memset(out, 0xff, 32); return; }
    }
    fe r;
    fe64_to_fe(r, h);
    memcpy(out, r.v, 32);
}
// saturated form -> fe64 -> (sum of `terms` copies, alternating nothing) -> saturated form
void hh_fe64_roundtrip(u32* out, const u32* in, int terms) {
    fe f, r;
    memcpy(f.v, in, 32);
    fe64 x, acc;
    fe64_from_fe(x, f);
    acc = x;
    for (int i = 1; i < terms; i++) fe64_add(acc, acc, x);
    fe64_to_fe(r, acc);
    memcpy(out, r.v, 32);
}
int hh_scalarmult_split64(uint8_t* out0, uint8_t* out1, const uint8_t* s0, const uint8_t* s1, const uint8_t* point) {
    u32 w[8], s[8];
    ge_p3 p, r;
    load_words(w, point);
    int ok = (int)ristretto_decompress(p, w);
    std::vector<double> tbl(QQ_VBS64_TABLE_D);
    ge64_p3 p64, r64;
    fe64 d2;
    fe64_from_fe(d2, fe_2d());
    ge64_from_p3(p64, p);
    vbs64_build_tables(tbl.data(), p64, d2);
    load_words(s, s0);
    vbs64_scalarmult(r64, tbl.data(), s);
    ge64_to_p3(r, r64);
    ristretto_compress(w, r);
    store_words(out0, w);
    load_words(s, s1);
    vbs64_scalarmult(r64, tbl.data(), s);
    ge64_to_p3(r, r64);
    ristretto_compress(w, r);
    store_words(out1, w);
    return ok;
}

// lazy sums of products (sc_host.hpp: mul_wide_w32 / acc17_add / acc17_reduce) and the one-reduction difference shl_minus_wide:
// out_sum = sum_p a_p b_p mod l over n pairs of canonical 32-byte scalars (each pair added `repeat` times, to drive the 17th limb);
// out_u = (rzz 2^k - sb slo) mod l
void hh_sc_lazy(uint8_t* out_sum, uint8_t* out_u, const uint8_t* a, const uint8_t* b, size_t n, unsigned repeat, const uint8_t* rzz, int k,
                const uint8_t* sb, const uint8_t* slo) {
    uint32_t acc[17] = {0}, x[16];
    for (size_t p = 0; p < n; p++) {
        qq_sc::sc x1, x2;
        qq_sc::from_bytes(x1, a + 32 * p);
        qq_sc::from_bytes(x2, b + 32 * p);
        qq_sc::mul_wide_w32(x, x1, x2);
        for (unsigned r = 0; r < repeat; r++) qq_sc::acc17_add(acc, x);
    }
    qq_sc::to_bytes(out_sum, qq_sc::acc17_reduce(acc));
    qq_sc::sc r1, s1, s2;
    qq_sc::from_bytes(r1, rzz);
    qq_sc::from_bytes(s1, sb);
    qq_sc::from_bytes(s2, slo);
    qq_sc::mul_wide_w32(x, s1, s2);
    qq_sc::to_bytes(out_u, qq_sc::shl_minus_wide(r1, k, x));
}
}
