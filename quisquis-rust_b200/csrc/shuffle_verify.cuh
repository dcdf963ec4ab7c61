// Per-proof phases of the batched Bayer-Groth shuffle verifier (reference src/shuffle/*.rs, ROWS = COLUMNS = 3), written once
// for host and device: Merlin transcript + Z/l algebra of ONE proof, emitting every group equation as an MSM job (scalar,
// compressed point) into fixed slots.  qq_api_shuffle.inc runs them
//   * one GPU thread per proof (k_shuffle_pass_a / _pass_b / _final below): the proof bytes are uploaded once, the job lists,
//     transcripts and verdicts never leave the device (qq_verify_shuffle_batch), or
//   * on the host threads for the stand-alone leaf arguments (qq_verify_{svp,hadamard,product}_batch) and, as a measurement
//     knob (qq_verify_set_transcripts), for the whole shuffle proof.
// The MSMs themselves always run in the segmented-MSM kernels (kernels.cuh: k_straus / k_straus_coop).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "merlin_host.hpp"
#include "sc_host.hpp"

#ifndef QQ_ST_OK
#define QQ_ST_OK 0
#define QQ_ST_BAD_POINT 1
#define QQ_ST_BAD_SCALAR 2
#define QQ_ST_PROOF 6
#endif

#define QQ_SVP_PROOF_BYTES 352
#define QQ_HADAMARD_PROOF_BYTES 640
#define QQ_PRODUCT_PROOF_BYTES 1024
#define QQ_PRODUCT_STATEMENT_BYTES 192
#define QQ_MEXP_PROOF_BYTES 832
#define QQ_SHUFFLE_PROOF_BYTES 3776
#define QQ_SHUFFLE_STATEMENT_BYTES 352
#define QQ_SHUFFLE_MSMS_1 18       // batch 1: 0-3 Hadamard, 4-6 c_E, 7-13 product, 14 G, 15 H, 16 g_r, 17 h_r
#define QQ_SHUFFLE_TERMS_1 124
#define QQ_SHUFFLE_MSMS_2 14       // batch 2: pubkey a, b, E_K g (2), E_K h (2); commitment C_c, C_d, a, b, E_K c (2), E_K d (2)
#define QQ_SHUFFLE_TERMS_2 115

namespace qq_shuffle {

using qq_sc::sc;

// the fixed generators every phase needs, as 32-byte encodings in the memory space the phase runs in:
// B, Hp = BASE_PK_BTC_COMPRESSED (PedersenGens::default()); H, G[0..3) = VectorPedersenGens::new(4)
struct gens {
    const uint8_t *B, *Hp, *H, *G;
};

// Where one batch's MSM jobs go: every proof owns the same slots (MSM m = terms first[m] .. first[m + 1] - 1 of its
// terms_pp), proof-major; sc / pt are host or device arrays of 32-byte entries.
//
// Aggregate mode (agg != nullptr; qq_verify_shuffle_batch's fast path): only the MSMs whose RESULT is needed (G, H, g_r,
// h_r: they enter the transcript) keep a slot of their own - exact_slot[m] >= 0 names it in the (then smaller) sc / pt list.
// Every other MSM is an equation "sum == identity": its terms are multiplied by the random 128-bit weight w[m] of that check
// and appended to the proof's segment of ONE aggregated term list (asc / apt, cap entries per proof) that a single Pippenger
// MSM over all proofs evaluates; terms on the six fixed generators are summed into fixed[] instead (one term per generator
// for the whole batch).  Slots in skip_mask are not emitted (decode-only checks: their points occur in other equations).
#define QQ_JOB_MAX_MSMS 20
struct agg_ctx {                // per proof, lives in the frame of the phase that emits
    qq_sc::sc w[QQ_JOB_MAX_MSMS];
    qq_sc::sc fixed[6];         // B, Hp, H, G[0], G[1], G[2]
    uint32_t k;                 // next entry of the proof's segment
    uint32_t skip_mask;
    bool overflow;
};
struct job_sink {
    uint8_t *sc, *pt;
    uint32_t first[QQ_JOB_MAX_MSMS + 1];
    uint32_t msms_pp, terms_pp;
    size_t base;      // global index of the first proof of this batch (slices of a larger call)
    // aggregate mode
    agg_ctx* agg;
    uint8_t *asc, *apt;
    uint32_t cap;
    int8_t exact_slot[QQ_JOB_MAX_MSMS];
    const uint8_t *fB, *fHp, *fH, *fG;      // the fixed generators' addresses (gens), compared by pointer
    QQ_HOSTDEV static void put(uint8_t* sc_arr, uint8_t* pt_arr, size_t i, const qq_sc::sc& s, const uint8_t* point) {
#ifdef __CUDA_ARCH__
        // the job arrays are 32-byte aligned device buffers; so are the points of the uploaded proofs (every struct size
        // is a multiple of 32) - the generators and MSM outputs as well.  16-byte accesses instead of 64 byte moves.
        // Both halves of the point are loaded before anything is stored: through generic pointers the compiler must keep a
        // load behind every earlier store, and each load is a trip to L2 / HBM.
        uint4* ds = reinterpret_cast<uint4*>(&sc_arr[32 * i]);
        uint4* dp = reinterpret_cast<uint4*>(&pt_arr[32 * i]);
        if ((reinterpret_cast<uintptr_t>(point) & 15) == 0) {
            const uint4* sp = reinterpret_cast<const uint4*>(point);
            const uint4 p0 = sp[0], p1 = sp[1];
            ds[0] = make_uint4((uint32_t)s.v[0], (uint32_t)(s.v[0] >> 32), (uint32_t)s.v[1], (uint32_t)(s.v[1] >> 32));
            ds[1] = make_uint4((uint32_t)s.v[2], (uint32_t)(s.v[2] >> 32), (uint32_t)s.v[3], (uint32_t)(s.v[3] >> 32));
            dp[0] = p0;
            dp[1] = p1;
        } else {
            ds[0] = make_uint4((uint32_t)s.v[0], (uint32_t)(s.v[0] >> 32), (uint32_t)s.v[1], (uint32_t)(s.v[1] >> 32));
            ds[1] = make_uint4((uint32_t)s.v[2], (uint32_t)(s.v[2] >> 32), (uint32_t)s.v[3], (uint32_t)(s.v[3] >> 32));
            memcpy(&pt_arr[32 * i], point, 32);
        }
#else
        qq_sc::to_bytes(&sc_arr[32 * i], s);
        memcpy(&pt_arr[32 * i], point, 32);
#endif
    }
    QQ_HOSTDEV QQ_NOINLINE void set(size_t p, size_t m, size_t t, const qq_sc::sc& s, const uint8_t* point) const {
        if (agg == nullptr) {
            put(sc, pt, (p - base) * terms_pp + first[m] + t, s, point);
            return;
        }
        if (exact_slot[m] >= 0) {
            put(sc, pt, (p - base) * terms_pp + first[exact_slot[m]] + t, s, point);
            return;
        }
        if ((agg->skip_mask >> m) & 1u) return;
        qq_sc::sc ws = qq_sc::mul(agg->w[m], s);
        int f = point == fB ? 0 : point == fHp ? 1 : point == fH ? 2 : (point == fG ? 3 : point == fG + 32 ? 4 : point == fG + 64 ? 5 : -1);
        if (f >= 0) {
            agg->fixed[f] = qq_sc::add(agg->fixed[f], ws);
            return;
        }
        if (agg->k >= cap) {
            agg->overflow = true;
            return;
        }
        put(asc, apt, (p - base) * cap + agg->k, ws, point);
        agg->k++;
    }
};
#define QQ_SHUFFLE_AGG_CAP_1 64      // aggregated (non-fixed) terms of a proof that passes every scalar check: batch 1
#define QQ_SHUFFLE_AGG_CAP_2 102     // and batch 2

// 128-bit weights of the aggregated checks: Keccak-f over (fresh 32-byte entropy of this call, proof index, batch tag,
// counter) - unpredictable to whoever made the proofs, which is all a random linear combination needs.
QQ_HOSTDEV static inline void agg_weights(qq_sc::sc* w, int n, const uint8_t entropy[32], uint64_t p, uint64_t tag) {
    int got = 0;
    for (uint64_t ctr = 0; got < n; ctr++) {
        uint64_t st[25];
        for (int i = 0; i < 25; i++) st[i] = 0;
        memcpy(st, entropy, 32);
        st[4] = p;
        st[5] = tag;
        st[6] = ctr;
        st[7] = 0x71715f6232303061ULL;      // "qq_b200a": domain tag of this generator
        qq_keccak::f1600(st);
        for (int i = 0; i + 1 < 24 && got < n; i += 2) {
            w[got].v[0] = st[i] | 1ULL;      // never zero
            w[got].v[1] = st[i + 1];
            w[got].v[2] = 0;
            w[got].v[3] = 0;
            got++;
        }
    }
}

QQ_HOSTDEV static inline bool is_zero32(const uint8_t* p) {
    uint8_t acc = 0;
    for (int i = 0; i < 32; i++) acc |= p[i];
    return acc == 0;
}
QQ_HOSTDEV static inline bool differ32(const uint8_t* a, const uint8_t* b) {
    uint8_t acc = 0;
    for (int i = 0; i < 32; i++) acc |= (uint8_t)(a[i] ^ b[i]);
    return acc != 0;
}
QQ_HOSTDEV static inline void transcript_challenge(qq_merlin::transcript& tr, const char* label, sc& out) {
    uint8_t wide[64];
    tr.challenge_bytes(label, wide, 64);
    out = qq_sc::from_wide(wide);
}

// SVPProof::verify (reference src/shuffle/singlevalueproduct.rs:175-257) on a transcript that already carries everything
// before it: scalar checks, challenge, and the two identity MSMs into slots (m0, m0 + 1) of proof p.
//   a~_1 == b~_1;   x = H(transcript);   b x == b~_3;
//   x c_a + c_d == com(a~; r~)                                  [MSM m0:     x c_a + c_d - r~ H - sum a~_i G_i == 0]
//   x c_Delta + c_delta == com(x b~_{i+1} - b~_i a~_{i+1}; s~)  [MSM m0 + 1, two generators]
// Returns QQ_ST_OK / QQ_ST_BAD_SCALAR / QQ_ST_PROOF.
QQ_HOSTDEV static inline uint8_t svp_phase(qq_merlin::transcript& tr, const uint8_t* pr, const uint8_t* commitment_a,
                                           const uint8_t* b, const gens& g, const job_sink& jobs, size_t p, size_t m0) {
    using namespace qq_sc;
    sc at[3], bt[3], rt, st_, bb;
    bool canon = from_bytes(bb, b) && from_bytes(rt, pr + 288) && from_bytes(st_, pr + 320);
    for (int i = 0; i < 3; i++) canon = canon && from_bytes(at[i], pr + 96 + 32 * i) && from_bytes(bt[i], pr + 192 + 32 * i);
    if (!canon) return QQ_ST_BAD_SCALAR;
    if (at[0] != bt[0]) return QQ_ST_PROOF;
    tr.domain_sep("SingleValueProductProof");
    tr.append_point_var("DeltaSmall", pr + 32);
    tr.append_point_var("DeltaCapital", pr + 64);
    tr.append_point_var("d", pr);
    sc x;
    transcript_challenge(tr, "challenge", x);
    if (mul(bb, x) != bt[2]) return QQ_ST_PROOF;
    jobs.set(p, m0, 0, x, commitment_a);
    jobs.set(p, m0, 1, one(), pr);
    jobs.set(p, m0, 2, neg(rt), g.H);
    for (int i = 0; i < 3; i++) jobs.set(p, m0, 3 + i, neg(at[i]), g.G + 32 * i);
    jobs.set(p, m0 + 1, 0, x, pr + 64);
    jobs.set(p, m0 + 1, 1, one(), pr + 32);
    jobs.set(p, m0 + 1, 2, neg(st_), g.H);
    for (int i = 0; i < 2; i++) jobs.set(p, m0 + 1, 3 + i, neg(sub(mul(bt[i + 1], x), mul(bt[i], at[i + 1]))), g.G + 32 * i);
    return QQ_ST_OK;
}

// HadamardProof::verify (reference src/shuffle/hadamard.rs:249-389) on a running transcript; fills slots m0 .. m0 + 3 of
// proof p.  cm[k]: the three commitments of a / b / c.  With l(X) = prod (X - omega_j) and the Lagrange basis l_i(X):
//   l(x) c_a0 + sum_i l_i(x) c_a_i == com(a_bar; r_bar)  (and the same for b, c)                 [MSMs m0 .. m0 + 2]
//   l(x) sum_i x^i c_delta_i == com(a_bar o b_bar - c_bar; rho_bar)                              [MSM m0 + 3]
// The first three are only ever compared with the identity, so each is emitted multiplied by D = prod_i den_i, den_i =
// (w_i - w_j)(w_i - w_k): l_i(x) D = num_i(x) prod_{j != i} den_j needs no inversion in Z/l (D != 0: the omegas are distinct).
QQ_HOSTDEV static inline void hadamard_phase(qq_merlin::transcript& tr, const uint8_t* pr, const uint8_t* omega,
                                             const uint8_t* const cm[3], const gens& g, const job_sink& jobs, size_t p, size_t m0,
                                             uint8_t& pre, uint8_t& pre_detail) {
    using namespace qq_sc;
    pre = QQ_ST_OK;
    pre_detail = 0;
    sc w[3], bar[3][3], blind[3], rho_bar;
    bool canon = from_bytes(rho_bar, pr + 608);
    for (int i = 0; i < 3; i++) {
        canon = canon && from_bytes(w[i], omega + 32 * i) && from_bytes(blind[i], pr + 512 + 32 * i);
        for (int k = 0; k < 3; k++) canon = canon && from_bytes(bar[k][i], pr + 224 + 96 * k + 32 * i);
    }
    if (!canon) {
        pre = QQ_ST_BAD_SCALAR;
        return;
    }
    if (w[0] == w[1] || w[0] == w[2] || w[1] == w[2]) {
        pre = QQ_ST_PROOF;
        pre_detail = 1;
        return;
    }
    tr.domain_sep("HadamardProductProof");
    for (int i = 0; i < 3; i++) {
        tr.append_point_var("c_a", cm[0] + 32 * i);
        tr.append_point_var("c_b", cm[1] + 32 * i);
        tr.append_point_var("c_c", cm[2] + 32 * i);
    }
    tr.append_point_var("c_a_0", pr);
    tr.append_point_var("c_b_0", pr + 32);
    tr.append_point_var("c_c_0", pr + 64);
    for (int i = 0; i < 4; i++) tr.append_point_var("c_delta", pr + 96 + 32 * i);
    sc x;
    transcript_challenge(tr, "challenge", x);
    // l(x) and the Lagrange basis at x (polynomial::create_l_i_x_polynomial, src/shuffle/polynomial.rs:367-391)
    sc d[3] = {sub(x, w[0]), sub(x, w[1]), sub(x, w[2])};
    const sc lx = mul(mul(d[0], d[1]), d[2]);      // l(x)
    sc evD[4], D;                                  // l(x) D and l_i(x) D
    {
        sc den[3];
        for (int i = 0; i < 3; i++) den[i] = mul(sub(w[i], w[(i + 1) % 3]), sub(w[i], w[(i + 2) % 3]));
        sc p01 = mul(den[0], den[1]);
        D = mul(p01, den[2]);
        const sc cof[3] = {mul(den[1], den[2]), mul(den[0], den[2]), p01};
        evD[0] = mul(lx, D);
        for (int i = 0; i < 3; i++) evD[i + 1] = mul(mul(d[(i + 1) % 3], d[(i + 2) % 3]), cof[i]);
    }
    const sc nD = neg(D);
    for (int k = 0; k < 3; k++) {
        jobs.set(p, m0 + k, 0, evD[0], pr + 32 * k);
        for (int i = 0; i < 3; i++) jobs.set(p, m0 + k, 1 + i, evD[i + 1], cm[k] + 32 * i);
        jobs.set(p, m0 + k, 4, mul(nD, blind[k]), g.H);
        for (int i = 0; i < 3; i++) jobs.set(p, m0 + k, 5 + i, mul(nD, bar[k][i]), g.G + 32 * i);
    }
    sc xi = lx;
    for (int i = 0; i < 4; i++) {
        jobs.set(p, m0 + 3, i, xi, pr + 96 + 32 * i);
        xi = mul(xi, x);
    }
    jobs.set(p, m0 + 3, 4, neg(rho_bar), g.H);
    for (int i = 0; i < 3; i++) jobs.set(p, m0 + 3, 5 + i, neg(sub(mul(bar[0][i], bar[1][i]), bar[2][i])), g.G + 32 * i);
}
QQ_HOSTDEV static inline void hadamard_verdict(uint8_t pre, uint8_t pre_detail, const uint8_t* st4, const uint8_t* e4,
                                               uint8_t& status, uint8_t& detail) {
    status = pre;
    detail = pre_detail;
    if (pre != QQ_ST_OK) return;
    if (st4[0] || st4[1] || st4[2]) status = QQ_ST_BAD_POINT;
    else if (!is_zero32(e4) || !is_zero32(e4 + 32) || !is_zero32(e4 + 64)) { status = QQ_ST_PROOF; detail = 2; }
    else if (st4[3]) status = QQ_ST_BAD_POINT;
    else if (!is_zero32(e4 + 96)) { status = QQ_ST_PROOF; detail = 3; }
}

// ProductProof::verify (reference src/shuffle/product.rs:170-195) = MultiHadamardProof::verify (:325-389) -> ZeroProof::verify
// (:508-600) -> SVPProof::verify on one running transcript; fills slots m0 .. m0 + 6 of proof p:
//   0: c_B decodes                                   1: c_D[m+1] is the identity
//   2: c_A0 + sum x^i c_A_i == com(a; r)             3: sum x^(m-i) c_B'_i + c_Bm == com(b; s)   (c_B' folded into c_B)
//   4: sum x^k c_D_k == (a * b) B + t B_blinding     5, 6: the SVP equations
// pre: status before any group check, svp_pre: the SVP's scalar checks (after the zero argument in the reference's order).
// cA: the three column commitments the argument is about, or nullptr when the caller compares them itself.
QQ_HOSTDEV static inline void product_phase(qq_merlin::transcript& tr, const uint8_t* pr, const uint8_t* stm, const uint8_t* cA,
                                            const gens& g, const job_sink& jobs, size_t p, size_t m0, uint8_t& pre,
                                            uint8_t& pre_detail, uint8_t& svp_pre) {
    using namespace qq_sc;
    pre = QQ_ST_OK;
    pre_detail = 0;
    svp_pre = QQ_ST_OK;
    const uint8_t *cB = pr, *cA0 = pr + 96, *cBm = pr + 128, *cD = pr + 160, *zA = stm + 32;
    sc av[3], bv[3], r, s_, t;
    bool canon = from_bytes(r, pr + 576) && from_bytes(s_, pr + 608) && from_bytes(t, pr + 640);
    for (int i = 0; i < 3; i++) canon = canon && from_bytes(av[i], pr + 384 + 32 * i) && from_bytes(bv[i], pr + 480 + 32 * i);
    if (!canon) {
        pre = QQ_ST_BAD_SCALAR;
        return;
    }
    if (cA != nullptr && (differ32(cA, cB) || differ32(cA + 32, zA) || differ32(cA + 64, zA + 32))) {
        pre = QQ_ST_PROOF;
        pre_detail = 1;
        return;
    }
    if (differ32(stm, cB + 64)) {
        pre = QQ_ST_PROOF;
        pre_detail = 2;
        return;
    }
    tr.domain_sep("MultiHadamardProductProof");
    for (int i = 0; i < 3; i++) tr.append_point_var("BVectorCommitment", cB + 32 * i);
    sc xm, y;
    transcript_challenge(tr, "XChallenge", xm);
    transcript_challenge(tr, "YChallenge", y);
    tr.domain_sep("ZeroArgumentProof");
    tr.append_point_var("A0Commitment", cA0);
    tr.append_point_var("BmCommitment", cBm);
    for (int k = 0; k < 7; k++) tr.append_point_var("DCommitment", cD + 32 * k);
    sc x;
    transcript_challenge(tr, "challenge", x);
    sc xe[7];
    xe[0] = one();
    for (int k = 1; k < 7; k++) xe[k] = mul(xe[k - 1], x);
    sc xm2 = mul(xm, xm);
    for (int i = 0; i < 3; i++) jobs.set(p, m0, i, one(), cB + 32 * i);
    jobs.set(p, m0 + 1, 0, one(), cD + 32 * 4);
    // a: c_A0 + x zA_0 + x^2 zA_1 + x^3 com(-1, -1, -1; 0) - com(a; r)
    jobs.set(p, m0 + 2, 0, one(), cA0);
    jobs.set(p, m0 + 2, 1, xe[1], zA);
    jobs.set(p, m0 + 2, 2, xe[2], zA + 32);
    jobs.set(p, m0 + 2, 3, neg(r), g.H);
    for (int i = 0; i < 3; i++) jobs.set(p, m0 + 2, 4 + i, neg(add(av[i], xe[3])), g.G + 32 * i);
    // b: x^3 (xm c_B0) + x^2 (xm^2 c_B1) + x (xm c_B1 + xm^2 c_B2) + c_Bm - com(b; s)
    jobs.set(p, m0 + 3, 0, one(), cBm);
    jobs.set(p, m0 + 3, 1, mul(xe[3], xm), cB);
    jobs.set(p, m0 + 3, 2, add(mul(xe[2], xm2), mul(xe[1], xm)), cB + 32);
    jobs.set(p, m0 + 3, 3, mul(xe[1], xm2), cB + 64);
    jobs.set(p, m0 + 3, 4, neg(s_), g.H);
    for (int i = 0; i < 3; i++) jobs.set(p, m0 + 3, 5 + i, neg(bv[i]), g.G + 32 * i);
    // a * b = sum a_i b_i y^(i+1)
    sc abb = zero(), yi = y;
    for (int i = 0; i < 3; i++) {
        abb = add(abb, mul(mul(av[i], bv[i]), yi));
        yi = mul(yi, y);
    }
    for (int k = 0; k < 7; k++) jobs.set(p, m0 + 4, k, xe[k], cD + 32 * k);
    jobs.set(p, m0 + 4, 7, neg(abb), g.B);
    jobs.set(p, m0 + 4, 8, neg(t), g.Hp);
    svp_pre = svp_phase(tr, pr + 672, stm + 128, stm + 160, g, jobs, p, m0 + 5);
    if (svp_pre == QQ_ST_BAD_SCALAR) pre = QQ_ST_BAD_SCALAR;
}
// verdict of one product argument from its seven MSM results (status bytes st7, encodings e7)
QQ_HOSTDEV static inline void product_verdict(uint8_t pre, uint8_t pre_detail, uint8_t svp_pre, const uint8_t* st7,
                                              const uint8_t* e7, uint8_t& status, uint8_t& detail) {
    status = pre;
    detail = pre_detail;
    if (pre != QQ_ST_OK) return;
    if (st7[0]) { status = QQ_ST_BAD_POINT; detail = 10; return; }
    for (int m = 1; m <= 4; m++) {
        if (st7[m]) { status = QQ_ST_BAD_POINT; detail = 11; return; }
        if (!is_zero32(e7 + 32 * m)) { status = QQ_ST_PROOF; detail = (uint8_t)(2 + m); return; }
    }
    if (svp_pre != QQ_ST_OK) { status = QQ_ST_PROOF; detail = 7; return; }
    for (int m = 5; m <= 6; m++) {
        if (st7[m]) { status = QQ_ST_BAD_POINT; detail = 12; return; }
        if (!is_zero32(e7 + 32 * m)) { status = QQ_ST_PROOF; detail = 7; return; }
    }
}

// transcript of one multi-exponentiation argument -> x^0 .. x^5
QQ_HOSTDEV static inline void mexp_transcript(qq_merlin::transcript& tr, const uint8_t* mp, bool commitment_form, sc xe[6]) {
    tr.domain_sep(commitment_form ? "MultiExponentialElgamalCommmitmentProof" : "MultiExponentialPubKeyProof");
    tr.append_point_var("A0Commitment", mp);
    for (int k = 0; k < 6; k++) {
        tr.append_point_var("BKCommitment", mp + 32 + 32 * k);
        tr.append_point_var("EK0Commitment", mp + 224 + 32 * k);
        tr.append_point_var("EK1Commitment", mp + 416 + 32 * k);
    }
    sc x;
    transcript_challenge(tr, "xchallenege", x);
    xe[0] = qq_sc::one();
    for (int k = 1; k < 6; k++) xe[k] = qq_sc::mul(xe[k - 1], x);
}
// verify_multiexpo_scalars as two identity MSMs (slots m0, m0 + 1; 8 terms each)
QQ_HOSTDEV static inline void mexp_scalar_jobs(const uint8_t* mp, const uint8_t* cA, const sc xe[6], const sc av[3], const sc& r,
                                               const sc& b, const sc& s_, const gens& g, const job_sink& jobs, size_t p, size_t m0) {
    using namespace qq_sc;
    jobs.set(p, m0, 0, one(), mp);
    for (int i = 0; i < 3; i++) jobs.set(p, m0, 1 + i, xe[1 + i], cA + 32 * i);
    jobs.set(p, m0, 4, neg(r), g.H);
    for (int i = 0; i < 3; i++) jobs.set(p, m0, 5 + i, neg(av[i]), g.G + 32 * i);
    for (int k = 0; k < 6; k++) jobs.set(p, m0 + 1, k, xe[k], mp + 32 + 32 * k);
    jobs.set(p, m0 + 1, 6, neg(b), g.B);
    jobs.set(p, m0 + 1, 7, neg(s_), g.Hp);
}
// The E_K equations have 16 or 17 terms: sum_k x^k E_k[k] - sum_j a_j (x^2 P_{0,j} + x P_{1,j} + P_{2,j}) - .. == 0 with
// P_{r,j} = acc[3 r + j] + off.  k_straus handles 10 terms per pass over the 252 doublings, and one long instance holds up
// its whole wave, so each equation is evaluated as TWO MSMs of 8 and 8 (9) terms - slot m gets terms 0..7, slot m + 1 the
// NEGATED terms 8.. - and the check becomes enc(left) == enc(right) (canonical encodings).
QQ_HOSTDEV static inline void ek_set(const job_sink& jobs, size_t p, size_t m, int t, const sc& s, const uint8_t* point) {
    if (t < 8) jobs.set(p, m, t, s, point);
    else jobs.set(p, m + 1, t - 8, qq_sc::neg(s), point);
}
QQ_HOSTDEV static inline void mexp_ek_job(const uint8_t* ek, const uint8_t* accounts, size_t off, const sc xe[6], const sc av[3],
                                          const job_sink& jobs, size_t p, size_t m) {
    using namespace qq_sc;
    for (int k = 0; k < 6; k++) ek_set(jobs, p, m, k, xe[k], ek + 32 * k);
    for (int rr = 0; rr < 3; rr++)
        for (int j = 0; j < 3; j++)
            ek_set(jobs, p, m, 6 + 3 * rr + j, neg(mul(av[j], xe[2 - rr])), accounts + 128 * (3 * rr + j) + off);
}

// =================================================================================================================
// ShuffleProof::verify (reference src/shuffle/shuffle.rs:547-712): state of one proof between the phases
// =================================================================================================================
struct proof_state {
    qq_merlin::transcript tr;
    sc expx[9];                 // x, x^2, .. x^9
    uint8_t had_pre, had_det, prod_pre, prod_det, svp_pre, b_ok, pk_pre, cm_pre;
    uint8_t st, sg, dt;         // verdict so far: status, stage, detail (sg == 0: undecided)
    uint8_t clean;              // aggregate mode: 1 = every scalar check passed and the proof's equations are in the aggregate
    uint8_t pad[4];
    sc fixed[6];                // aggregate mode: this proof's scalars on the fixed generators (both batches)
    QQ_HOSTDEV explicit proof_state(const qq_merlin::transcript& t0)
        : tr(t0), had_pre(0), had_det(0), prod_pre(0), prod_det(0), svp_pre(0), b_ok(0), pk_pre(0), cm_pre(0), st(QQ_ST_OK), sg(0), dt(0), clean(0) {
        for (int i = 0; i < 9; i++) expx[i] = qq_sc::zero();
        for (int i = 0; i < 4; i++) pad[i] = 0;
        for (int i = 0; i < 6; i++) fixed[i] = qq_sc::zero();
    }
    QQ_HOSTDEV void fail(uint8_t s, uint8_t stage, uint8_t detail) {
        st = s;
        sg = stage;
        dt = detail;
    }
};

// pass A: transcript through the Hadamard and product arguments (every challenge there depends on proof bytes only); scalar
// check prod (y i + x^i - z) == b; the 18 MSMs of batch 1.  pr / stm / in: this proof's ShuffleProof, ShuffleStatement and
// input accounts (9 x 128 B).  S.tr must hold Transcript::new(label) + Verifier::new(label).
QQ_HOSTDEV static inline void pass_a(proof_state& S, const job_sink& j1, size_t p, const uint8_t* pr, const uint8_t* stm,
                                     const uint8_t* in, const gens& g) {
    using namespace qq_sc;
    const uint8_t *cA = pr, *ctau = pr + 96, *cB = pr + 192, *cBd = pr + 288;
    qq_merlin::transcript& tr = S.tr;
    for (int i = 0; i < 3; i++) {
        tr.append_point_var("ACommitment", cA + 32 * i);
        tr.append_point_var("tauCommitment", ctau + 32 * i);
    }
    sc x;
    transcript_challenge(tr, "xChallenge", x);
    sc* ex = S.expx;
    ex[0] = x;
    for (int i = 1; i < 9; i++) ex[i] = mul(ex[i - 1], x);
    for (int i = 0; i < 3; i++) {
        tr.append_point_var("BCommitment", cB + 32 * i);
        tr.append_point_var("BDashCommitment", cBd + 32 * i);
    }
    const uint8_t* cm[3] = {cBd, ctau, cB};
    hadamard_phase(tr, pr + 384, stm, cm, g, j1, p, 0, S.had_pre, S.had_det);
    if (S.had_pre != QQ_ST_OK) return;     // the verdict is the Hadamard argument's; nothing later is looked at
    sc y, z;
    transcript_challenge(tr, "yChallenge", y);
    transcript_challenge(tr, "zChallenge", z);
    sc bstm;
    if (!from_bytes(bstm, stm + 96 + 160)) {
        S.had_pre = QQ_ST_BAD_SCALAR;
        return;
    }
    sc product = one();
    for (int i = 0; i < 9; i++) product = mul(product, sub(add(mul(y, from_u64((uint64_t)i + 1)), ex[i]), z));
    S.b_ok = product == bstm ? 1 : 0;
    sc nz = neg(z);
    for (int i = 0; i < 3; i++) {
        j1.set(p, 4 + i, 0, y, cA + 32 * i);
        j1.set(p, 4 + i, 1, one(), cB + 32 * i);
        for (int k = 0; k < 3; k++) j1.set(p, 4 + i, 2 + k, nz, g.G + 32 * k);
        // aggregate mode: c_E_i == the product argument's c_B_1 / zero_statement.c_A (MultiHadamardProof::verify's first check)
        // as an equation: c_E_i - dec(those bytes) == identity
        if (j1.agg) j1.set(p, 4 + i, 5, neg(one()), i == 0 ? pr + 1024 : stm + 96 + 32 * i);
    }
    product_phase(tr, pr + 1024, stm + 96, nullptr, g, j1, p, 7, S.prod_pre, S.prod_det, S.svp_pre);
    // G, H = sum x^i pk_i;  g_r = z G + c G_dash, h_r = z H + c H_dash with G, H expanded over the keys
    sc dc, dz;
    if (!from_bytes(dc, pr + 3712) || !from_bytes(dz, pr + 3744)) {
        S.had_pre = QQ_ST_BAD_SCALAR;
        return;
    }
    if (j1.agg) {
        // Aggregate mode.  An accepted proof carries the encodings of G and H itself: the pubkey argument checks
        // E_k_0[3] == enc(G), E_k_1[3] == enc(H) ("Verify Em == C").  So G and H are TAKEN from those bytes - the transcript
        // absorbs them, g_r = z G + c G_dash and h_r become two-term MSMs on them - and "sum x^i pk_i == dec(E_k[3])" joins the
        // aggregated equations.  If the claim is false the aggregate fails and the exact form gives the verdict.
        const uint8_t *Gclaim = pr + 2048 + 224 + 96, *Hclaim = pr + 2048 + 416 + 96;
        for (int i = 0; i < 9; i++) {
            j1.set(p, 14, i, ex[i], in + 128 * i);
            j1.set(p, 15, i, ex[i], in + 128 * i + 32);
        }
        j1.set(p, 14, 9, neg(one()), Gclaim);
        j1.set(p, 15, 9, neg(one()), Hclaim);
        j1.set(p, 16, 0, dz, Gclaim);
        j1.set(p, 16, 1, dc, stm + 288);
        j1.set(p, 17, 0, dz, Hclaim);
        j1.set(p, 17, 1, dc, stm + 320);
        return;
    }
    for (int i = 0; i < 9; i++) {
        j1.set(p, 14, i, ex[i], in + 128 * i);
        j1.set(p, 15, i, ex[i], in + 128 * i + 32);
        sc zx = mul(dz, ex[i]);
        j1.set(p, 16, i, zx, in + 128 * i);
        j1.set(p, 17, i, zx, in + 128 * i + 32);
    }
    j1.set(p, 16, 9, dc, stm + 288);
    j1.set(p, 17, 9, dc, stm + 320);
}

// pass B: verdicts of batch 1 (e: 18 x 32 B encodings, s: 18 status bytes); DDH transcript (absorbs G, H, g_r, h_r) and
// challenge check; transcripts of the two multi-exponentiation arguments; the 14 MSMs of batch 2.
// eG / sG: encodings and status bytes of G, H, g_r, h_r (= e + 32 * 14, s + 14 in the exact form).
// Aggregate mode (j2.agg): e / s are not read (the other 14 MSMs of batch 1 are equations inside the aggregate); the proof
// is `clean` when every check that does not need a group result passed and both multi-exponentiation arguments were
// emitted - anything else leaves the verdict to the exact form.  Returns clean (always false in the exact form).
QQ_HOSTDEV static inline bool pass_b(proof_state& S, const job_sink& j2, size_t p, const uint8_t* pr, const uint8_t* stm,
                                     const uint8_t* in, const uint8_t* out, const uint8_t* e, const uint8_t* s, const uint8_t* eG,
                                     const uint8_t* sG, const gens& g) {
    using namespace qq_sc;
    const bool aggregate = j2.agg != nullptr;
    uint8_t vs, vd;
    if (aggregate) {
        if (S.had_pre != QQ_ST_OK || !S.b_ok || S.prod_pre != QQ_ST_OK || S.svp_pre != QQ_ST_OK) return false;
        if (sG[0] || sG[1] || sG[2] || sG[3]) return false;
    } else {
        hadamard_verdict(S.had_pre, S.had_det, s, e, vs, vd);
        if (vs != QQ_ST_OK) return S.fail(vs, 1, vd), false;
        if (!S.b_ok) return S.fail(QQ_ST_PROOF, 2, 0), false;
        if (s[4] || s[5] || s[6]) return S.fail(QQ_ST_BAD_POINT, 3, 0), false;
        {   // MultiHadamardProof::verify's first check against the computed c_E
            const uint8_t *cBp = pr + 1024, *zA = stm + 96 + 32;
            if (S.prod_pre != QQ_ST_BAD_SCALAR && (differ32(e + 32 * 4, cBp) || differ32(e + 32 * 5, zA) || differ32(e + 32 * 6, zA + 32)))
                return S.fail(QQ_ST_PROOF, 4, 1), false;
        }
        product_verdict(S.prod_pre, S.prod_det, S.svp_pre, s + 7, e + 32 * 7, vs, vd);
        if (vs != QQ_ST_OK) return S.fail(vs, 4, vd), false;
        if (sG[0] || sG[1]) return S.fail(QQ_ST_BAD_POINT, 5, 0), false;
        if (sG[2] || sG[3]) return S.fail(sG[2] == QQ_ST_BAD_SCALAR || sG[3] == QQ_ST_BAD_SCALAR ? QQ_ST_BAD_SCALAR : QQ_ST_BAD_POINT, 6, 0), false;
    }
    const uint8_t *Gc = eG, *Hc = eG + 32;
    qq_merlin::transcript& tr = S.tr;
    tr.domain_sep("DDHTupleProof");
    tr.append_point_var("g", Gc);
    tr.append_point_var("g_dash", stm + 288);
    tr.append_point_var("h", Hc);
    tr.append_point_var("h_dash", stm + 320);
    tr.append_point_var("gr", eG + 64);
    tr.append_point_var("hr", eG + 96);
    uint8_t chal[32];
    tr.get_challenge("Challenge", chal);
    if (differ32(chal, pr + 3712)) {
        if (!aggregate) S.fail(QQ_ST_PROOF, 6, 0);
        return false;
    }
    // ---- pubkey argument (c_A = c_B_dash, base_pk = (B, H_pedersen), pk_GH = (G, H)), then the commitment argument
    const sc* ex = S.expx;
    for (int which = 0; which < 2; which++) {
        const uint8_t* mp = pr + (which ? 2880 : 2048);
        uint8_t& pre = which ? S.cm_pre : S.pk_pre;
        sc av[3], r, b, s_, t;
        bool canon = from_bytes(r, mp + 704) && from_bytes(b, mp + 736) && from_bytes(s_, mp + 768) && from_bytes(t, mp + 800);
        for (int i = 0; i < 3; i++) canon = canon && from_bytes(av[i], mp + 608 + 32 * i);
        if (!canon) {
            pre = 100;      // non-canonical scalar
            return false;
        }
        if (!is_zero32(mp + 32 + 32 * 3)) {
            pre = 1;
            return false;
        }
        if (which == 0 && (differ32(Gc, mp + 224 + 96) || differ32(Hc, mp + 416 + 96))) {
            pre = 2;
            return false;
        }
        sc xe[6];
        size_t m0 = 0;
        if (which == 1) {
            for (int i = 0; i < 9; i++) {
                j2.set(p, 6, i, ex[i], in + 128 * i + 64);
                j2.set(p, 7, i, ex[i], in + 128 * i + 96);
            }
            if (aggregate) {      // C_c == E_k_0[3], C_d == E_k_1[3] ("Verify Em == C") as equations
                j2.set(p, 6, 9, neg(one()), mp + 224 + 96);
                j2.set(p, 7, 9, neg(one()), mp + 416 + 96);
            }
            m0 = 8;
        }
        mexp_transcript(tr, mp, which == 1, xe);
        mexp_scalar_jobs(mp, which ? pr + 192 : pr + 288, xe, av, r, b, s_, g, j2, p, m0);
        if (which == 0) {
            mexp_ek_job(mp + 224, out, 0, xe, av, j2, p, 2);
            ek_set(j2, p, 2, 15, neg(b), g.B);
            mexp_ek_job(mp + 416, out, 32, xe, av, j2, p, 4);
            ek_set(j2, p, 4, 15, neg(b), g.Hp);
        } else {
            mexp_ek_job(mp + 224, out, 64, xe, av, j2, p, 10);
            ek_set(j2, p, 10, 15, neg(t), Gc);
            mexp_ek_job(mp + 416, out, 96, xe, av, j2, p, 12);
            ek_set(j2, p, 12, 15, neg(b), g.B);
            ek_set(j2, p, 12, 16, neg(t), Hc);
        }
    }
    return aggregate;
}

// Aggregate mode: sinks and weights of the two passes.  Batch 1: slots 16, 17 (g_r, h_r) are exact MSMs 0, 1 of the exact
// list {2, 2}; slots 14, 15 (G, H against the proof's own encodings of them) are aggregated equations like the rest; slot 7
// (c_B decodes) is skipped (the same points carry scalars in slot 10); 15 weights.
// Batch 2: no exact slots; the right half of an E_K pair carries minus the weight of the left half (ek_set negated its
// terms for the enc(left) == enc(right) comparison of the exact form); 10 weights.
QQ_HOSTDEV static inline void agg_begin_a(agg_ctx& a, const uint8_t entropy[32], size_t p) {
    sc w[15];
    agg_weights(w, 15, entropy, (uint64_t)p, 1);
    int k = 0;
    for (int m = 0; m < QQ_JOB_MAX_MSMS; m++) a.w[m] = qq_sc::zero();
    for (int m = 0; m < 16; m++)
        if (m != 7) a.w[m] = w[k++];
    for (int i = 0; i < 6; i++) a.fixed[i] = qq_sc::zero();
    a.k = 0;
    a.skip_mask = 1u << 7;
    a.overflow = false;
}
QQ_HOSTDEV static inline void agg_begin_b(agg_ctx& a, const uint8_t entropy[32], size_t p) {
    sc w[10];
    agg_weights(w, 10, entropy, (uint64_t)p, 2);
    for (int m = 0; m < QQ_JOB_MAX_MSMS; m++) a.w[m] = qq_sc::zero();
    const int single[6] = {0, 1, 6, 7, 8, 9}, pair[4] = {2, 4, 10, 12};
    int k = 0;
    for (int i = 0; i < 6; i++) a.w[single[i]] = w[k++];
    for (int i = 0; i < 4; i++) {
        a.w[pair[i]] = w[k++];
        a.w[pair[i] + 1] = qq_sc::neg(a.w[pair[i]]);
    }
    for (int i = 0; i < 6; i++) a.fixed[i] = qq_sc::zero();
    a.k = 0;
    a.skip_mask = 0;
    a.overflow = false;
}

// final verdict from batch 2 (e: 14 x 32 B encodings, s: 14 status bytes); no-op when an earlier phase decided
QQ_HOSTDEV static inline void pass_final(proof_state& S, const uint8_t* pr, const uint8_t* e, const uint8_t* s) {
    if (!(S.st == QQ_ST_OK && S.sg == 0)) return;
    // true = this check ends the verification
#define QQ_SHF_CHECK(m, stg, d)                                                                  \
    if (s[m]) { S.fail(s[m] == QQ_ST_BAD_SCALAR ? QQ_ST_BAD_SCALAR : QQ_ST_BAD_POINT, stg, d); return; } \
    if (!is_zero32(e + 32 * (m))) { S.fail(QQ_ST_PROOF, stg, d); return; }
    // an E_K equation: left half == negated right half
#define QQ_SHF_CHECK_PAIR(m, stg, d)                                                             \
    for (int h = 0; h < 2; h++)                                                                  \
        if (s[(m) + h]) { S.fail(s[(m) + h] == QQ_ST_BAD_SCALAR ? QQ_ST_BAD_SCALAR : QQ_ST_BAD_POINT, stg, d); return; } \
    if (differ32(e + 32 * (m), e + 32 * ((m) + 1))) { S.fail(QQ_ST_PROOF, stg, d); return; }
    if (S.pk_pre == 100) return S.fail(QQ_ST_BAD_SCALAR, 7, 0);
    if (S.pk_pre) return S.fail(QQ_ST_PROOF, 7, S.pk_pre);
    QQ_SHF_CHECK(0, 7, 3)
    QQ_SHF_CHECK(1, 7, 4)
    QQ_SHF_CHECK_PAIR(2, 7, 5)
    QQ_SHF_CHECK_PAIR(4, 7, 5)
    if (S.cm_pre == 100) return S.fail(QQ_ST_BAD_SCALAR, 8, 0);
    if (S.cm_pre) return S.fail(QQ_ST_PROOF, 8, S.cm_pre);
    if (s[6] || s[7]) return S.fail(QQ_ST_BAD_POINT, 8, 0);
    const uint8_t* mp = pr + 2880;
    if (differ32(e + 32 * 6, mp + 224 + 96) || differ32(e + 32 * 7, mp + 416 + 96)) return S.fail(QQ_ST_PROOF, 8, 2);
    QQ_SHF_CHECK(8, 8, 3)
    QQ_SHF_CHECK(9, 8, 4)
    QQ_SHF_CHECK_PAIR(10, 8, 5)
    QQ_SHF_CHECK_PAIR(12, 8, 5)
#undef QQ_SHF_CHECK
#undef QQ_SHF_CHECK_PAIR
}

}  // namespace qq_shuffle

// =================================================================================================================
// transcript kernels: one thread per proof, one warp per block (4 096 proofs = 128 warps, spread over the SMs)
// =================================================================================================================
#ifdef __CUDACC__
namespace qq_shuffle {

struct dev_inputs {
    const uint8_t *in, *out, *stm, *proof;     // nproofs x (9 x 128 | 9 x 128 | 352 | 3776) B, device
    size_t nproofs;
};
// job slots of one batch pre-filled with 0 * B (a proof that stops at a scalar pre-check leaves them untouched) + CSR offsets
__global__ void k_shuffle_jobs_prefill(job_sink j, size_t nproofs, const uint8_t* __restrict__ B, uint32_t* __restrict__ offs) {
    size_t total = nproofs * j.terms_pp;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    uint4 b0 = reinterpret_cast<const uint4*>(B)[0], b1 = reinterpret_cast<const uint4*>(B)[1];
    uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        reinterpret_cast<uint4*>(j.sc)[2 * t] = z;
        reinterpret_cast<uint4*>(j.sc)[2 * t + 1] = z;
        reinterpret_cast<uint4*>(j.pt)[2 * t] = b0;
        reinterpret_cast<uint4*>(j.pt)[2 * t + 1] = b1;
        size_t p = t / j.terms_pp, r = t % j.terms_pp;
        if (r < j.msms_pp) offs[p * j.msms_pp + r] = (uint32_t)(p * j.terms_pp + j.first[r]);
        if (t == 0) offs[nproofs * j.msms_pp] = (uint32_t)total;
    }
}
__global__ void __launch_bounds__(32) k_shuffle_pass_a(dev_inputs d, gens g, job_sink j1, qq_merlin::transcript tr0,
                                                       proof_state* __restrict__ states) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.nproofs) return;
    proof_state S(tr0);
    pass_a(S, j1, p, d.proof + QQ_SHUFFLE_PROOF_BYTES * p, d.stm + QQ_SHUFFLE_STATEMENT_BYTES * p, d.in + 9 * 128 * p, g);
    states[p] = S;
}
__global__ void __launch_bounds__(32) k_shuffle_pass_b(dev_inputs d, gens g, job_sink j2, proof_state* __restrict__ states,
                                                       const uint8_t* __restrict__ e1, const uint8_t* __restrict__ s1) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.nproofs) return;
    proof_state S = states[p];
    const uint8_t *e = e1 + 32 * QQ_SHUFFLE_MSMS_1 * p, *s = s1 + QQ_SHUFFLE_MSMS_1 * p;
    pass_b(S, j2, p, d.proof + QQ_SHUFFLE_PROOF_BYTES * p, d.stm + QQ_SHUFFLE_STATEMENT_BYTES * p, d.in + 9 * 128 * p,
           d.out + 9 * 128 * p, e, s, e + 32 * 14, s + 14, g);
    states[p] = S;
}
// verdict bytes: out3 = status[nproofs] | stage[nproofs] | detail[nproofs]
__global__ void __launch_bounds__(32) k_shuffle_final(dev_inputs d, proof_state* __restrict__ states, const uint8_t* __restrict__ e2,
                                                      const uint8_t* __restrict__ s2, uint8_t* __restrict__ out3) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.nproofs) return;
    proof_state& S = states[p];
    pass_final(S, d.proof + QQ_SHUFFLE_PROOF_BYTES * p, e2 + 32 * QQ_SHUFFLE_MSMS_2 * p, s2 + QQ_SHUFFLE_MSMS_2 * p);
    out3[p] = S.st;
    out3[d.nproofs + p] = S.sg;
    out3[2 * d.nproofs + p] = S.dt;
}


// ---- aggregate mode --------------------------------------------------------------------------------------------------
struct entropy32 {
    uint8_t b[32];
};
// n MSM terms = 0 * B
__global__ void k_shuffle_terms_clear(uint8_t* __restrict__ sc, uint8_t* __restrict__ pt, size_t n, const uint8_t* __restrict__ B) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    uint4 b0 = reinterpret_cast<const uint4*>(B)[0], b1 = reinterpret_cast<const uint4*>(B)[1];
    uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        reinterpret_cast<uint4*>(sc)[2 * t] = z;
        reinterpret_cast<uint4*>(sc)[2 * t + 1] = z;
        reinterpret_cast<uint4*>(pt)[2 * t] = b0;
        reinterpret_cast<uint4*>(pt)[2 * t + 1] = b1;
    }
}
__global__ void __launch_bounds__(32) k_shuffle_pass_a_agg(dev_inputs d, gens g, job_sink j1, qq_merlin::transcript tr0, entropy32 ent,
                                                           proof_state* __restrict__ states) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.nproofs) return;
    agg_ctx A;
    agg_begin_a(A, ent.b, p);
    j1.agg = &A;
    proof_state S(tr0);
    pass_a(S, j1, p, d.proof + QQ_SHUFFLE_PROOF_BYTES * p, d.stm + QQ_SHUFFLE_STATEMENT_BYTES * p, d.in + 9 * 128 * p, g);
    for (int i = 0; i < 6; i++) S.fixed[i] = A.fixed[i];
    S.clean = A.overflow ? 0 : 1;
    states[p] = S;
}
// ex / sx: the exact MSMs g_r, h_r of every proof (2 x 32 B, 2 status bytes per proof); G and H are the proof's own encodings
// of them (see pass_a).  A proof that is not clean drops out of the aggregate (its scalars are zeroed) and is verified in the
// exact form afterwards.
__global__ void __launch_bounds__(32) k_shuffle_pass_b_agg(dev_inputs d, gens g, job_sink j1, job_sink j2, entropy32 ent,
                                                           proof_state* __restrict__ states, const uint8_t* __restrict__ ex,
                                                           const uint8_t* __restrict__ sx, uint8_t* __restrict__ clean_out) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= d.nproofs) return;
    proof_state S = states[p];
    agg_ctx A;
    agg_begin_b(A, ent.b, p);
    j2.agg = &A;
    bool clean = S.clean != 0;
    if (clean) {
        const uint8_t* pr = d.proof + QQ_SHUFFLE_PROOF_BYTES * p;
        alignas(16) uint8_t eG[128];
        uint8_t sG[4] = {0, 0, sx[2 * p], sx[2 * p + 1]};
        memcpy(eG, pr + 2048 + 224 + 96, 32);
        memcpy(eG + 32, pr + 2048 + 416 + 96, 32);
        memcpy(eG + 64, ex + 64 * p, 64);
        clean = pass_b(S, j2, p, pr, d.stm + QQ_SHUFFLE_STATEMENT_BYTES * p, d.in + 9 * 128 * p, d.out + 9 * 128 * p, nullptr, nullptr,
                       eG, sG, g) && !A.overflow;
    }
    uint4 z = make_uint4(0, 0, 0, 0);
    if (clean) {
        for (int i = 0; i < 6; i++) states[p].fixed[i] = qq_sc::add(S.fixed[i], A.fixed[i]);
    } else {
        // out of the aggregate: zero scalars on a decodable point (an undecodable point of THIS proof must not send the
        // whole slice to the exact form)
        for (int i = 0; i < 6; i++) states[p].fixed[i] = qq_sc::zero();
        const uint4 b0 = reinterpret_cast<const uint4*>(g.B)[0], b1 = reinterpret_cast<const uint4*>(g.B)[1];
        uint4* a1 = reinterpret_cast<uint4*>(j1.asc + (size_t)32 * j1.cap * (p - j1.base));
        uint4* p1 = reinterpret_cast<uint4*>(j1.apt + (size_t)32 * j1.cap * (p - j1.base));
        for (uint32_t t = 0; t < 2 * j1.cap; t++) {
            a1[t] = z;
            p1[t] = (t & 1) ? b1 : b0;
        }
        uint4* a2 = reinterpret_cast<uint4*>(j2.asc + (size_t)32 * j2.cap * (p - j2.base));
        uint4* p2 = reinterpret_cast<uint4*>(j2.apt + (size_t)32 * j2.cap * (p - j2.base));
        for (uint32_t t = 0; t < 2 * j2.cap; t++) {
            a2[t] = z;
            p2[t] = (t & 1) ? b1 : b0;
        }
    }
    states[p].clean = clean ? 1 : 0;
    clean_out[p] = clean ? 1 : 0;
}
// the six fixed-generator terms of the aggregate: scalar i = sum over the proofs of states[p].fixed[i].  grid = 6 blocks for
// the whole batch (group_size >= nproofs), 6 G blocks for the grouped form: block 6 grp + i sums the proofs
// [grp * group_size, (grp + 1) * group_size) into term 6 grp + i.
__global__ void __launch_bounds__(256) k_shuffle_fixed_sum(const proof_state* __restrict__ states, size_t nproofs, size_t group_size, gens g,
                                                           uint8_t* __restrict__ out_sc, uint8_t* __restrict__ out_pt) {
    __shared__ qq_sc::sc part[256];
    const int i = blockIdx.x % 6;
    const size_t grp = blockIdx.x / 6, lo = grp * group_size, hi = lo + group_size < nproofs ? lo + group_size : nproofs;
    qq_sc::sc acc = qq_sc::zero();
    for (size_t p = lo + threadIdx.x; p < hi; p += blockDim.x) acc = qq_sc::add(acc, states[p].fixed[i]);
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int h = 128; h > 0; h >>= 1) {
        if ((int)threadIdx.x < h) part[threadIdx.x] = qq_sc::add(part[threadIdx.x], part[threadIdx.x + h]);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint8_t* pt = i == 0 ? g.B : i == 1 ? g.Hp : i == 2 ? g.H : g.G + 32 * (i - 3);
        job_sink::put(out_sc, out_pt, 6 * grp + (size_t)i, part[0], pt);
    }
}
// group of every term of the aggregated list [nproofs x cap1 | nproofs x cap2 | 6 per group] (grouped form, qq_msm_grouped's layout)
__global__ void k_shuffle_group_of(size_t nproofs, size_t cap1, size_t cap2, size_t group_size, size_t groups, unsigned int* __restrict__ group_of) {
    const size_t n1 = nproofs * cap1, n2 = nproofs * cap2, total = n1 + n2 + 6 * groups;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride)
        group_of[t] = (unsigned int)(t < n1 ? (t / cap1) / group_size : t < n1 + n2 ? ((t - n1) / cap2) / group_size : (t - n1 - n2) / 6);
}

}  // namespace qq_shuffle
#endif
