// Per-transcript phase of the batched Bulletproofs range-proof verifier, written once for host and device
// (RangeProof::verify_multiple / verify_single of the `bulletproofs` crate as the reference calls them, src/accounts/verifier.rs:
// 504-555; crate not vendored, restated from its published algorithm): the Merlin transcript of one proof chain, the challenges
// y, z, x, w, u_1..u_k, one inversion in Z/l for all of them, the record the fold kernel reads (rangeproof.cuh) and the
// proof-specific terms of the aggregated MSM.  qq_api_rangeproof.inc runs it one GPU thread per transcript (k_rp_transcripts:
// the proof bytes are the only upload) or, as a measurement knob, on the host threads.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "merlin_host.hpp"
#include "sc_host.hpp"
#include "shuffle_verify.cuh"      // is_zero32, QQ_ST_*

namespace qq_rp {

using qq_sc::sc;

#define QQ_RP_MAX_LG 10      // n m <= 64 x 16 = 1024
#define QQ_RP_MAX_PARTIES 16

// one record per (sub-)proof (all scalars canonical, already multiplied by the batch weight rho)
struct record {
    sc neg_rz;                      // -rho z
    sc rz;                          //  rho z
    sc ra;                          //  rho a
    sc rb;                          //  rho b
    sc allinv;                      // (u_1 .. u_k)^-1
    sc usq[QQ_RP_MAX_LG];           // u^2 in creation order (the crate's challenges_sq)
    sc yinv_pow[QQ_RP_MAX_LG];      // y^-(2^j)
    sc rzz_zj[QQ_RP_MAX_PARTIES];   // rho z^2 z^j
};

struct shape {
    uint32_t n_bits, m, chain, N, lg, T, proof_bytes;
    bool has_domain;
    char domain_label[48];
};

QQ_HOSTDEV static inline void append_u64(qq_merlin::transcript& tr, const char* label, uint64_t v) {
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(v >> (8 * i));
    tr.append_message(label, b, 8);
}
QQ_HOSTDEV static inline sc challenge(qq_merlin::transcript& tr, const char* label) {
    uint8_t wide[64];
    tr.challenge_bytes(label, wide, 64);
    return qq_sc::from_wide(wide);
}

// One transcript (a chain of sh.chain proofs over sh.m commitments each).  tr: Transcript::new + Verifier::new (or the
// imported state), before the domain separator.  proofs / commitments: this transcript's chain.  rec / s_out / p_out: the
// chain's records and its sh.chain x sh.T MSM terms (scalars, compressed points).  sB / sBt: the transcript's share of the
// scalars on B and B_blinding.  Returns the status before any group check: QQ_ST_OK, QQ_ST_BAD_SCALAR
// (RangeProof::from_bytes FormatError) or QQ_ST_PROOF (an identity point where the crate's validate_and_append_point refuses
// it, a zero challenge); everything of a rejected transcript is left out of the aggregate (zero scalars on a decodable point).
QQ_HOSTDEV static inline uint8_t transcript_phase(qq_merlin::transcript& tr, const shape& sh, const uint8_t* proofs,
                                                  const uint8_t* commitments, const uint8_t* entropy, const uint8_t* B,
                                                  record* rec, uint8_t* s_out, uint8_t* p_out, sc& sB, sc& sBt) {
    using namespace qq_sc;
    using qq_shuffle::is_zero32;
    const int lg = (int)sh.lg;
    uint8_t pre = QQ_ST_OK;
    sB = zero();
    sBt = zero();
    if (sh.has_domain) tr.domain_sep(sh.domain_label);
    for (uint32_t q = 0; q < sh.chain && pre == QQ_ST_OK; q++) {
        const uint8_t* pr = proofs + (size_t)sh.proof_bytes * q;
        const uint8_t* V = commitments + (size_t)32 * sh.m * q;
        const uint8_t *A = pr, *S = pr + 32, *T1 = pr + 64, *T2 = pr + 96, *LR = pr + 224;
        sc t_x, t_xb, e_b, a, b;
        if (!(from_bytes(t_x, pr + 128) && from_bytes(t_xb, pr + 160) && from_bytes(e_b, pr + 192) && from_bytes(a, LR + 64 * lg) &&
              from_bytes(b, LR + 64 * lg + 32))) {
            pre = QQ_ST_BAD_SCALAR;
            break;
        }
        tr.append_message("dom-sep", (const uint8_t*)"rangeproof v1", 13);
        append_u64(tr, "n", sh.n_bits);
        append_u64(tr, "m", sh.m);
        for (uint32_t j = 0; j < sh.m; j++) tr.append_message("V", V + 32 * j, 32);
        if (is_zero32(A) || is_zero32(S)) { pre = QQ_ST_PROOF; break; }      // validate_and_append_point
        tr.append_message("A", A, 32);
        tr.append_message("S", S, 32);
        const sc y = challenge(tr, "y"), z = challenge(tr, "z");
        if (is_zero32(T1) || is_zero32(T2)) { pre = QQ_ST_PROOF; break; }
        tr.append_message("T_1", T1, 32);
        tr.append_message("T_2", T2, 32);
        const sc x = challenge(tr, "x");
        tr.append_message("t_x", pr + 128, 32);
        tr.append_message("t_x_blinding", pr + 160, 32);
        tr.append_message("e_blinding", pr + 192, 32);
        const sc w = challenge(tr, "w");
        tr.append_message("dom-sep", (const uint8_t*)"ipp v1", 6);
        append_u64(tr, "n", sh.N);
        sc u[QQ_RP_MAX_LG + 1], usq[QQ_RP_MAX_LG];      // u_1 .. u_k, then y; replaced by their inverses below
        bool bad = false;
        for (int k = 0; k < lg && !bad; k++) {
            if (is_zero32(LR + 64 * k) || is_zero32(LR + 64 * k + 32)) { bad = true; break; }
            tr.append_message("L", LR + 64 * k, 32);
            tr.append_message("R", LR + 64 * k + 32, 32);
            u[k] = challenge(tr, "u");
            usq[k] = mul(u[k], u[k]);
            bad = is_zero(u[k]);                          // probability 2^-252
        }
        u[lg] = y;
        if (bad || is_zero(y)) { pre = QQ_ST_PROOF; break; }
        // weights: rho for the batch, c = the crate's random scalar, from a fork of the transcript keyed with fresh entropy
        sc rho, c;
        {
            qq_merlin::transcript fork = tr;
            fork.append_message("rng", entropy, 32);
            uint8_t wb[128];
            fork.challenge_bytes("batch-weights", wb, 128);
            rho = from_wide(wb);
            c = from_wide(wb + 64);
        }
        {   // every u_k and y inverted with ONE inversion (Montgomery's trick)
            sc prefix[QQ_RP_MAX_LG + 2];
            prefix[0] = one();
            for (int k = 0; k <= lg; k++) prefix[k + 1] = mul(prefix[k], u[k]);
#ifdef __CUDA_ARCH__
            sc inv_all = invert_fixed(prefix[lg + 1]);        // no divergent branches, operands in registers (sc_host.hpp)
#else
            sc inv_all = invert_vartime(prefix[lg + 1]);      // public challenges: the variable-time inversion is fine
#endif
            for (int k = lg; k >= 0; k--) {
                sc v = u[k];
                u[k] = mul(inv_all, prefix[k]);
                inv_all = mul(inv_all, v);
            }
        }
        const sc* uinv = u;
        record& r = rec[q];
        memset((void*)&r, 0, sizeof(r));
        sc allinv = one();
        for (int k = 0; k < lg; k++) allinv = mul(allinv, uinv[k]);
        r.allinv = allinv;
        sc yp = uinv[lg];
        for (int j = 0; j < lg; j++) {
            r.yinv_pow[j] = yp;
            yp = mul(yp, yp);
        }
        const sc zz = mul(z, z);
        r.rz = mul(rho, z);
        r.neg_rz = neg(r.rz);
        r.ra = mul(rho, a);
        r.rb = mul(rho, b);
        uint8_t* so = s_out + (size_t)32 * sh.T * q;
        const sc rx = mul(rho, x), rcx = mul(rx, c);
        to_bytes(so, rho);
        to_bytes(so + 32, rx);
        to_bytes(so + 64, rcx);
        to_bytes(so + 96, mul(rcx, x));
        for (int k = 0; k < lg; k++) {
            r.usq[k] = usq[k];
            to_bytes(so + 32 * (4 + k), mul(rho, usq[k]));
            to_bytes(so + 32 * (4 + lg + k), mul(rho, mul(uinv[k], uinv[k])));
        }
        sc rzz = mul(rho, zz), zj = one(), sum_z = zero();
        for (uint32_t j = 0; j < sh.m; j++) {
            r.rzz_zj[j] = mul(rzz, zj);
            to_bytes(so + 32 * (4 + 2 * lg + j), mul(c, r.rzz_zj[j]));
            sum_z = add(sum_z, zj);
            zj = mul(zj, z);
        }
        // delta(n, m, y, z) = (z - z^2) sum_{i < nm} y^i - z^3 (2^n - 1) sum_{j < m} z^j
        sc sum_y = one(), ypow = y;
        for (int j = 0; j < lg; j++) {
            sum_y = mul(sum_y, add(one(), ypow));
            ypow = mul(ypow, ypow);
        }
        const sc sum_2 = from_u64(sh.n_bits == 64 ? ~0ull : ((1ull << sh.n_bits) - 1));
        const sc delta = sub(mul(sub(z, zz), sum_y), mul(mul(mul(zz, z), sum_2), sum_z));
        const sc base_s = add(mul(w, sub(t_x, mul(a, b))), mul(c, sub(delta, t_x)));
        sB = add(sB, mul(rho, base_s));
        sBt = add(sBt, mul(rho, neg(add(e_b, mul(c, t_xb)))));
        // the proof's own points
        uint8_t* po = p_out + (size_t)32 * sh.T * q;
        qq_sc::copy_aligned16(po, pr, 128);                                            // A, S, T_1, T_2
        for (int k = 0; k < lg; k++) {
            qq_sc::copy_aligned16(po + 32 * (4 + k), pr + 224 + 64 * k, 32);           // L_k
            qq_sc::copy_aligned16(po + 32 * (4 + lg + k), pr + 224 + 64 * k + 32, 32);  // R_k
        }
        qq_sc::copy_aligned16(po + 32 * (4 + 2 * lg), V, (size_t)32 * sh.m);
    }
    if (pre != QQ_ST_OK) {      // left out of the aggregate: zero scalars on a decodable point
        memset((void*)rec, 0, sh.chain * sizeof(record));
        memset(s_out, 0, (size_t)sh.chain * sh.T * 32);
        for (size_t t = 0; t < (size_t)sh.chain * sh.T; t++) memcpy(p_out + 32 * t, B, 32);
        sB = zero();
        sBt = zero();
    }
    return pre;
}

}  // namespace qq_rp

#ifdef __CUDACC__
namespace qq_rp {
struct entropy32 {
    uint8_t b[32];
};
// one thread per transcript.  states: nullptr (every transcript starts from tr0 = Transcript::new + Verifier::new) or the
// serialised states of qq_transcript_capture; sB_out / sBt_out / pre_out: per transcript.
__global__ void __launch_bounds__(32) k_rp_transcripts(shape sh, qq_merlin::transcript tr0, const uint8_t* __restrict__ states,
                                                       const uint8_t* __restrict__ proofs, const uint8_t* __restrict__ commitments,
                                                       entropy32 ent, const uint8_t* __restrict__ B, size_t nproofs,
                                                       record* __restrict__ rec, uint8_t* __restrict__ s_out,
                                                       uint8_t* __restrict__ p_out, sc* __restrict__ sB_out,
                                                       sc* __restrict__ sBt_out, uint8_t* __restrict__ pre_out) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nproofs) return;
    qq_merlin::transcript tr = tr0;
    sc sB = qq_sc::zero(), sBt = qq_sc::zero();
    uint8_t pre;
    record* r = rec + p * sh.chain;
    uint8_t *so = s_out + p * (size_t)sh.chain * sh.T * 32, *po = p_out + p * (size_t)sh.chain * sh.T * 32;
    if (states != nullptr && !tr.import_state(states + qq_merlin::transcript::STATE_BYTES * p)) {
        // a state the sigma verification never wrote (its proof failed before the capture) or foreign bytes: rejected
        pre = QQ_ST_PROOF;
        memset((void*)r, 0, sh.chain * sizeof(record));
        memset(so, 0, (size_t)sh.chain * sh.T * 32);
        for (size_t t = 0; t < (size_t)sh.chain * sh.T; t++) memcpy(po + 32 * t, B, 32);
    } else {
        pre = transcript_phase(tr, sh, proofs + p * (size_t)sh.chain * sh.proof_bytes, commitments + p * (size_t)sh.chain * sh.m * 32,
                               ent.b, B, r, so, po, sB, sBt);
    }
    sB_out[p] = sB;
    sBt_out[p] = sBt;
    pre_out[p] = pre;
}
}  // namespace qq_rp
#endif
