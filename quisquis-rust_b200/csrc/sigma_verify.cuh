// Per-proof phases of the batched sigma-protocol verifiers of the reference's src/accounts/verifier.rs (and the DDH tuple proof
// of src/shuffle/ddh.rs), written once for host and device.  Every verifier recomputes a few commitments - 2- and 3-term MSMs
// over points of the accounts - absorbs them into the proof's Merlin transcript and compares the challenge it draws with the
// one in the proof.  Two phases per proof:
//   emit   : the MSM jobs of the proof (scalar, compressed point) into its slots of a segmented batch
//   finish : status of the MSMs -> the reference's Err / panic mapping, then the transcript script and the challenge comparison
// qq_api_sigma.inc runs them one GPU thread per proof (k_sigma_emit / k_sigma_finish: accounts, responses and challenges are the
// only upload, one status byte per proof comes back) or, with qq_verify_set_transcripts(ctx, 0), on the host threads.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "merlin_host.hpp"
#include "sc_host.hpp"
#include "shuffle_verify.cuh"      // QQ_ST_*, differ32

#ifndef QQ_ST_PANIC
#define QQ_ST_PANIC 7
#endif

namespace qq_sigma {

enum kind_t {
    DLOG = 0,           // Verifier::verify_update_account_verifier            verifier.rs:223-292
    DELTA_COMPACT = 1,  // Verifier::verify_delta_compact_verifier             verifier.rs:138-209
    ACCOUNT = 2,        // Verifier::verify_account_verifier_bulletproof       verifier.rs:396-470
    ZERO_BALANCE = 3,   // Verifier::zero_balance_account[_vector]_verifier    verifier.rs:593-680
    DESTROY = 4,        // Verifier::destroy_account_verifier                  verifier.rs:693-735
    SAME_VALUE = 5,     // Verifier::verify_same_value_compact_verifier        verifier.rs:747-806
    DARK_TX = 6,        // Verifier::verify_update_account_dark_tx_verifier    verifier.rs:818-917
    DDH = 7             // DDHProof::verify_ddh_proof                          shuffle/ddh.rs:109-142
};

// inputs of one call, in the memory space the phase runs in.  a0 / a1: account arrays (nproofs x n x 128 B), or for DDH g / h;
// p0 / p1: 32-byte arrays (SAME_VALUE commitments; DDH g_dash / h_dash); z0 / z1 / z2: response arrays; x: nproofs x 32 B.
struct inputs {
    const uint8_t *a0, *a1, *p0, *p1, *z0, *z1, *z2, *x, *base_pk;
    uint32_t n;          // accounts per proof (1 for SAME_VALUE, DDH)
    int vector_form;     // ZERO_BALANCE only
    int kind;
};
QQ_HOSTDEV static inline uint32_t msms_per_proof(int kind, uint32_t n) {
    return kind == DLOG ? 2 * n : kind == ZERO_BALANCE || kind == DESTROY ? 2 * n : kind == SAME_VALUE || kind == DDH ? 2 : 4 * n;
}
QQ_HOSTDEV static inline uint32_t terms_per_proof(int kind, uint32_t n) {
    return kind == DLOG ? 6 * n : kind == ZERO_BALANCE || kind == DESTROY ? 4 * n : kind == SAME_VALUE ? 6 : kind == DDH ? 4 : 10 * n;
}

struct emitter {
    uint8_t *sc, *pt;
    uint32_t* offs;      // this proof's CSR entries (msms_per_proof of them)
    uint32_t t;          // next term (global index)
    uint32_t m;
    // 32 bytes, as two 16-byte moves when both ends are 16-byte aligned (device: the term arrays and the uploaded inputs are;
    // a byte-wise memcpy made k_sigma_emit - 54 terms per DLOG proof - as slow as the transcripts: 0.32 ms)
    QQ_HOSTDEV static void copy32(uint8_t* d, const uint8_t* s) {
#ifdef __CUDA_ARCH__
        if (((reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(s)) & 15) == 0) {
            const uint4 a = reinterpret_cast<const uint4*>(s)[0], b = reinterpret_cast<const uint4*>(s)[1];
            reinterpret_cast<uint4*>(d)[0] = a;
            reinterpret_cast<uint4*>(d)[1] = b;
            return;
        }
#endif
        memcpy(d, s, 32);
    }
    QQ_HOSTDEV emitter& term(const uint8_t* scalar, const uint8_t* point) {
        copy32(sc + 32 * (size_t)t, scalar);
        copy32(pt + 32 * (size_t)t, point);
        t++;
        return *this;
    }
    QQ_HOSTDEV void begin() { offs[m++] = t; }
};

// the MSM jobs of proof p.  sc / pt: the batch's term arrays; offs: CSR array of the batch (offs[msms] is written by the caller).
QQ_HOSTDEV static inline void emit(const inputs& in, size_t p, uint8_t* sc, uint8_t* pt, uint32_t* offs) {
    const uint32_t n = in.n, mp = msms_per_proof(in.kind, n), tp = terms_per_proof(in.kind, n);
    emitter e{sc, pt, offs + p * mp, (uint32_t)(p * tp), 0};
    const uint8_t* xp = in.x + 32 * p;
    alignas(16) uint8_t negx[32];
    qq_merlin::sc_negate(negx, xp);     // a non-canonical x is caught by the MSM's scalar check on x itself
    switch (in.kind) {
    case DLOG:
        // a_i = delta_i.comm - input_i.comm;  e11_i = z_i input_i.gr + x a_i.c,  e12_i = z_i input_i.grsk + x a_i.d, as 3-term MSMs
        // (z, x, -x) x (key, delta.comm, input.comm): the difference is never encoded and decoded again
        for (uint32_t i = 0; i < n; i++) {
            const size_t it = p * n + i;
            const uint8_t *ia = in.a0 + 128 * it, *de = in.a1 + 128 * it;
            for (int h = 0; h < 2; h++) {
                e.begin();
                e.term(in.z0 + 32 * it, ia + 32 * h).term(xp, de + 64 + 32 * h).term(negx, ia + 64 + 32 * h);
            }
        }
        break;
    case DELTA_COMPACT:
        // e = zr gr + x c,  f = zr grsk + x d + zv B,  for the delta account (zr1) and the epsilon account (zr2)
        for (uint32_t i = 0; i < n; i++) {
            const size_t it = p * n + i;
            for (int side = 0; side < 2; side++) {
                const uint8_t* acc = (side ? in.a1 : in.a0) + 128 * it;
                const uint8_t* zr = (side ? in.z2 : in.z1) + 32 * it;
                e.begin();
                e.term(zr, acc).term(xp, acc + 64);
                e.begin();
                e.term(zr, acc + 32).term(xp, acc + 96).term(in.z0 + 32 * it, in.base_pk);
            }
        }
        break;
    case ACCOUNT:
        // e_delta = zsk gr_d + x grsk_d;  f_delta = zv G + zsk c_d + x d_d;  e_epsilon = x c_e + zr G;  f_epsilon = zv G + zr H + x d_e
        for (uint32_t i = 0; i < n; i++) {
            const size_t it = p * n + i;
            const uint8_t *d = in.a0 + 128 * it, *ep = in.a1 + 128 * it;
            const uint8_t *zv = in.z0 + 32 * it, *zsk = in.z1 + 32 * it, *zr = in.z2 + 32 * it;
            e.begin();
            e.term(zsk, d).term(xp, d + 32);
            e.begin();
            e.term(zv, in.base_pk).term(zsk, d + 64).term(xp, d + 96);
            e.begin();
            e.term(xp, ep + 64).term(zr, in.base_pk);
            e.begin();
            e.term(zv, in.base_pk).term(zr, in.base_pk + 32).term(xp, ep + 96);
        }
        break;
    case ZERO_BALANCE:
        // e = z gr + x c,  f = z grsk + x d
        for (uint32_t i = 0; i < n; i++) {
            const size_t it = p * n + i;
            const uint8_t* a = in.a0 + 128 * it;
            e.begin();
            e.term(in.z0 + 32 * it, a).term(xp, a + 64);
            e.begin();
            e.term(in.z0 + 32 * it, a + 32).term(xp, a + 96);
        }
        break;
    case DESTROY:
        // e = z gr + x grsk,  f = z c + x d
        for (uint32_t i = 0; i < n; i++) {
            const size_t it = p * n + i;
            const uint8_t* a = in.a0 + 128 * it;
            e.begin();
            e.term(in.z0 + 32 * it, a).term(xp, a + 32);
            e.begin();
            e.term(in.z0 + 32 * it, a + 64).term(xp, a + 96);
        }
        break;
    case SAME_VALUE: {
        // f_enc = zr grsk + x d + zv B;  f_pedersen = zr B_blinding + x commitment + zv B
        const uint8_t* a = in.a0 + 128 * p;
        e.begin();
        e.term(in.z1 + 32 * p, a + 32).term(xp, a + 96).term(in.z0 + 32 * p, in.base_pk);
        e.begin();
        e.term(in.z1 + 32 * p, in.base_pk + 32).term(xp, in.p0 + 32 * p).term(in.z0 + 32 * p, in.base_pk);
        break;
    }
    case DARK_TX: {
        // the reference's order: every e (e_gr = z0 gr_d + x gr_o, e_grsk), then every f (f_c = z1 gr_d + x (o.c - d.c), f_d) as 3-term MSMs
        const uint8_t *z0 = in.z0 + 64 * p, *z1 = in.z0 + 64 * p + 32;
        for (uint32_t i = 0; i < n; i++) {
            const uint8_t *d = in.a0 + 128 * (p * n + i), *o = in.a1 + 128 * (p * n + i);
            e.begin();
            e.term(z0, d).term(xp, o);
            e.begin();
            e.term(z0, d + 32).term(xp, o + 32);
        }
        for (uint32_t i = 0; i < n; i++) {
            const uint8_t *d = in.a0 + 128 * (p * n + i), *o = in.a1 + 128 * (p * n + i);
            e.begin();
            e.term(z1, d).term(xp, o + 64).term(negx, d + 64);
            e.begin();
            e.term(z1, d + 32).term(xp, o + 96).term(negx, d + 96);
        }
        break;
    }
    case DDH:
        // g_r = z G + challenge G_dash,  h_r = z H + challenge H_dash     (x = the proof's challenge, z0 = its response)
        e.begin();
        e.term(in.z0 + 32 * p, in.a0 + 32 * p).term(xp, in.p0 + 32 * p);
        e.begin();
        e.term(in.z0 + 32 * p, in.a1 + 32 * p).term(xp, in.p1 + 32 * p);
        break;
    }
}

// status of proof p from its MSMs (e: msms x 32 B encodings, st: msms status bytes) and its transcript.  tr: Transcript::new +
// Verifier::new.  A non-canonical scalar anywhere wins; an undecodable point gives what the reference's control flow gives there
// (Err("... Failed") = QQ_ST_BAD_POINT; the `unwrap()` of `d.comm - i.comm` in the dark-tx verifier = QQ_ST_PANIC).
// Returns true when the transcript ran to the challenge (its state is then what a following range proof continues from).
QQ_HOSTDEV static inline bool finish(const inputs& in, size_t p, qq_merlin::transcript& tr, const uint8_t* e, const uint8_t* st,
                                     uint8_t& status) {
    const uint32_t n = in.n, mp = msms_per_proof(in.kind, n);
    uint32_t first_bad = mp;
    for (uint32_t m = 0; m < mp; m++) {
        if (st[m] == QQ_ST_BAD_SCALAR) {
            status = QQ_ST_BAD_SCALAR;
            return false;
        }
        if (st[m] && first_bad == mp) first_bad = m;
    }
    if (first_bad != mp) {
        status = (in.kind == DARK_TX && first_bad >= 2 * n) ? QQ_ST_PANIC : QQ_ST_BAD_POINT;
        return false;
    }
    const char* chal_label = "challenge";
    switch (in.kind) {
    case DLOG:
        tr.domain_sep("DLOGProof");
        for (uint32_t i = 0; i < n; i++) {
            const uint8_t *ia = in.a0 + 128 * (p * n + i), *de = in.a1 + 128 * (p * n + i);
            tr.append_point_var("inputgr", ia);
            tr.append_point_var("inputgrsk", ia + 32);
            tr.append_point_var("outputgr", de);
            tr.append_point_var("outputgrsk", de + 32);
        }
        for (uint32_t i = 0; i < n; i++) {
            tr.append_point_var("commitmentgr", e + 64 * i);
            tr.append_point_var("commitmentgrsk", e + 64 * i + 32);
        }
        chal_label = "chal";
        break;
    case DELTA_COMPACT:
    case ACCOUNT:
        tr.domain_sep(in.kind == DELTA_COMPACT ? "VerifyDeltaCompact" : "VerifyAccountProof");
        for (uint32_t i = 0; i < n; i++) {
            tr.append_account_var("delta_account", in.a0 + 128 * (p * n + i));
            tr.append_account_var("epsilon_account", in.a1 + 128 * (p * n + i));
        }
        for (uint32_t i = 0; i < n; i++) {
            const uint8_t* q = e + 128 * i;
            tr.append_point_var("e_delta", q);
            tr.append_point_var("f_delta", q + 32);
            tr.append_point_var("e_epsilon", q + 64);
            tr.append_point_var("f_epsilon", q + 96);
        }
        break;
    case ZERO_BALANCE:
        // the vector form keeps the reference verifier's spelling b"ZeroBalanceAccounVectorProof" (verifier.rs:605)
        tr.domain_sep(in.vector_form ? "ZeroBalanceAccounVectorProof" : "ZeroBalanceAccountProof");
        for (uint32_t i = 0; i < n; i++)
            tr.append_account_var(in.vector_form ? "anonymity_account" : "zero_account", in.a0 + 128 * (p * n + i));
        for (uint32_t i = 0; i < n; i++) {
            tr.append_point_var("e", e + 64 * i);
            tr.append_point_var("f", e + 64 * i + 32);
        }
        break;
    case DESTROY:
        tr.domain_sep("DestroyAccountProof");
        for (uint32_t i = 0; i < n; i++) tr.append_account_var("account", in.a0 + 128 * (p * n + i));
        for (uint32_t i = 0; i < n; i++) {
            tr.append_point_var("e", e + 64 * i);
            tr.append_point_var("f", e + 64 * i + 32);
        }
        break;
    case SAME_VALUE:
        tr.append_account_var("encrypted_account", in.a0 + 128 * p);
        tr.append_point_var("G", in.base_pk);
        tr.append_point_var("H", in.base_pk + 32);
        tr.append_point_var("d", in.p0 + 32 * p);
        tr.append_point_var("f_delta", e);
        tr.append_point_var("f_epsilon", e + 32);
        break;
    case DARK_TX:
        tr.domain_sep("VerifyUpdateAccountDarkTx");
        for (uint32_t i = 0; i < n; i++) {
            tr.append_account_var("account", in.a0 + 128 * (p * n + i));
            tr.append_account_var("updatedaccount", in.a1 + 128 * (p * n + i));
        }
        for (uint32_t i = 0; i < n; i++) {
            tr.append_point_var("commitmentgr", e + 64 * i);
            tr.append_point_var("commitmentgrsk", e + 64 * i + 32);
        }
        for (uint32_t i = 0; i < n; i++) {
            tr.append_point_var("commitmentc", e + 64 * (n + i));
            tr.append_point_var("commitmentd", e + 64 * (n + i) + 32);
        }
        break;
    case DDH:
        tr.domain_sep("DDHTupleProof");
        tr.append_point_var("g", in.a0 + 32 * p);
        tr.append_point_var("g_dash", in.p0 + 32 * p);
        tr.append_point_var("h", in.a1 + 32 * p);
        tr.append_point_var("h_dash", in.p1 + 32 * p);
        tr.append_point_var("gr", e);
        tr.append_point_var("hr", e + 32);
        chal_label = "Challenge";
        break;
    }
    uint8_t chal[32];
    tr.get_challenge(chal_label, chal);
    status = qq_shuffle::differ32(chal, in.x + 32 * p) ? QQ_ST_PROOF : QQ_ST_OK;
    return true;
}

}  // namespace qq_sigma

#ifdef __CUDACC__
namespace qq_sigma {
// one thread per proof
__global__ void __launch_bounds__(64) k_sigma_emit(inputs in, size_t nproofs, uint8_t* __restrict__ sc, uint8_t* __restrict__ pt,
                                                   uint32_t* __restrict__ offs) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nproofs) return;
    emit(in, p, sc, pt, offs);
    if (p == nproofs - 1) offs[nproofs * msms_per_proof(in.kind, in.n)] = (uint32_t)(nproofs * terms_per_proof(in.kind, in.n));
}
// capture (may be nullptr): capture_n serialised transcript states, written for proofs whose transcript ran to the challenge
__global__ void __launch_bounds__(32) k_sigma_finish(inputs in, size_t nproofs, qq_merlin::transcript tr0, const uint8_t* __restrict__ e,
                                                     const uint8_t* __restrict__ st, uint8_t* __restrict__ status,
                                                     uint8_t* __restrict__ capture, size_t capture_n) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nproofs) return;
    const uint32_t mp = msms_per_proof(in.kind, in.n);
    qq_merlin::transcript tr = tr0;
    uint8_t s = QQ_ST_PROOF;
    bool ran = finish(in, p, tr, e + 32 * (size_t)mp * p, st + (size_t)mp * p, s);
    status[p] = s;
    if (capture != nullptr && p < capture_n && ran) tr.export_state(capture + (size_t)qq_merlin::transcript::STATE_BYTES * p);
}
}  // namespace qq_sigma
#endif
