// quisquis.hpp -- host-side mirror (C++17, header-only) of the reference's operator interface for the hot path, above
// the C ABI of include/qq_b200.h.  Same type names, method names, argument meaning and error behaviour as
//   src/keys.rs:33-126 (trait PublicKey), src/ristretto/keys.rs:76-282 (RistrettoPublicKey),
//   src/elgamal/elgamal.rs:18-236 (ElGamalCommitment), src/accounts/accounts.rs:47-347 (Account),
//   src/accounts/verifier.rs:91-99,138-917 (Verifier, every sigma-protocol verifier)
// plus the *_batch forms the Rust shim adds.  The reference is Rust; Rust is not available in this
// image, so this is the compiled host language closest to it.  Link with -lqq_b200.  No CPU fallback.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/qq_b200.h"

namespace quisquis {

using Scalar = std::array<uint8_t, 32>;               // canonical little-endian, < l
using CompressedRistretto = std::array<uint8_t, 32>;

// the reference panics (`.unwrap()` on a failed decompress); C++ callers get this exception instead
struct Panic : std::runtime_error {
    Panic() : std::runtime_error("called `Option::unwrap()` on a `None` value") {}
};
// Err(&'static str) of the reference
struct Err : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class Gpu {
  public:
    explicit Gpu(int device = 0) {
        int rc = qq_init(&ctx_, device);
        if (rc != QQ_OK) throw std::runtime_error("qq_init failed (" + std::to_string(rc) + "): no usable sm_100 GPU, no CPU fallback");
    }
    ~Gpu() { qq_destroy(ctx_); }
    Gpu(const Gpu&) = delete;
    Gpu& operator=(const Gpu&) = delete;
    qq_ctx* ctx() const { return ctx_; }
    void check(int rc, const char* what) const {
        if (rc != QQ_OK) throw std::runtime_error(std::string(what) + ": " + qq_last_error(ctx_));
    }
    static Gpu& instance() {
        static Gpu g(0);
        return g;
    }

  private:
    qq_ctx* ctx_ = nullptr;
};

struct RistrettoSecretKey {
    Scalar s;
};

struct RistrettoPublicKey {
    CompressedRistretto gr, grsk;  // src/ristretto/keys.rs:76-79

    std::array<uint8_t, 64> as_bytes() const {  // :113-120
        std::array<uint8_t, 64> b;
        std::memcpy(b.data(), gr.data(), 32);
        std::memcpy(b.data() + 32, grsk.data(), 32);
        return b;
    }
    static RistrettoPublicKey from_bytes(const uint8_t* p) {  // :127-134 (does not validate)
        RistrettoPublicKey k;
        std::memcpy(k.gr.data(), p, 32);
        std::memcpy(k.grsk.data(), p + 32, 32);
        return k;
    }
    static RistrettoPublicKey generate_base_pk() {  // :171-177 -> BASE_PK_BTC_COMPRESSED
        static const uint8_t B[64] = {
            0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f,
            0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76,
            0x8c, 0x92, 0x40, 0xb4, 0x56, 0xa9, 0xe6, 0xdc, 0x65, 0xc3, 0x77, 0xa1, 0x04, 0x8d, 0x74, 0x5f,
            0x94, 0xa0, 0x8c, 0xdb, 0x7f, 0x44, 0xcb, 0xcd, 0x7b, 0x46, 0xf3, 0x40, 0x48, 0x87, 0x11, 0x34};
        return from_bytes(B);
    }
    // update_public_key(p, rscalar) = (r*gr, r*grsk)   :146-148
    static RistrettoPublicKey update_public_key(const RistrettoPublicKey& p, const Scalar& rscalar) {
        auto in = p.as_bytes();
        uint8_t out[64], st;
        Gpu& g = Gpu::instance();
        g.check(qq_update_public_key_batch(g.ctx(), in.data(), rscalar.data(), out, &st, 1), "update_public_key");
        if (st == QQ_ST_BAD_POINT) throw Panic();
        return from_bytes(out);
    }
    static std::vector<RistrettoPublicKey> update_public_key_batch(const std::vector<RistrettoPublicKey>& p,
                                                                   const std::vector<Scalar>& r) {
        size_t n = p.size();
        std::vector<uint8_t> in(n * 64), sc(n * 32), out(n * 64), st(n);
        for (size_t i = 0; i < n; i++) {
            auto b = p[i].as_bytes();
            std::memcpy(&in[i * 64], b.data(), 64);
            std::memcpy(&sc[i * 32], r[i].data(), 32);
        }
        Gpu& g = Gpu::instance();
        g.check(qq_update_public_key_batch(g.ctx(), in.data(), sc.data(), out.data(), st.data(), n), "update_public_key_batch");
        std::vector<RistrettoPublicKey> res(n);
        for (size_t i = 0; i < n; i++) {
            if (st[i] == QQ_ST_BAD_POINT) throw Panic();
            res[i] = from_bytes(&out[i * 64]);
        }
        return res;
    }
    // verify_public_key_update(u, p, rscalar) -> bool   :161-169
    static bool verify_public_key_update(const RistrettoPublicKey& u, const RistrettoPublicKey& p, const Scalar& rscalar) {
        auto ub = u.as_bytes(), pb = p.as_bytes();
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_verify_public_key_update_batch(g.ctx(), ub.data(), pb.data(), rscalar.data(), &st, 1), "verify_public_key_update");
        if (st == QQ_ST_BAD_POINT) throw Panic();
        return st == QQ_ST_OK;
    }
    // verify_keypair(&self, privkey) -> Result<(), &'static str>   :187-195
    void verify_keypair(const RistrettoSecretKey& sk) const {
        uint8_t acc[128] = {0}, st;
        auto b = as_bytes();
        std::memcpy(acc, b.data(), 64);
        Scalar zero{};
        Gpu& g = Gpu::instance();
        g.check(qq_verify_account_batch(g.ctx(), acc, sk.s.data(), zero.data(), &st, 1), "verify_keypair");
        if (st == QQ_ST_BAD_POINT) throw Err("Error::Decompression Failed");
        if (st == QQ_ST_KEYPAIR) throw Err("Invalid Account::Keypair Verification Failed");
    }
    bool operator==(const RistrettoPublicKey& o) const { return gr == o.gr && grsk == o.grsk; }  // byte equality :241-247
};

struct ElGamalCommitment {
    CompressedRistretto c, d;  // src/elgamal/elgamal.rs:18-22

    std::array<uint8_t, 64> to_bytes() const {
        std::array<uint8_t, 64> b;
        std::memcpy(b.data(), c.data(), 32);
        std::memcpy(b.data() + 32, d.data(), 32);
        return b;
    }
    static ElGamalCommitment from_raw(const uint8_t* p) {
        ElGamalCommitment k;
        std::memcpy(k.c.data(), p, 32);
        std::memcpy(k.d.data(), p + 32, 32);
        return k;
    }
    // generate_commitment(p, rscalar, bl_scalar) = (r*gr, bl*B + r*grsk)   :41-53
    static ElGamalCommitment generate_commitment(const RistrettoPublicKey& p, const Scalar& rscalar, const Scalar& bl_scalar) {
        auto pb = p.as_bytes();
        uint8_t out[64], st;
        Gpu& g = Gpu::instance();
        g.check(qq_generate_commitment_batch(g.ctx(), pb.data(), rscalar.data(), bl_scalar.data(), out, &st, 1), "generate_commitment");
        if (st == QQ_ST_BAD_POINT) throw Panic();
        return from_raw(out);
    }
    // add_commitments(a, b)   :65-69
    static ElGamalCommitment add_commitments(const ElGamalCommitment& a, const ElGamalCommitment& b) { return addsub(a, b, 0); }
    // impl Sub   :201-218
    ElGamalCommitment operator-(const ElGamalCommitment& o) const { return addsub(*this, o, 1); }
    // impl Mul<&Scalar>   :220-236
    ElGamalCommitment operator*(const Scalar& s) const {
        auto b = to_bytes();
        uint8_t out[64], st;
        Gpu& g = Gpu::instance();
        g.check(qq_mul_commitment_batch(g.ctx(), b.data(), s.data(), out, &st, 1), "mul_commitment");
        if (st == QQ_ST_BAD_POINT) throw Panic();
        return from_raw(out);
    }
    // decommit(&self, pr) -> CompressedRistretto   :106-108
    CompressedRistretto decommit(const RistrettoSecretKey& pr) const {
        auto b = to_bytes();
        CompressedRistretto out;
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_decommit_batch(g.ctx(), b.data(), pr.s.data(), out.data(), &st, 1), "decommit");
        if (st == QQ_ST_BAD_POINT) throw Panic();
        return out;
    }
    // decommit_value(&self, pr) -> Option<Scalar>   :119-122 (search space 2^search_bits instead of the reference's 2^64 walk)
    std::optional<uint64_t> decommit_value(const RistrettoSecretKey& pr, int search_bits = 40) const {
        auto b = to_bytes();
        uint64_t v = 0;
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_decommit_value_batch(g.ctx(), b.data(), pr.s.data(), search_bits, &v, &st, 1), "decommit_value");
        if (st == QQ_ST_BAD_POINT) throw Panic();
        if (st != QQ_ST_OK) return std::nullopt;
        return v;
    }
    bool operator==(const ElGamalCommitment& o) const { return c == o.c && d == o.d; }

  private:
    static ElGamalCommitment addsub(const ElGamalCommitment& a, const ElGamalCommitment& b, int neg) {
        auto ab = a.to_bytes(), bb = b.to_bytes();
        uint8_t out[64], st;
        Gpu& g = Gpu::instance();
        g.check(qq_add_commitments_batch(g.ctx(), ab.data(), bb.data(), neg, out, &st, 1), "add_commitments");
        if (st == QQ_ST_BAD_POINT) throw Panic();
        return from_raw(out);
    }
};

struct Account {
    RistrettoPublicKey pk;
    ElGamalCommitment comm;  // src/accounts/accounts.rs:47-53

    static Account set_account(const RistrettoPublicKey& pk, const ElGamalCommitment& comm) { return Account{pk, comm}; }
    std::array<uint8_t, 128> to_bytes() const {
        std::array<uint8_t, 128> b;
        auto p = pk.as_bytes();
        auto c = comm.to_bytes();
        std::memcpy(b.data(), p.data(), 64);
        std::memcpy(b.data() + 64, c.data(), 64);
        return b;
    }
    static Account from_raw(const uint8_t* p) { return Account{RistrettoPublicKey::from_bytes(p), ElGamalCommitment::from_raw(p + 64)}; }

    // update_account(a, bl, update_key_scalar, generate_commitment_scalar)   :143-154
    static Account update_account(const Account& a, const Scalar& bl, const Scalar& update_key_scalar,
                                  const Scalar& generate_commitment_scalar) {
        return update_account_batch({a}, {bl}, {update_key_scalar}, {generate_commitment_scalar})[0];
    }
    static std::vector<Account> update_account_batch(const std::vector<Account>& a, const std::vector<Scalar>& bl,
                                                     const std::vector<Scalar>& u, const std::vector<Scalar>& c) {
        size_t n = a.size();
        std::vector<uint8_t> acc(n * 128), b(n * 32), uu(n * 32), cc(n * 32), out(n * 128), st(n);
        for (size_t i = 0; i < n; i++) {
            auto ab = a[i].to_bytes();
            std::memcpy(&acc[i * 128], ab.data(), 128);
            std::memcpy(&b[i * 32], bl[i].data(), 32);
            std::memcpy(&uu[i * 32], u[i].data(), 32);
            std::memcpy(&cc[i * 32], c[i].data(), 32);
        }
        Gpu& g = Gpu::instance();
        g.check(qq_update_account_batch(g.ctx(), acc.data(), b.data(), uu.data(), cc.data(), out.data(), st.data(), n), "update_account");
        std::vector<Account> res(n);
        for (size_t i = 0; i < n; i++) {
            if (st[i] == QQ_ST_BAD_POINT) throw Panic();
            res[i] = from_raw(&out[i * 128]);
        }
        return res;
    }
    // verify_account(&self, sk, bl) -> Result<(), &'static str>   :81-84
    void verify_account(const RistrettoSecretKey& sk, const Scalar& bl) const {
        auto ab = to_bytes();
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_verify_account_batch(g.ctx(), ab.data(), sk.s.data(), bl.data(), &st, 1), "verify_account");
        if (st == QQ_ST_BAD_POINT) throw Err("Error::Decompression Failed");
        if (st == QQ_ST_KEYPAIR) throw Err("Invalid Account::Keypair Verification Failed");
        if (st == QQ_ST_COMMIT) throw Err("Invalid Account::Commitment Verification Failed");
    }
    // verify_account_update: exactly 9 accounts, bl = 0 (quirk kept: shorter input is an index panic)   :173-193
    static bool verify_account_update(const std::vector<Account>& updated_input_accounts, const std::vector<Account>& accounts,
                                      const std::vector<Scalar>& updated_keys_scalar,
                                      const std::vector<Scalar>& generate_commitment_scalar) {
        if (accounts.size() < 9 || updated_keys_scalar.size() < 9 || generate_commitment_scalar.size() < 9)
            throw std::out_of_range("index out of bounds: the len is < 9");
        std::vector<Account> a(accounts.begin(), accounts.begin() + 9);
        std::vector<Scalar> u(updated_keys_scalar.begin(), updated_keys_scalar.begin() + 9);
        std::vector<Scalar> c(generate_commitment_scalar.begin(), generate_commitment_scalar.begin() + 9);
        std::vector<Scalar> zero(9, Scalar{});
        auto upd = update_account_batch(a, zero, u, c);
        size_t m = std::min<size_t>(9, updated_input_accounts.size());
        for (size_t i = 0; i < m; i++)
            if (!(upd[i] == updated_input_accounts[i])) return false;
        return true;
    }
    // create_delta_and_epsilon_accounts(a, bl, base_pk) with rscalar supplied by the caller   :198-220
    static std::pair<std::vector<Account>, std::vector<Account>> create_delta_and_epsilon_accounts(
        const std::vector<Account>& a, const std::vector<Scalar>& bl, const RistrettoPublicKey& base_pk,
        const std::vector<Scalar>& rscalar) {
        size_t n = a.size();
        std::vector<uint8_t> acc(n * 128), b(n * 32), r(n * 32), d(n * 128), e(n * 128), st(n);
        for (size_t i = 0; i < n; i++) {
            auto ab = a[i].to_bytes();
            std::memcpy(&acc[i * 128], ab.data(), 128);
            std::memcpy(&b[i * 32], bl[i].data(), 32);
            std::memcpy(&r[i * 32], rscalar[i].data(), 32);
        }
        auto bp = base_pk.as_bytes();
        Gpu& g = Gpu::instance();
        g.check(qq_delta_epsilon_batch(g.ctx(), acc.data(), b.data(), r.data(), bp.data(), d.data(), e.data(), st.data(), n), "delta_epsilon");
        std::vector<Account> dv(n), ev(n);
        for (size_t i = 0; i < n; i++) {
            if (st[i] == QQ_ST_BAD_POINT) throw Panic();
            dv[i] = from_raw(&d[i * 128]);
            ev[i] = from_raw(&e[i * 128]);
        }
        return {dv, ev};
    }
    bool operator==(const Account& o) const { return pk == o.pk && comm == o.comm; }  // :350-356
};

struct Verifier {
    // multiscalar_multiplication(&scalars, &points) -> Option<RistrettoPoint> (returned compressed)   verifier.rs:91-99
    static std::optional<CompressedRistretto> multiscalar_multiplication(const std::vector<Scalar>& combined_scalars,
                                                                         const std::vector<CompressedRistretto>& point) {
        size_t n = combined_scalars.size();
        std::vector<uint8_t> s(n * 32), p(n * 32);
        for (size_t i = 0; i < n; i++) {
            std::memcpy(&s[i * 32], combined_scalars[i].data(), 32);
            std::memcpy(&p[i * 32], point[i].data(), 32);
        }
        CompressedRistretto out;
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_msm(g.ctx(), s.data(), p.data(), n, out.data(), &st), "multiscalar_multiplication");
        if (st != QQ_ST_OK) return std::nullopt;
        return out;
    }
    // verify_delta_identity_check(&[Account]) -> Result<(), &'static str>   verifier.rs:566-581
    static void verify_delta_identity_check(const std::vector<Account>& epsilon_accounts) {
        size_t n = epsilon_accounts.size();
        std::vector<uint8_t> acc(n * 128);
        for (size_t i = 0; i < n; i++) {
            auto ab = epsilon_accounts[i].to_bytes();
            std::memcpy(&acc[i * 128], ab.data(), 128);
        }
        uint8_t v;
        Gpu& g = Gpu::instance();
        g.check(qq_delta_identity_check(g.ctx(), acc.data(), n, &v), "verify_delta_identity_check");
        if (v == QQ_ST_BAD_POINT) throw Panic();
        if (v != QQ_ST_OK) throw Err("Identity sum verify: Failed");
    }
    // ---- sigma-protocol verifiers (src/accounts/verifier.rs:138-917).  The reference passes a `Verifier` carrying a Merlin
    // transcript; here the two labels that built it (Transcript::new(label), Verifier::new(label, ..)) are the trailing
    // arguments, defaulting to the labels of the reference's own tests.  Ok(()) = return, Err(msg) = throw Err(msg).
    static std::vector<uint8_t> pack(const std::vector<Account>& v) {
        std::vector<uint8_t> b(v.size() * 128);
        for (size_t i = 0; i < v.size(); i++) {
            auto ab = v[i].to_bytes();
            std::memcpy(&b[i * 128], ab.data(), 128);
        }
        return b;
    }
    static std::vector<uint8_t> pack(const std::vector<Scalar>& v) {
        std::vector<uint8_t> b(v.size() * 32);
        for (size_t i = 0; i < v.size(); i++) std::memcpy(&b[i * 32], v[i].data(), 32);
        return b;
    }
    static void verdict(uint8_t st, const char* bad_point, const char* mismatch) {
        if (st == QQ_ST_PANIC) throw Panic();
        if (st == QQ_ST_BAD_POINT) {
            if (!bad_point) throw Panic();
            throw Err(bad_point);
        }
        if (st != QQ_ST_OK) throw Err(mismatch);
    }
    // verifier.rs:138-209
    static void verify_delta_compact_verifier(const std::vector<Account>& delta_accounts, const std::vector<Account>& epsilon_accounts,
                                              const std::vector<Scalar>& zv_vector, const std::vector<Scalar>& zr1_vector,
                                              const std::vector<Scalar>& zr2_vector, const Scalar& x,
                                              const char* transcript_label = "DeltaCompact", const char* verifier_label = "DLEQProof") {
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_verify_delta_compact_batch(g.ctx(), transcript_label, verifier_label, pack(delta_accounts).data(),
                                              pack(epsilon_accounts).data(), pack(zv_vector).data(), pack(zr1_vector).data(),
                                              pack(zr2_vector).data(), x.data(), zv_vector.size(), 1, &st),
                "verify_delta_compact_verifier");
        verdict(st, "Delta Compact Proof Verify: Failed", "Dleq Proof Verify: Failed");
    }
    // verifier.rs:223-292
    static void verify_update_account_verifier(const std::vector<Account>& updated_input_accounts,
                                               const std::vector<Account>& updated_delta_accounts, const std::vector<Scalar>& z_vector,
                                               const Scalar& x, const char* transcript_label = "UpdateAccount",
                                               const char* verifier_label = "DLOGProof") {
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_verify_update_account_dlog_batch(g.ctx(), transcript_label, verifier_label, pack(updated_input_accounts).data(),
                                                    pack(updated_delta_accounts).data(), pack(z_vector).data(), x.data(),
                                                    z_vector.size(), 1, &st),
                "verify_update_account_verifier");
        verdict(st, nullptr, "DLOG Proof Verify: Failed");
    }
    // verifier.rs:396-470 (and the sigma-protocol part of verify_account_verifier, :305-381)
    static void verify_account_verifier_bulletproof(const std::vector<Account>& updated_delta_account_sender,
                                                    const std::vector<Account>& account_epsilon_sender, const RistrettoPublicKey& base_pk,
                                                    const std::vector<Scalar>& zv, const std::vector<Scalar>& zsk,
                                                    const std::vector<Scalar>& zr, const Scalar& x,
                                                    const char* transcript_label = "SenderAccountProof",
                                                    const char* verifier_label = "DLOGProof") {
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_verify_account_sigma_batch(g.ctx(), transcript_label, verifier_label, pack(updated_delta_account_sender).data(),
                                              pack(account_epsilon_sender).data(), base_pk.as_bytes().data(), pack(zv).data(),
                                              pack(zsk).data(), pack(zr).data(), x.data(), zv.size(), 1, &st),
                "verify_account_verifier_bulletproof");
        verdict(st, "Account Verify: Failed", "sender account verification failed");
    }
    // The reference keeps ONE running transcript per Verifier: verify_account_verifier_bulletproof and the range proof that
    // follows share it (verifier.rs:1603-1628).  keep_transcript() before the sigma verification, then pass the state on.
    using TranscriptState = std::vector<uint8_t>;
    static TranscriptState keep_transcript() {
        TranscriptState s(qq_transcript_state_bytes());
        Gpu& g = Gpu::instance();
        g.check(qq_transcript_capture(g.ctx(), s.data(), 1), "keep_transcript");
        return s;
    }
    // verifier.rs:504-523: one aggregated 64-bit range proof (RangeProof::to_bytes()) over the d components of the accounts
    static void verify_non_negative_sender_receiver_bulletproof_batch_verifier(const std::vector<Account>& epsilon_account,
                                                                               const std::vector<uint8_t>& proof,
                                                                               const TranscriptState* transcript = nullptr,
                                                                               const char* transcript_label = "SenderAccountProof",
                                                                               const char* verifier_label = "BulletProof") {
        size_t m = epsilon_account.size(), lg = 0;
        while (((size_t)1 << lg) < 64 * m) lg++;
        if (m == 0 || m > 16 || (m & (m - 1)) || proof.size() != (9 + 2 * lg) * 32) throw Err("Bulletproof verification failed");
        std::vector<uint8_t> cm;
        for (const Account& a : epsilon_account) {
            auto b = a.to_bytes();
            cm.insert(cm.end(), b.begin() + 96, b.begin() + 128);
        }
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_verify_range_proof_batch(g.ctx(), transcript_label, verifier_label, transcript ? transcript->data() : nullptr,
                                            "AggregateBulletProof", cm.data(), proof.data(), 64, m, 1, 1, &st),
                "verify_non_negative_sender_receiver_bulletproof_batch_verifier");
        if (st) throw Err("Bulletproof verification failed");
    }
    // verifier.rs:534-555: one single-value proof per account, chained on one transcript
    static void verify_non_negative_sender_receiver_bulletproof_vector_verifier(const std::vector<Account>& epsilon_account,
                                                                                const std::vector<std::vector<uint8_t>>& proof_vector,
                                                                                const TranscriptState* transcript = nullptr,
                                                                                const char* transcript_label = "SenderAccountProof",
                                                                                const char* verifier_label = "BulletProof") {
        size_t k = std::min(epsilon_account.size(), proof_vector.size());      // zip() stops at the shorter one
        if (k == 0) return;
        std::vector<uint8_t> cm, pr;
        for (size_t i = 0; i < k; i++) {
            if (proof_vector[i].size() != (9 + 12) * 32) throw Err("Bulletproof verification failed");
            auto b = epsilon_account[i].to_bytes();
            cm.insert(cm.end(), b.begin() + 96, b.begin() + 128);
            pr.insert(pr.end(), proof_vector[i].begin(), proof_vector[i].end());
        }
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_verify_range_proof_batch(g.ctx(), transcript_label, verifier_label, transcript ? transcript->data() : nullptr,
                                            "AggregateBulletProof", cm.data(), pr.data(), 64, 1, k, 1, &st),
                "verify_non_negative_sender_receiver_bulletproof_vector_verifier");
        if (st) throw Err("Bulletproof verification failed");
    }
    // verifier.rs:593-634 (domain separator as the reference's verifier spells it)
    static void zero_balance_account_vector_verifier(const std::vector<Account>& anonymity_accounts, const std::vector<Scalar>& z,
                                                     const Scalar& x, const char* transcript_label = "ZeroBalanceAccount",
                                                     const char* verifier_label = "DLOGProof") {
        if (anonymity_accounts.size() != z.size()) throw std::logic_error("assertion failed: anonymity_accounts.len() == z.len()");
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_verify_zero_balance_batch(g.ctx(), transcript_label, verifier_label, pack(anonymity_accounts).data(),
                                             pack(z).data(), x.data(), z.size(), 1, 1, &st),
                "zero_balance_account_vector_verifier");
        verdict(st, "Zero balance Account Verify: Failed", "Zero balance account verification failed");
    }
    // verifier.rs:647-680
    static void zero_balance_account_verifier(const Account& account, const Scalar& z, const Scalar& x,
                                              const char* transcript_label = "ZeroBalanceAccount",
                                              const char* verifier_label = "DLOGProof") {
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_verify_zero_balance_batch(g.ctx(), transcript_label, verifier_label, account.to_bytes().data(), z.data(),
                                             x.data(), 1, 1, 0, &st),
                "zero_balance_account_verifier");
        verdict(st, "Zero balance Account Verify: Failed", "Zero balance account verification failed");
    }
    // verifier.rs:693-735
    static void destroy_account_verifier(const std::vector<Account>& accounts, const std::vector<Scalar>& z, const Scalar& x,
                                         const char* transcript_label = "DestroyAccount", const char* verifier_label = "DLOGProof") {
        if (accounts.size() != z.size()) throw std::logic_error("assertion failed: accounts.len() == z.len()");
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_verify_destroy_account_batch(g.ctx(), transcript_label, verifier_label, pack(accounts).data(), pack(z).data(),
                                                x.data(), z.size(), 1, &st),
                "destroy_account_verifier");
        verdict(st, "Destroy Account Verify: Failed", "Destroy account verification failed");
    }
    // verifier.rs:747-806; the proof is SigmaProof::Dleq(zv, zr, _, x) with one response each
    static void verify_same_value_compact_verifier(const Account& enc_account, const CompressedRistretto& commitment, const Scalar& zv,
                                                   const Scalar& zr, const Scalar& x) {
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_verify_same_value_compact_batch(g.ctx(), enc_account.to_bytes().data(), commitment.data(), zv.data(), zr.data(),
                                                   x.data(), 1, &st),
                "verify_same_value_compact_verifier");
        verdict(st, "Delta Compact Proof Verify: Failed", "Same Value Proof Verify: Failed");
    }
    // verifier.rs:818-917
    static void verify_update_account_dark_tx_verifier(const std::vector<Account>& delta_updated_accounts,
                                                       const std::vector<Account>& output_accounts, const std::vector<Scalar>& z_vector,
                                                       const Scalar& x, const char* transcript_label = "UpdateAccount",
                                                       const char* verifier_label = "DLOGProof") {
        if (delta_updated_accounts.size() != output_accounts.size())
            throw Err("Length of delta_updated_accounts and output_accounts is not same");
        uint8_t st;
        Gpu& g = Gpu::instance();
        g.check(qq_verify_update_account_dark_tx_batch(g.ctx(), transcript_label, verifier_label, pack(delta_updated_accounts).data(),
                                                       pack(output_accounts).data(), pack(z_vector).data(), x.data(),
                                                       output_accounts.size(), 1, &st),
                "verify_update_account_dark_tx_verifier");
        verdict(st, "Update Account: DLOG Proof Verify: Failed", "Update Output Challenge : DLOG Proof Verify: Failed");
    }
};

}  // namespace quisquis
