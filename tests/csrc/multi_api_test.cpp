// GPU test of the multi-device handle of the C ABI (include/qq_b200.h: qq_init_multi ...) in plain C++: the same batches through
// qq_multi over every visible GPU (at most 8; one is enough to run) and through a single qq_ctx must give identical bytes, and
// one MSM split over the devices must equal the single-device MSM.  Inputs are valid points made with qq_fixed_base_batch.
// Prints "MULTI_API_TEST OK devices=<n>" on success.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/qq_b200.h"

#define CHECK(c)                                                               \
    do {                                                                       \
        if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); return 1; } \
    } while (0)

static uint64_t rng_state = 0x5155495351554953ull;
static uint64_t next64() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}
static void rand_scalars(std::vector<uint8_t>& v, size_t n) {      // < 2^252: canonical
    v.resize(32 * n);
    for (size_t i = 0; i < 4 * n; i++) {
        uint64_t x = next64();
        std::memcpy(&v[8 * i], &x, 8);
    }
    for (size_t i = 0; i < n; i++) v[32 * i + 31] &= 0x0f;
}

int main(int argc, char** argv) {
    int want = argc > 1 ? std::atoi(argv[1]) : 8;
    qq_ctx* one = nullptr;
    CHECK(qq_init(&one, 0) == QQ_OK);
    // how many devices can be opened
    std::vector<int> devs;
    for (int d = 0; d < want; d++) {
        qq_ctx* probe = nullptr;
        if (d == 0 || qq_init(&probe, d) == QQ_OK) {
            devs.push_back(d);
            if (probe) qq_destroy(probe);
        } else {
            break;
        }
    }
    qq_multi* m = nullptr;
    int dup[2] = {0, 0};
    CHECK(qq_init_multi(&m, dup, 2) == QQ_ERR_ARG);
    CHECK(qq_init_multi(&m, devs.data(), (int)devs.size()) == QQ_OK);
    CHECK(qq_multi_device_count(m) == (int)devs.size());
    const size_t n = 3001;      // not a multiple of the device count
    std::vector<uint8_t> s[4], acc(128 * n), bl, u, c, st(n), out1(128 * n), outm(128 * n), st1(n), stm(n);
    for (int k = 0; k < 4; k++) {
        rand_scalars(s[k], n);
        std::vector<uint8_t> pts(32 * n);
        CHECK(qq_fixed_base_batch(one, QQ_BASE_B, s[k].data(), pts.data(), st.data(), n) == QQ_OK);
        for (size_t i = 0; i < n; i++) std::memcpy(&acc[128 * i + 32 * k], &pts[32 * i], 32);
    }
    rand_scalars(bl, n);
    rand_scalars(u, n);
    rand_scalars(c, n);
    acc[128 * 7 + 31] |= 0x80;     // one undecodable account (bit 255 set is never canonical): status 1 at the same place either way
    CHECK(qq_update_account_batch(one, acc.data(), bl.data(), u.data(), c.data(), out1.data(), st1.data(), n) == QQ_OK);
    CHECK(qq_multi_update_account_batch(m, acc.data(), bl.data(), u.data(), c.data(), outm.data(), stm.data(), n) == QQ_OK);
    CHECK(out1 == outm && st1 == stm);
    size_t bad = 0;
    for (size_t i = 0; i < n; i++) bad += st1[i] != 0;
    CHECK(bad >= 1 && st1[7] != 0);
    // generate_commitment
    std::vector<uint8_t> pk(64 * n), cm1(64 * n), cmm(64 * n);
    for (size_t i = 0; i < n; i++) std::memcpy(&pk[64 * i], &out1[128 * ((i + 8) % n == 7 ? 8 : (i + 8) % n)], 64);
    CHECK(qq_generate_commitment_batch(one, pk.data(), u.data(), bl.data(), cm1.data(), st1.data(), n) == QQ_OK);
    CHECK(qq_multi_generate_commitment_batch(m, pk.data(), u.data(), bl.data(), cmm.data(), stm.data(), n) == QQ_OK);
    CHECK(cm1 == cmm && st1 == stm);
    // one MSM over 70 001 points split over the devices == the single-device MSM (Pippenger on every slice)
    const size_t nm = 70001;
    std::vector<uint8_t> hs, as, pts(32 * nm), stp(nm);
    rand_scalars(hs, nm);
    rand_scalars(as, nm);
    CHECK(qq_fixed_base_batch(one, QQ_BASE_B, hs.data(), pts.data(), stp.data(), nm) == QQ_OK);
    uint8_t r1[32], rm[32], s1 = 9, sm = 9;
    CHECK(qq_msm(one, as.data(), pts.data(), nm, r1, &s1) == QQ_OK);
    CHECK(qq_multi_msm(m, as.data(), pts.data(), nm, rm, &sm) == QQ_OK);
    CHECK(s1 == 0 && sm == 0 && std::memcmp(r1, rm, 32) == 0);
    // tiny MSM (fewer terms than devices) and a bad point in the last slice
    CHECK(qq_msm(one, as.data(), pts.data(), 3, r1, &s1) == QQ_OK);
    CHECK(qq_multi_msm(m, as.data(), pts.data(), 3, rm, &sm) == QQ_OK);
    CHECK(s1 == 0 && sm == 0 && std::memcmp(r1, rm, 32) == 0);
    pts[32 * (nm - 1)] ^= 0x01;
    pts[32 * (nm - 1) + 31] |= 0x80;      // high bit set: never a valid encoding
    CHECK(qq_multi_msm(m, as.data(), pts.data(), nm, rm, &sm) == QQ_OK);
    CHECK(sm == QQ_ST_BAD_POINT);
    uint8_t zero[32] = {0};
    CHECK(std::memcmp(rm, zero, 32) == 0);
    qq_destroy_multi(m);
    qq_destroy(one);
    std::printf("MULTI_API_TEST OK devices=%d\n", (int)devs.size());
    return 0;
}
