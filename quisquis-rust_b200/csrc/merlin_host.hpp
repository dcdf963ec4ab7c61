// Merlin transcripts: STROBE-128 over Keccak-f[1600] (merlin.cool; crate `merlin`, a dependency of the
// reference: src/accounts/transcript.rs:10) and the reference's TranscriptProtocol extension (transcript.rs:55-82).
// Callable from host and device code: the sigma verifiers run one transcript per proof on the host threads, the shuffle
// verifier runs them one-thread-per-proof in its transcript kernels (shuffle_verify.cuh).  The group arithmetic that
// produces the points a transcript absorbs runs in the MSM kernels either way.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>

#include "keccak_host.hpp"
#include "sc_host.hpp"

namespace qq_merlin {

class transcript;

class strobe128 {
    friend class transcript;

  public:
    enum { FLAG_I = 1, FLAG_A = 2, FLAG_C = 4, FLAG_T = 8, FLAG_M = 16, FLAG_K = 32 };

  private:
    static const int R = 166;
    alignas(8) uint8_t st[200];
    uint8_t pos, pos_begin, cur_flags;

    QQ_HOSTDEV void permute() {
        qq_keccak::f1600(reinterpret_cast<uint64_t*>(st));   // little-endian host (x86-64 / aarch64 LE) and device
    }
    QQ_HOSTDEV void run_f() {
        st[pos] ^= pos_begin;
        st[pos + 1] ^= 0x04;
        st[R + 1] ^= 0x80;
        permute();
        pos = 0;
        pos_begin = 0;
    }
    // The state is addressed as 25 little-endian 64-bit words (what Keccak-f permutes): absorbing / squeezing moves up to
    // eight bytes per step with two shifts instead of one load-xor-store per byte.  On the device the message words of a
    // block of up to 32 bytes are loaded into registers BEFORE the state is touched: d is a generic pointer the compiler
    // cannot tell apart from st[], and interleaved it would serialise one global-memory latency per load.
    QQ_HOSTDEV uint64_t* words() { return reinterpret_cast<uint64_t*>(st); }
    QQ_HOSTDEV static uint64_t load_le(const uint8_t* d, unsigned c) {      // c <= 8 bytes, little endian
        if (c == 8 && (reinterpret_cast<uintptr_t>(d) & 7) == 0) return *reinterpret_cast<const uint64_t*>(d);
        uint64_t v = 0;
        for (unsigned i = 0; i < c; i++) v |= (uint64_t)d[i] << (8 * i);
        return v;
    }
    QQ_HOSTDEV void xor_at(unsigned at, uint64_t v, unsigned c) {           // c <= 8 bytes of v into st[at .. at + c)
        const unsigned q = at >> 3, sh = (at & 7) * 8;
        uint64_t* w = words();
        w[q] ^= v << sh;
        if (sh + 8 * c > 64) w[q + 1] ^= v >> (64 - sh);
    }
    QQ_HOSTDEV QQ_NOINLINE void absorb(const uint8_t* d, size_t n) {
        while (n) {
            const unsigned blk = n < 32 ? (unsigned)n : 32u;
            uint64_t m[4];
#pragma unroll
            for (unsigned k = 0; k < 4; k++) m[k] = 8 * k < blk ? load_le(d + 8 * k, blk - 8 * k < 8 ? blk - 8 * k : 8u) : 0;
#pragma unroll
            for (unsigned k = 0; k < 4; k++) {
                if (8 * k >= blk) break;
                unsigned c = blk - 8 * k < 8 ? blk - 8 * k : 8u;
                uint64_t v = m[k];
                while (c) {                                   // at most two rounds: a word that straddles the rate boundary
                    unsigned room = (unsigned)(R - pos), t = c < room ? c : room;
                    xor_at(pos, t == 8 ? v : (v & ((1ULL << (8 * t)) - 1)), t);
                    pos = (uint8_t)(pos + t);
                    if (pos == R) run_f();
                    v = t == 8 ? 0 : v >> (8 * t);
                    c -= t;
                }
            }
            d += blk;
            n -= blk;
        }
    }
    QQ_HOSTDEV QQ_NOINLINE void squeeze(uint8_t* d, size_t n) {
        while (n) {
            unsigned room = (unsigned)(R - pos), c = n < 8 ? (unsigned)n : 8u;
            if (c > room) c = room;
            const unsigned q = pos >> 3, sh = (pos & 7) * 8;
            uint64_t* w = words();
            uint64_t v = w[q] >> sh;
            const uint64_t mask = c == 8 ? ~0ULL : ((1ULL << (8 * c)) - 1);
            w[q] &= ~(mask << sh);                                          // squeezed bytes are zeroed (STROBE's PRF)
            if (sh + 8 * c > 64) {
                v |= w[q + 1] << (64 - sh);
                w[q + 1] &= ~(mask >> (64 - sh));
            }
            v &= mask;
            for (unsigned i = 0; i < c; i++) d[i] = (uint8_t)(v >> (8 * i));
            pos = (uint8_t)(pos + c);
            d += c;
            n -= c;
            if (pos == R) run_f();
        }
    }
    // ---- fused operations: one pass over "operation header | label | length | operation header | bytes" with the cursor in registers --
    // A Merlin append_message is meta-AD(label) | meta-AD(le32 length, continued) | AD(message): five separate absorbs of 2, |label|,
    // 4, 2 and n bytes.  Done through the member functions above, every piece reloads pos / pos_begin from the object (local memory
    // in the transcript kernels), and two neighbouring pieces read-modify-write the same state word through memory: chains of
    // dependent local-memory round trips (31 % of the transcript kernels' time for 15 % of their instructions, ncu source view).
    // The cursor keeps pos / pos_begin and the pending bytes of the current state word in registers: every state word is
    // loaded, xored and stored exactly once per operation, and the accesses of different words are independent.
    struct cursor {
        strobe128& s;
        uint64_t* w;
        unsigned pos, pos_begin;
        uint64_t acc;      // bytes absorbed since the last flush, at their place inside state word (pos >> 3)
        QQ_HOSTDEV explicit cursor(strobe128& st_) : s(st_), w(st_.words()), pos(st_.pos), pos_begin(st_.pos_begin), acc(0) {}
        QQ_HOSTDEV void flush_word(unsigned q) {
            w[q] ^= acc;
            acc = 0;
        }
        QQ_HOSTDEV void run_f() {                     // acc is empty here
            w[pos >> 3] ^= (uint64_t)pos_begin << (8 * (pos & 7));
            w[(pos + 1) >> 3] ^= (uint64_t)0x04 << (8 * ((pos + 1) & 7));
            w[(R + 1) >> 3] ^= (uint64_t)0x80 << (8 * ((R + 1) & 7));
            s.permute();
            pos = 0;
            pos_begin = 0;
        }
        // the low c <= 8 bytes of v (the bytes above them must be zero)
        QQ_HOSTDEV void put(uint64_t v, unsigned c) {
            while (c) {
                const unsigned o = pos & 7, room_word = 8 - o, room_rate = (unsigned)R - pos;
                unsigned t = c < room_word ? c : room_word;
                if (t > room_rate) t = room_rate;
                acc |= v << (8 * o);                  // bytes beyond the word (or beyond t) are cut by the shift / re-fed below
                if (t < 8 && o + t < 8) acc &= ~0ULL >> (8 * (8 - o - t));
                pos += t;
                if ((pos & 7) == 0 || pos == (unsigned)R) flush_word((pos - 1) >> 3);
                if (pos == (unsigned)R) run_f();
                v = t == 8 ? 0 : v >> (8 * t);
                c -= t;
            }
        }
        QQ_HOSTDEV void begin(uint8_t flags) {        // begin_op for flags without C / K (no forced permutation)
            const unsigned old_begin = pos_begin;
            pos_begin = pos + 1;
            put((uint64_t)old_begin | ((uint64_t)flags << 8), 2);
        }
        QQ_HOSTDEV void label_len(const char* label, uint32_t n) {
            uint64_t v = 0;
            unsigned cnt = 0;
            for (unsigned i = 0; label[i]; i++) {
                v |= (uint64_t)(uint8_t)label[i] << (8 * cnt);
                if (++cnt == 8) {
                    put(v, 8);
                    v = 0;
                    cnt = 0;
                }
            }
            if (cnt <= 4) {
                put(v | ((uint64_t)n << (8 * cnt)), cnt + 4);
            } else {
                put(v, cnt);
                put(n, 4);
            }
        }
        QQ_HOSTDEV void finish(uint8_t flags) {
            if (pos & 7) flush_word(pos >> 3);
            s.pos = (uint8_t)pos;
            s.pos_begin = (uint8_t)pos_begin;
            s.cur_flags = flags;
        }
    };
    // meta-AD(label | le32(n)) as one operation
    QQ_HOSTDEV QQ_NOINLINE void op_label_len(uint8_t flags, const char* label, uint32_t n) {
        cursor c(*this);
        c.begin(flags);
        c.label_len(label, n);
        c.finish(flags);
    }
    // the whole append_message: meta-AD(label | le32(n)), then AD(msg)
    QQ_HOSTDEV QQ_NOINLINE void op_append(const char* label, const uint8_t* d, size_t n) {
        // the common case (points, scalars, wide challenges): the message words are loaded before the state is touched - d is a
        // generic pointer the compiler cannot tell apart from st[], interleaved it would serialise one memory latency per load
        const bool words = n <= 64 && (n & 7) == 0 && (reinterpret_cast<uintptr_t>(d) & 7) == 0;
        uint64_t m[8];
        const unsigned nw = words ? (unsigned)(n >> 3) : 0;
#pragma unroll
        for (unsigned k = 0; k < 8; k++) m[k] = k < nw ? reinterpret_cast<const uint64_t*>(d)[k] : 0;
        cursor c(*this);
        c.begin(FLAG_M | FLAG_A);
        c.label_len(label, (uint32_t)n);
        c.begin(FLAG_A);
        if (words) {
#pragma unroll
            for (unsigned k = 0; k < 8; k++)
                if (k < nw) c.put(m[k], 8);
        } else {
            while (n) {
                const unsigned t = n < 8 ? (unsigned)n : 8u;
                c.put(load_le(d, t), t);
                d += t;
                n -= t;
            }
        }
        c.finish(FLAG_A);
    }
    QQ_HOSTDEV QQ_NOINLINE void begin_op(uint8_t flags, bool more) {
        if (more) return;   // continuation of the current operation (same flags by construction)
        uint8_t old_begin = pos_begin;
        pos_begin = (uint8_t)(pos + 1);
        cur_flags = flags;
        uint8_t hdr[2] = {old_begin, flags};
        absorb(hdr, 2);
        if ((flags & (FLAG_C | FLAG_K)) && pos != 0) run_f();
    }

  public:
    QQ_HOSTDEV explicit strobe128(const char* protocol_label) : pos(0), pos_begin(0), cur_flags(0) {
        memset(st, 0, sizeof st);
        const uint8_t init[6] = {1, R + 2, 1, 0, 1, 96};
        memcpy(st, init, 6);
        memcpy(st + 6, "STROBEv1.0.2", 12);
        permute();
        meta_ad((const uint8_t*)protocol_label, label_len(protocol_label), false);
    }
    QQ_HOSTDEV static size_t label_len(const char* s) {
        size_t n = 0;
        while (s[n]) n++;
        return n;
    }
    // serialised form (qq_transcript_state_bytes): 200 state bytes | pos | pos_begin | cur_flags | tag | 4 zero bytes
    static const int STATE_BYTES = 208;
    static const uint8_t STATE_TAG = 0xa5;
    QQ_HOSTDEV void export_state(uint8_t* out) const {
        memcpy(out, st, 200);
        out[200] = pos; out[201] = pos_begin; out[202] = cur_flags; out[203] = STATE_TAG;
        out[204] = out[205] = out[206] = out[207] = 0;
    }
    // false (and the object is left untouched) when the bytes are not a state export_state wrote: wrong tag (e.g. the all-zero
    // entry of a proof whose sigma check ended before the capture) or positions outside the sponge rate
    QQ_HOSTDEV bool import_state(const uint8_t* in) {
        if (in[203] != STATE_TAG || in[200] >= R || in[201] > R) return false;
        memcpy(st, in, 200);
        pos = in[200]; pos_begin = in[201]; cur_flags = in[202];
        return true;
    }
    QQ_HOSTDEV void meta_ad(const uint8_t* d, size_t n, bool more) {
        begin_op(FLAG_M | FLAG_A, more);
        absorb(d, n);
    }
    QQ_HOSTDEV void ad(const uint8_t* d, size_t n, bool more) {
        begin_op(FLAG_A, more);
        absorb(d, n);
    }
    QQ_HOSTDEV void prf(uint8_t* d, size_t n, bool more) {
        begin_op(FLAG_I | FLAG_A | FLAG_C, more);
        squeeze(d, n);
    }
};

// merlin::Transcript + the reference's TranscriptProtocol
class transcript {
    strobe128 s;

  public:
    static const int STATE_BYTES = strobe128::STATE_BYTES;
    QQ_HOSTDEV transcript(const uint8_t* label, size_t n) : s("Merlin v1.0") { append_message("dom-sep", label, n); }
    QQ_HOSTDEV void export_state(uint8_t* out) const { s.export_state(out); }
    QQ_HOSTDEV bool import_state(const uint8_t* in) { return s.import_state(in); }
    QQ_HOSTDEV void append_message(const char* label, const uint8_t* msg, size_t n) {
        s.op_append(label, msg, n);
    }
    QQ_HOSTDEV void challenge_bytes(const char* label, uint8_t* out, size_t n) {
        s.op_label_len(strobe128::FLAG_M | strobe128::FLAG_A, label, (uint32_t)n);
        s.prf(out, n, false);
    }
    QQ_HOSTDEV void domain_sep(const char* label) { append_message("dom-sep", (const uint8_t*)label, strobe128::label_len(label)); }
    QQ_HOSTDEV void append_point_var(const char* label, const uint8_t point[32]) {
        append_message("ptvar", (const uint8_t*)label, strobe128::label_len(label));
        append_message("val", point, 32);
    }
    QQ_HOSTDEV void append_scalar_var(const char* label, const uint8_t scalar[32]) { append_message(label, scalar, 32); }
    QQ_HOSTDEV void append_account_var(const char* label, const uint8_t acc[128]) {
        append_message("acvar", (const uint8_t*)label, strobe128::label_len(label));
        append_message("gr", acc, 32);
        append_message("grsk", acc + 32, 32);
        append_message("commc", acc + 64, 32);
        append_message("commd", acc + 96, 32);
    }
    // get_challenge: 64 challenge bytes reduced mod l (Scalar::from_bytes_mod_order_wide), canonical 32 bytes out
    QQ_HOSTDEV void get_challenge(const char* label, uint8_t out[32]);
};

// ---- scalars mod l on the host (only what the verifiers need: wide reduction, negation) --------------------------------
static const uint64_t L_WORDS[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0ULL, 0x1000000000000000ULL};

// r = x mod l for a 512-bit little-endian x: binary long division (512 shift-compare-subtract steps; a few hundred ns)
QQ_HOSTDEV static inline void sc_reduce_wide(uint8_t out[32], const uint8_t in[64]) {
    const uint64_t L_WORDS[4] = QQ_SC_L_WORDS;      // function-local: visible to device code too
    uint64_t r[5] = {0, 0, 0, 0, 0};
    for (int bit = 511; bit >= 0; bit--) {
        // r = 2 r + bit
        for (int i = 4; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
        r[0] = (r[0] << 1) | ((in[bit >> 3] >> (bit & 7)) & 1);
        // if r >= l: r -= l      (r < 2 l < 2^254 always)
        bool ge = r[4] != 0;
        if (!ge) {
            ge = true;
            for (int i = 3; i >= 0; i--) {
                if (r[i] != L_WORDS[i]) {
                    ge = r[i] > L_WORDS[i];
                    break;
                }
            }
        }
        if (ge) {
            unsigned __int128 borrow = 0;
            for (int i = 0; i < 4; i++) {
                unsigned __int128 d = (unsigned __int128)r[i] - L_WORDS[i] - (uint64_t)borrow;
                r[i] = (uint64_t)d;
                borrow = (d >> 64) & 1;
            }
            r[4] -= (uint64_t)borrow;
        }
    }
    memcpy(out, r, 32);
}
// out = -s mod l for canonical s
QQ_HOSTDEV static inline void sc_negate(uint8_t out[32], const uint8_t s[32]) {
    const uint64_t L_WORDS[4] = QQ_SC_L_WORDS;
    uint64_t a[4];
    memcpy(a, s, 32);
    if ((a[0] | a[1] | a[2] | a[3]) == 0) {
        memset(out, 0, 32);
        return;
    }
    unsigned __int128 borrow = 0;
    uint64_t r[4];
    for (int i = 0; i < 4; i++) {
        unsigned __int128 d = (unsigned __int128)L_WORDS[i] - a[i] - (uint64_t)borrow;
        r[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    memcpy(out, r, 32);
}
QQ_HOSTDEV inline void transcript::get_challenge(const char* label, uint8_t out[32]) {
    uint8_t wide[64];
    challenge_bytes(label, wide, 64);
    qq_sc::to_bytes(out, qq_sc::from_wide(wide));      // special-form reduction (sc_host.hpp); sc_reduce_wide is the slow cross-check
}

}  // namespace qq_merlin
