"""CPU tests (no GPU): the exact device arithmetic (quisquis-rust_b200/csrc/*.cuh compiled for the host by
tests/csrc/host_harness.cpp -- test infrastructure only) against the big-int oracle, over the full 256-bit range."""
import ctypes
import os
import random
import subprocess

import pytest

import ristretto_ref as R
from qq_testlib import invalid_encodings

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hh():
    so = os.path.join(HERE, "csrc", "libhost_harness.so")
    src = os.path.join(HERE, "csrc", "host_harness.cpp")
    csrc = os.path.join(os.path.dirname(HERE), "quisquis-rust_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc) if f.endswith((".cuh", ".inc")))
    if not os.path.exists(so) or os.path.getmtime(so) < max(newest, os.path.getmtime(src)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-shared", "-fPIC", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.hh_fb_table_words.restype = ctypes.c_size_t
    return lib


A8 = ctypes.c_uint32 * 8


def limbs(x):
    return A8(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def val(l):
    return sum(int(l[i]) << (32 * i) for i in range(8))


EDGES = [0, 1, 2, 18, 19, 20, 37, 38, 39, R.P - 1, R.P, R.P + 1, 2 * R.P - 1, 2 * R.P, 2 * R.P + 1, 2**256 - 1,
         2**256 - 2, 2**256 - 38, 2**256 - 39, 2**255, 2**255 - 1, 2**255 - 19, 2**255 + 18, 2**128, 2**128 - 1,
         2**128 + 1, (2**128 - 1) << 128, 2**256 - 2**128, 2**224, 2**32 - 1, (2**32 - 1) << 224]


def rand_fe(rnd):
    """Field elements in the saturated form: ANY value in [0, 2^256), biased towards the carry/borrow corner cases."""
    r = rnd.random()
    if r < 0.25:
        return rnd.choice(EDGES)
    if r < 0.35:
        return (2**256 - 1) ^ rnd.getrandbits(rnd.choice([6, 20, 70]))
    if r < 0.45:
        return rnd.getrandbits(rnd.choice([6, 40, 130]))
    if r < 0.6:  # equal / nearly equal 128-bit halves: the |f0 - f1| path of the Karatsuba multiplication
        h = rnd.getrandbits(128)
        return h | ((h ^ rnd.getrandbits(rnd.choice([0, 1, 3, 33]))) << 128)
    if r < 0.7:  # limbs of all-ones / all-zeros
        return sum((0xFFFFFFFF if rnd.random() < 0.5 else 0) << (32 * i) for i in range(8))
    return rnd.getrandbits(256)


def test_field_constants(hh):
    consts = [R.D, 2 * R.D % R.P, R.SQRT_M1, R.INVSQRT_A_MINUS_D, R.SQRT_AD_MINUS_ONE, R.ONE_MINUS_D_SQ, R.D_MINUS_ONE_SQ]
    for i, v in enumerate(consts):
        o = ctypes.create_string_buffer(32)
        hh.hh_fe_const(i, o)
        assert int.from_bytes(o.raw, "little") == v


def test_field_ops_over_the_full_saturated_range(hh):
    """mul (Karatsuba and schoolbook), sq, add, sub, neg and canonical encoding for arbitrary 256-bit inputs."""
    rnd = random.Random(11)
    for it in range(30000):
        f, g = rand_fe(rnd), rand_fe(rnd)
        o, o2, o3 = A8(), A8(), A8()
        hh.hh_fe_mul_limbs(o, limbs(f), limbs(g))
        assert val(o) % R.P == f * g % R.P, (hex(f), hex(g))
        hh.hh_fe_mul_school_limbs(o, limbs(f), limbs(g))
        assert val(o) % R.P == f * g % R.P, (hex(f), hex(g))
        hh.hh_fe_sq_limbs(o, limbs(f))
        assert val(o) % R.P == f * f % R.P, hex(f)
        hh.hh_fe_addsub_limbs(o, o2, o3, limbs(f), limbs(g))
        assert val(o) % R.P == (f + g) % R.P and val(o2) % R.P == (f - g) % R.P and val(o3) % R.P == -f % R.P
        b = ctypes.create_string_buffer(32)
        hh.hh_fe_tobytes_limbs(b, limbs(f))
        assert int.from_bytes(b.raw, "little") == f % R.P


def test_canonical_encoding_edges(hh):
    for x in EDGES:
        b = ctypes.create_string_buffer(32)
        hh.hh_fe_tobytes_limbs(b, limbs(x))
        assert int.from_bytes(b.raw, "little") == x % R.P
        o = A8()
        hh.hh_fe_frombytes(o, x.to_bytes(32, "little"))
        assert val(o) == x & (2**255 - 1)      # bit 255 ignored, like dalek's FieldElement::from_bytes


def test_sqrt_ratio_and_invert(hh):
    rnd = random.Random(12)
    cases = [(0, 1), (1, 0), (0, 0), (5, R.P - 1), (R.P - 1, 1)] + [(rnd.randrange(R.P), rnd.randrange(R.P)) for _ in range(150)]
    for u, v in cases:
        o = ctypes.create_string_buffer(32)
        ok = hh.hh_sqrt_ratio_i(o, u.to_bytes(32, "little"), v.to_bytes(32, "little"))
        assert (bool(ok), int.from_bytes(o.raw, "little")) == R.sqrt_ratio_i(u, v)
        hh.hh_fe_invert(o, v.to_bytes(32, "little"))
        assert int.from_bytes(o.raw, "little") == pow(v, R.P - 2, R.P)


def test_decompress_compress_and_reject_rules(hh):
    rnd = random.Random(13)

    def check(b):
        o = ctypes.create_string_buffer(128)
        ok = hh.hh_decompress(o, b)
        e = R.decompress(b)
        assert bool(ok) == (e is not None), b.hex()
        if e is not None:
            assert [int.from_bytes(o.raw[i * 32:(i + 1) * 32], "little") for i in range(4)] == list(e)
            o2 = ctypes.create_string_buffer(32)
            hh.hh_compress_xyzt(o2, o.raw)
            assert o2.raw == b
    for _ in range(150):
        check(R.compress(R.mul(rnd.randrange(R.L), R.BASEPOINT)))
    check(bytes(32))
    for name, enc in invalid_encodings():
        check(enc)
    for i in range(1500):
        b = bytearray(rnd.randbytes(32))
        if i % 2:
            b[31] &= 0x7f
            b[0] &= 0xfe
        check(bytes(b))
    # projective (Z != 1) inputs compress to the same bytes
    for _ in range(60):
        p = R.mul(rnd.randrange(R.L), R.BASEPOINT)
        z = rnd.randrange(1, R.P)
        xb = b"".join((c * z % R.P).to_bytes(32, "little") for c in p)
        o = ctypes.create_string_buffer(32)
        hh.hh_compress_xyzt(o, xb)
        assert o.raw == R.compress(p)


def test_group_law_and_scalar_mult(hh):
    rnd = random.Random(14)

    def rp():
        return R.compress(R.mul(rnd.randrange(R.L), R.BASEPOINT))
    for i in range(80):
        a, b = rp(), rp()
        if i == 0:
            b = a
        if i == 1:
            b = R.compress(R.neg(R.decompress(a)))
        if i == 2:
            b = bytes(32)
        o = ctypes.create_string_buffer(32)
        hh.hh_add(o, a, b)
        assert o.raw == R.compress(R.add(R.decompress(a), R.decompress(b)))
        hh.hh_sub_madd(o, a, b)
        assert o.raw == R.compress(R.sub(R.decompress(a), R.decompress(b)))
        n = rnd.randrange(1, 9)
        hh.hh_dbl(o, a, n)
        assert o.raw == R.compress(R.mul(2**n, R.decompress(a)))
        r = hh.hh_eq(a, b)
        assert (r & 1) == int(R.eq(R.decompress(a), R.decompress(b))) and (r >> 1) == int(a == bytes(32))
    for i in range(120):
        s = [0, 1, 2, R.L - 1, 8, R.L - 8][i] if i < 6 else rnd.randrange(R.L)
        p = bytes(32) if i == 6 else rp()
        o = ctypes.create_string_buffer(32)
        assert hh.hh_scalarmult(o, s.to_bytes(32, "little"), p) == 3
        assert o.raw == R.compress(R.mul(s, R.decompress(p)))


def test_split_variable_base(hh):
    # vbs_*: four 16-digit quarters over P, 2^64 P, 2^128 P, 2^192 P -- same results as the plain window method
    rnd = random.Random(16)
    edge = [0, 1, 2, R.L - 1, 8, R.L - 8, 2**64, 2**64 - 1, 2**128 + 2**64, 2**192 - 1, 2**252, 8 * (16**64 - 1) // 15 % R.L]
    for i in range(40):
        s0 = edge[i] if i < len(edge) else rnd.randrange(R.L)
        s1 = edge[-1 - i] if i < len(edge) else rnd.randrange(R.L)
        p = bytes(32) if i == 3 else R.compress(R.mul(rnd.randrange(R.L), R.BASEPOINT))
        o0, o1 = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        assert hh.hh_scalarmult_split(o0, o1, s0.to_bytes(32, "little"), s1.to_bytes(32, "little"), p) == 1
        assert o0.raw == R.compress(R.mul(s0, R.decompress(p)))
        assert o1.raw == R.compress(R.mul(s1, R.decompress(p)))


def test_halve_and_double_compress(hh):
    # enc(s P) == double-and-compress((s / 2 mod l) P): dalek ristretto.rs double_and_compress_batch restated
    rnd = random.Random(17)
    inv2 = pow(2, R.L - 2, R.L)
    for i in range(60):
        s = [0, 1, 2, R.L - 1, R.L - 2, 3][i] if i < 6 else rnd.randrange(R.L)
        o = ctypes.create_string_buffer(32)
        hh.hh_sc_halve(o, s.to_bytes(32, "little"))
        assert int.from_bytes(o.raw, "little") == s * inv2 % R.L
        p = bytes(32) if i == 7 else R.compress(R.mul(rnd.randrange(R.L), R.BASEPOINT))
        z = (1 if i % 3 == 0 else rnd.randrange(1, R.P)).to_bytes(32, "little")
        zero = hh.hh_halve_dblcompress(o, s.to_bytes(32, "little"), p, z)
        want = R.compress(R.mul(s, R.decompress(p)))
        assert o.raw == want
        assert zero == int(want == bytes(32))


def test_from_uniform_bytes_elligator(hh):
    # RistrettoPoint::from_uniform_bytes restated on the device limbs; pinned by the reference's own constant:
    # BASE_PK_BTC_COMPRESSED[1] = from_uniform_bytes(SHA3-512(enc(B)))  (src/ristretto/constants.rs:17-20)
    import hashlib
    o = ctypes.create_string_buffer(32)
    hh.hh_from_uniform(o, hashlib.sha3_512(R.BASEPOINT_COMPRESSED).digest())
    assert o.raw == R.PEDERSEN_H_COMPRESSED
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "rfc9496.json")))
    for v in gold["hash_to_group_sha512"]:      # RFC 9496 Appendix A.3
        hh.hh_from_uniform(o, hashlib.sha512(v["label"].encode()).digest())
        assert o.raw.hex() == v["encoding"]
    rnd = random.Random(18)
    for i in range(60):
        b = bytes(64) if i == 0 else (b"\xff" * 64 if i == 1 else rnd.randbytes(64))
        hh.hh_from_uniform(o, b)
        assert o.raw == R.compress(R.from_uniform_bytes(b)), i


@pytest.mark.parametrize("W", [4, 6])
def test_fixed_base_tables(hh, W):
    rnd = random.Random(15)
    n = hh.hh_fb_table_words(W)
    tbl = (ctypes.c_uint32 * n)()
    for base, pt in ((R.BASEPOINT_COMPRESSED, R.BASEPOINT), (R.PEDERSEN_H_COMPRESSED, R.PEDERSEN_H)):
        hh.hh_fb_build(tbl, W, base)
        for i in range(30):
            s = [0, 1, 2, R.L - 1, 8, R.L - 8][i] if i < 6 else rnd.randrange(R.L)
            o = ctypes.create_string_buffer(32)
            hh.hh_fb_mult(o, tbl, W, s.to_bytes(32, "little"))
            assert o.raw == R.compress(R.mul(s, pt))


def test_host_scalar_field(hh):
    """quisquis-rust_b200/csrc/sc_host.hpp (Barrett arithmetic mod l used by the verifier drivers) against Python integers:
    edge values, random values, wide reduction, inversion, non-canonical inputs refused."""
    import random
    L = R.L
    rnd = random.Random(11)
    edge = [0, 1, 2, L - 1, L - 2, (L - 1) // 2, 2**252, 2**252 - 1, 2**128, 2**64 - 1, 2**64, 2**192 + 5]
    vals = edge + [rnd.randrange(L) for _ in range(200)]
    out = (ctypes.c_uint8 * 32)()

    def op(k, a, b):
        ab = (ctypes.c_uint8 * 32).from_buffer_copy(a.to_bytes(32, "little"))
        bb = (ctypes.c_uint8 * 32).from_buffer_copy(b.to_bytes(32, "little"))
        ok = hh.hh_sc_op(out, k, ab, bb)
        return int.from_bytes(bytes(out), "little") if ok else None
    for i, a in enumerate(vals):
        b = vals[(i * 7 + 3) % len(vals)]
        assert op(0, a, b) == (a + b) % L
        assert op(1, a, b) == (a - b) % L
        assert op(2, a, b) == (a * b) % L
        assert op(4, a, b) == (a + (b << 256)) % L
        assert op(6, a, b) == (a * b) % L                  # the device code's 32-bit limb product (mul_w32)
        assert op(7, a, b) == (a + (b << 256)) % L         # and its wide reduction (reduce512_w32)
    for a in edge[1:] + vals[-60:]:
        inv = op(3, a, 0)
        assert inv * a % L == 1
        assert op(5, a, 0) == inv              # binary extended Euclid (invert_vartime) == a^(l - 2)
        assert op(8, a, 0) == inv              # the branch-free fixed-round form the GPU runs (invert_fixed)
    assert op(5, 0, 0) == 0 and op(8, 0, 0) == 0
    for a in [2**k for k in range(1, 252, 7)] + [L - 2**k for k in range(1, 252, 7)] + [rnd.randrange(1, L) for _ in range(300)]:
        assert op(8, a, 0) * a % L == 1
    wides = [2**512 - 1, (L << 256) + L - 1, (2**256 - 1) << 256, 2**511, (L - 1) ** 2, L * L, L * L - 1, 2**252, 2**252 - 1, L, L - 1,
             2 * L, (2**252) * (2**260 - 1), 2**504 + 2**252 - 1, (1 << 512) - (1 << 252)]
    wides += [rnd.getrandbits(512) for _ in range(400)] + [rnd.getrandbits(512) | ((2**260 - 1) << 252) for _ in range(50)]
    wides += [rnd.getrandbits(rnd.choice([10, 100, 250, 253, 300, 385, 400])) for _ in range(200)]
    for wide in wides:
        assert op(4, wide & (2**256 - 1), wide >> 256) == wide % L
        assert op(7, wide & (2**256 - 1), wide >> 256) == wide % L
    assert op(2, L, 1) is None and op(0, 1, 2**256 - 1) is None


def test_host_keccak_and_merlin(hh):
    """The host side of the batched verifiers (keccak_host.hpp, merlin_host.hpp compiled with g++): SHA3-512 / SHAKE256 equal
    hashlib on messages around the rate boundaries; Merlin reproduces the merlin crate's conformance vector and equals the
    oracle's restatement on random scripts of appends and challenges (absorbs crossing the STROBE rate, long challenges)."""
    import ctypes
    import hashlib
    import random
    from merlin_ref import Transcript
    h = hh
    h.hh_sha3_512.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    h.hh_shake256.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    h.hh_merlin_script.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    rnd = random.Random(5)
    for n in (0, 1, 31, 32, 71, 72, 73, 135, 136, 137, 143, 144, 200, 1000):
        msg = bytes(rnd.randrange(256) for _ in range(n))
        o = ctypes.create_string_buffer(64)
        h.hh_sha3_512(o, msg, n)
        assert o.raw == hashlib.sha3_512(msg).digest(), n
        o = ctypes.create_string_buffer(300)
        h.hh_shake256(o, 300, msg, n)
        assert o.raw == hashlib.shake_256(msg).digest(300), n

    def run(label, ops):
        script = b""
        for kind, lab, data in ops:
            script += bytes([kind, len(lab)]) + lab
            script += (len(data) if kind == 0 else data).to_bytes(2, "little") + (data if kind == 0 else b"")
        o = ctypes.create_string_buffer(max([1] + [d for k, _, d in ops if k == 1]))
        h.hh_merlin_script(o, label, len(label), script, len(script))
        return o.raw
    # merlin crate, transcript.rs test "equivalence_simple"
    got = run(b"test protocol", [(0, b"some label", b"some data"), (1, b"challenge", 32)])
    assert got.hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    for trial in range(40):
        label = bytes(rnd.randrange(97, 123) for _ in range(rnd.randrange(1, 20)))
        ops, tr, last = [], Transcript(label), b""
        for _ in range(rnd.randrange(1, 30)):
            lab = bytes(rnd.randrange(97, 123) for _ in range(rnd.randrange(1, 16)))
            if rnd.random() < 0.75:
                data = bytes(rnd.randrange(256) for _ in range(rnd.choice((0, 1, 8, 32, 32, 32, 64, 128, 165, 166, 167, 400))))
                ops.append((0, lab, data))
                tr.append_message(lab, data)
            else:
                k = rnd.choice((1, 32, 64, 64, 128, 200))
                ops.append((1, lab, k))
                last = tr.challenge_bytes(lab, k)
        k = 64
        ops.append((1, b"final", k))
        last = tr.challenge_bytes(b"final", k)
        got = run(label, ops)
        assert got[:k] == last, trial


def test_shuffle_phases_shared_with_the_transcript_kernels(hh):
    """The per-proof phases of the shuffle verifier (quisquis-rust_b200/csrc/shuffle_verify.cuh: pass A, pass B, final verdict
    - the code k_shuffle_pass_a / _pass_b / _final run one GPU thread per proof) compiled for the host, MSM batches evaluated
    with the host-compiled device arithmetic: the golden proofs are accepted, tampered ones rejected at the stage the oracle's
    restatement of ShuffleProof::verify (src/shuffle/shuffle.rs:547-712) names.  Also the serialised transcript state."""
    import numpy as np
    import shuffle_ref as F
    from qq_testlib import Stream, cat
    from test_gpu_parity import _shuffle_blobs, _shuffle_code, shuffle_cases
    h = hh
    u8p = ctypes.c_char_p
    h.hh_shuffle_verify.argtypes = [u8p, u8p, u8p, u8p, u8p, u8p, ctypes.c_size_t, u8p, u8p, u8p, u8p, u8p]
    h.hh_transcript_state_bytes.restype = ctypes.c_size_t
    xpc = F.XpcGens(4)
    xpc_bytes = xpc.h + b"".join(xpc.g[:3])        # compressed H | G[0..3)
    base_pk = R.BASEPOINT_COMPRESSED + R.PEDERSEN_H_COMPRESSED

    def run(si, so, stm, pr, n, label=b"ShuffleProof"):
        st, sg, dt = (ctypes.create_string_buffer(n) for _ in range(3))
        h.hh_shuffle_verify(label, b"Shuffle", si, so, stm, pr, n, base_pk, xpc_bytes, st, sg, dt)
        return list(st.raw), list(sg.raw), list(dt.raw)
    raw = np.fromfile(os.path.join(HERE, "golden", "shuffle_proofs.bin"), dtype=np.uint8).reshape(-1, 6432)
    n = raw.shape[0]
    cols = lambda rec: tuple(np.ascontiguousarray(rec[:, a:b]).tobytes() for a, b in ((0, 1152), (1152, 2304), (2304, 2656), (2656, 6432)))  # noqa: E731
    st, sg, dt = run(*cols(raw), n)
    assert st == [0] * n and sg == [0] * n
    bad = raw.copy()
    bad[0, 1152 + 5] ^= 1                 # an output account
    bad[1, 2656 + 3776 - 1 - 32] ^= 1     # top byte of the DDH challenge
    bad[2, 2656 + 384 + 608] ^= 1         # Hadamard rho_bar
    st, sg, dt = run(*cols(bad), n)
    assert st[0] != 0 and st[1] != 0 and st[2] != 0 and st[3:] == [0] * (n - 3)
    assert sg[2] == 1 and sg[1] == 6
    # other label: the first challenge moves, the Hadamard argument rejects
    st, sg, dt = run(*cols(raw[:1]), 1, label=b"Other")
    assert (st[0], sg[0]) == (6, 1)
    # the 23 accept / reject cases of the GPU parity test, against the oracle's verifier
    cases = shuffle_cases(Stream(b"shuffle-gpu"))
    expect = [_shuffle_code(F.shuffle_verify(F.new_transcript(b"ShuffleProof", b"Shuffle"), k[2], k[3], k[0], k[1], xpc)) for k in cases]
    blobs = [_shuffle_blobs(k[2], k[3]) for k in cases]
    st, sg, dt = run(b"".join(b"".join(k[0]) for k in cases), b"".join(b"".join(k[1]) for k in cases),
                     b"".join(b[1] for b in blobs), b"".join(b[0] for b in blobs), len(cases))
    for i, e in enumerate(expect):
        assert (st[i], sg[i]) == e[:2] and (e[2] is None or dt[i] == e[2]), (i, e, (st[i], sg[i], dt[i]))
    # ---- aggregate form (fast path): clean proofs, one weighted sum over all proofs == identity
    h.hh_shuffle_verify_aggregate.argtypes = [u8p, u8p, u8p, u8p, u8p, u8p, ctypes.c_size_t, u8p, u8p, u8p, u8p, u8p,
                                              ctypes.POINTER(ctypes.c_uint32)]

    def run_agg(si, so, stm, pr, n, entropy=bytes(range(32))):
        clean, ident = ctypes.create_string_buffer(n), ctypes.create_string_buffer(1)
        counts = (ctypes.c_uint32 * (2 * n))()
        h.hh_shuffle_verify_aggregate(b"ShuffleProof", b"Shuffle", si, so, stm, pr, n, base_pk, xpc_bytes, entropy, clean, ident, counts)
        return list(clean.raw), ident.raw[0], list(counts)
    clean, ident, counts = run_agg(*cols(raw), n)
    assert clean == [1] * n and ident == 1
    assert counts == [64, 102] * n                       # QQ_SHUFFLE_AGG_CAP_1 / _2 are exact for a clean proof
    clean, ident, counts = run_agg(*cols(raw), n, entropy=bytes(32))
    assert clean == [1] * n and ident == 1               # other weights, same verdict
    clean, ident, counts = run_agg(*cols(bad), n)
    assert clean[1] == 0                                 # DDH challenge: scalar-level, caught before the aggregate
    assert clean[0] == 1 and clean[2] == 1 and ident == 0    # a flipped output account / rho_bar only show in the group equations
    only_scalar_bad = raw.copy()
    only_scalar_bad[1, 2656 + 3776 - 1 - 32] ^= 1
    clean, ident, counts = run_agg(*cols(only_scalar_bad), n)
    assert clean == [1, 0] + [1] * (n - 2) and ident == 1    # the dirty proof left the aggregate, the rest still verifies
    # every group-level tampering of the 23 cases breaks the aggregate when it is the only proof in it
    for i, k in enumerate(cases):
        cl, idn, _ = run_agg(b"".join(k[0]), b"".join(k[1]), blobs[i][1], blobs[i][0], 1)
        assert (cl[0] == 1 and idn == 1) == (expect[i][0] == 0), (i, expect[i], cl, idn)
    # serialised transcript state: round trip, corrupted tag / position refused
    state = ctypes.create_string_buffer(h.hh_transcript_state_bytes())
    assert h.hh_transcript_state_bytes() == 208
    assert h.hh_transcript_state_roundtrip(b"abc", 3, state) == 1
    assert state.raw[203] == 0xa5


# ---- FP64-pipe arithmetic (csrc/fe64.cuh): exact by construction; checked here against big integers ------------------
F64_OFF = [(85 * i + 3) // 4 for i in range(13)]
Q12 = ctypes.c_longlong * 12


def test_fe64_column_bound_interval_arithmetic():
    # the claim in the header of fe64.cuh: a column of products of operands with ta and tb carried terms stays below
    # 2^(o_k + 53) whenever ta * tb <= 16 (the group law uses at most 3 x 3 = 9, 12 with the doubled squaring)
    assert F64_OFF == [0, 22, 43, 64, 85, 107, 128, 149, 170, 192, 213, 234, 255]
    from fractions import Fraction
    N = [Fraction(2) ** (F64_OFF[i + 1] - 1) for i in range(12)]
    N[2] = N[2] * (1 + Fraction(1, 2**19))           # second-lap excess of the carry
    for k in range(12):
        tot = Fraction(0)
        for i in range(12):
            j = k - i
            tot += N[i] * N[j] if j >= 0 else N[i] * N[j + 12] * 19 / Fraction(2) ** 255
        assert 16 * tot < Fraction(2) ** (F64_OFF[k] + 53), k
        for i in range(12):                            # every product is a multiple of the column's unit
            j = k - i
            assert F64_OFF[i] + F64_OFF[j if j >= 0 else j + 12] - (0 if j >= 0 else 255) >= F64_OFF[k]


def test_fe64_products_and_conversions(hh):
    rnd = random.Random(64)
    out = A8()
    for x in EDGES + [rand_fe(rnd) for _ in range(300)]:
        for terms in (1, 2, 5, 12):
            hh.hh_fe64_roundtrip(out, limbs(x), terms)
            assert val(out) % R.P == terms * x % R.P and val(out) < 2**256

    def raw_val(q):
        return sum(int(q[i]) << F64_OFF[i] for i in range(12))

    def carried(scale):
        # signed limbs with |q_i| <= scale * 2^(s_i - 1), biased towards the extremes
        q = []
        for i in range(12):
            m = scale << (F64_OFF[i + 1] - F64_OFF[i] - 1)
            r = rnd.random()
            q.append(m if r < 0.2 else -m if r < 0.4 else rnd.randint(-m, m))
        return q
    for ta, tb in [(1, 1), (2, 2), (3, 3), (4, 4), (4, 3), (2, 8), (16, 1), (1, 16)]:
        for _ in range(60):
            qa, qb = carried(ta), carried(tb)
            if rnd.random() < 0.15:
                qa = [s * (ta << (F64_OFF[i + 1] - F64_OFF[i] - 1)) for i, s in enumerate([rnd.choice([1, -1])] * 12)]
                qb = [s * (tb << (F64_OFF[i + 1] - F64_OFF[i] - 1)) for i, s in enumerate([rnd.choice([1, -1])] * 12)]
            hh.hh_fe64_op_raw(out, Q12(*qa), Q12(*qb), 0)
            assert val(out) != 2**256 - 1 and val(out) % R.P == raw_val(qa) * raw_val(qb) % R.P, (ta, tb)
            if ta * ta <= 16:
                hh.hh_fe64_op_raw(out, Q12(*qa), Q12(*qb), 1)
                assert val(out) % R.P == raw_val(qa) ** 2 % R.P
            if 2 * ta * ta <= 16:
                hh.hh_fe64_op_raw(out, Q12(*qa), Q12(*qb), 2)
                assert val(out) % R.P == 2 * raw_val(qa) ** 2 % R.P


def test_fe64_split_variable_base(hh):
    # the FP64 warps' scalar multiplication (vbs64_*) gives the encodings of the integer path
    rnd = random.Random(65)
    edge = [0, 1, 2, R.L - 1, 8, R.L - 8, 2**64, 2**64 - 1, 2**128 + 2**64, 2**192 - 1, 2**252, 8 * (16**64 - 1) // 15 % R.L]
    for i in range(30):
        s0 = edge[i] if i < len(edge) else rnd.randrange(R.L)
        s1 = edge[-1 - i] if i < len(edge) else rnd.randrange(R.L)
        p = bytes(32) if i == 3 else R.compress(R.mul(rnd.randrange(R.L), R.BASEPOINT))
        o0, o1 = ctypes.create_string_buffer(32), ctypes.create_string_buffer(32)
        assert hh.hh_scalarmult_split64(o0, o1, s0.to_bytes(32, "little"), s1.to_bytes(32, "little"), p) == 1
        assert o0.raw == R.compress(R.mul(s0, R.decompress(p)))
        assert o1.raw == R.compress(R.mul(s1, R.decompress(p)))


def test_host_scalar_field_lazy_sums(hh):
    """The range-proof fold's lazy arithmetic (sc_host.hpp): products added up unreduced in 17 limbs and reduced once (including
    sums that spill into the 17th limb), and (a 2^k - x) mod l with one reduction, against Python integers."""
    import random
    L = R.L
    rnd = random.Random(12)
    edge = [0, 1, L - 1, L - 2, 2**252, 2**252 - 1, (L - 1) // 2]
    s32 = lambda v: v.to_bytes(32, "little")
    osum, ou = (ctypes.c_uint8 * 32)(), (ctypes.c_uint8 * 32)()
    for trial in range(60):
        n = rnd.choice([1, 2, 7, 40])
        a = [rnd.choice(edge) if rnd.random() < 0.4 else rnd.randrange(L) for _ in range(n)]
        b = [rnd.choice(edge) if rnd.random() < 0.4 else rnd.randrange(L) for _ in range(n)]
        repeat = rnd.choice([1, 1, 3, 200, 5000])          # (L - 1)^2 * 40 * 5000 > 2^512: the top limb is used
        rzz, sb, slo = (rnd.choice(edge) if rnd.random() < 0.4 else rnd.randrange(L) for _ in range(3))
        k = rnd.choice([0, 1, 31, 32, 33, 63, rnd.randrange(64)])
        hh.hh_sc_lazy(osum, ou, b"".join(map(s32, a)), b"".join(map(s32, b)), n, repeat, s32(rzz), k, s32(sb), s32(slo))
        assert int.from_bytes(bytes(osum), "little") == repeat * sum(x * y for x, y in zip(a, b)) % L
        assert int.from_bytes(bytes(ou), "little") == (rzz * 2**k - sb * slo) % L
