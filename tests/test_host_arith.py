"""CPU tests (no GPU): the exact device arithmetic (quisquis-rust_b200/csrc/*.cuh compiled for the host by
tests/csrc/host_harness.cpp -- test infrastructure only) against the big-int oracle, including limb bounds."""
import ctypes
import os
import random
import subprocess

import pytest

import ristretto_ref as R
from qq_testlib import invalid_encodings

HERE = os.path.dirname(os.path.abspath(__file__))
OFFS = [0, 26, 51, 77, 102, 128, 153, 179, 204, 230]


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


def val(l):
    return sum(int(x) << o for x, o in zip(l, OFFS)) % R.P


A10 = ctypes.c_uint32 * 10


def test_field_constants(hh):
    consts = [R.D, 2 * R.D % R.P, R.SQRT_M1, R.INVSQRT_A_MINUS_D, R.SQRT_AD_MINUS_ONE, R.ONE_MINUS_D_SQ, R.D_MINUS_ONE_SQ]
    for i, v in enumerate(consts):
        o = ctypes.create_string_buffer(32)
        hh.hh_fe_const(i, o)
        assert int.from_bytes(o.raw, "little") == v


def test_field_mul_sq_at_the_limb_bounds(hh):
    rnd = random.Random(11)

    def rl(sc):
        return [rnd.randrange(0, int(sc * (1 << (26 if i % 2 == 0 else 25)))) for i in range(10)]

    def mx(sc):
        return [int(sc * (1 << (26 if i % 2 == 0 else 25))) - 1 for i in range(10)]
    for it in range(4000):
        sf, sg = rnd.choice([1.0, 2.0, 3.0, 4.0, 5.0, 9.0]), rnd.choice([1.0, 2.0, 3.0, 3.3])
        if sf * sg > 30:
            sf, sg = 9.0, 3.3
        f = mx(sf) if it % 5 == 0 else rl(sf)
        g = mx(sg) if it % 10 == 0 else rl(sg)
        o = A10()
        hh.hh_fe_mul_limbs(o, A10(*f), A10(*g))
        assert val(o) == val(f) * val(g) % R.P
        for i in range(10):
            assert o[i] < ((1 << 26) if i % 2 == 0 else (1 << 25) + (1 << 18))
        ff = mx(3.3) if it % 7 == 0 else rl(min(sf, 3.3))
        hh.hh_fe_sq_limbs(o, A10(*ff))
        assert val(o) == val(ff) ** 2 % R.P
        b = ctypes.create_string_buffer(32)
        hh.hh_fe_tobytes_limbs(b, A10(*f))
        assert int.from_bytes(b.raw, "little") == val(f)


def test_canonical_encoding_edges(hh):
    for x in [0, 1, R.P - 1, R.P, R.P + 1, 2**255 - 1, 2 * R.P - 1, 2 * R.P, 2 * R.P + 5, 19, 2**255 - 20]:
        l = [(x >> o) & ((1 << (26 if i % 2 == 0 else 25)) - 1) for i, o in enumerate(OFFS)]
        l[9] += (x >> 255) << 25
        b = ctypes.create_string_buffer(32)
        hh.hh_fe_tobytes_limbs(b, A10(*l))
        assert int.from_bytes(b.raw, "little") == x % R.P


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
