"""ctypes wrapper of oracle/_build/libqq_oracle.so (the plain-C restatement in oracle/qq_oracle.c).

CPU ORACLE -- test infrastructure.  Import only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libqq_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "qq_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_LIB), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-shared", "-fPIC", "-o", _LIB, src])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        if _lib.oq_init() != 0:
            raise RuntimeError("oq_init failed")
    return _lib


def threads():
    return lib().oq_threads()


def set_threads(n):
    lib().oq_set_threads(int(n))


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def _sz(n):
    return ctypes.c_size_t(n)


def update_account(acc, bl, u, c):
    acc, bl, u, c = _u8(acc), _u8(bl), _u8(u), _u8(c)
    n = bl.size // 32
    out, st = np.zeros(n * 128, np.uint8), np.zeros(n, np.uint8)
    lib().oq_update_account_batch(_p(acc), _p(bl), _p(u), _p(c), _p(out), _p(st), _sz(n))
    return out.reshape(n, 128), st


def verify_account(acc, sk, bl):
    acc, sk, bl = _u8(acc), _u8(sk), _u8(bl)
    n = sk.size // 32
    st = np.zeros(n, np.uint8)
    lib().oq_verify_account_batch(_p(acc), _p(sk), _p(bl), _p(st), _sz(n))
    return st


def update_public_key(pk, r):
    pk, r = _u8(pk), _u8(r)
    n = r.size // 32
    out, st = np.zeros(n * 64, np.uint8), np.zeros(n, np.uint8)
    lib().oq_update_public_key_batch(_p(pk), _p(r), _p(out), _p(st), _sz(n))
    return out.reshape(n, 64), st


def verify_public_key_update(upd, pk, r):
    upd, pk, r = _u8(upd), _u8(pk), _u8(r)
    n = r.size // 32
    st = np.zeros(n, np.uint8)
    lib().oq_verify_public_key_update_batch(_p(upd), _p(pk), _p(r), _p(st), _sz(n))
    return st


def generate_commitment(pk, r, v):
    pk, r, v = _u8(pk), _u8(r), _u8(v)
    n = r.size // 32
    out, st = np.zeros(n * 64, np.uint8), np.zeros(n, np.uint8)
    lib().oq_generate_commitment_batch(_p(pk), _p(r), _p(v), _p(out), _p(st), _sz(n))
    return out.reshape(n, 64), st


def add_commitments(a, b, negate_b=False):
    a, b = _u8(a), _u8(b)
    n = a.size // 64
    out, st = np.zeros(n * 64, np.uint8), np.zeros(n, np.uint8)
    lib().oq_add_commitments_batch(_p(a), _p(b), int(bool(negate_b)), _p(out), _p(st), _sz(n))
    return out.reshape(n, 64), st


def delta_epsilon(acc, bl, r, base_pk):
    acc, bl, r, base_pk = _u8(acc), _u8(bl), _u8(r), _u8(base_pk)
    n = bl.size // 32
    d, e, st = np.zeros(n * 128, np.uint8), np.zeros(n * 128, np.uint8), np.zeros(n, np.uint8)
    lib().oq_delta_epsilon_batch(_p(acc), _p(bl), _p(r), _p(base_pk), _p(d), _p(e), _p(st), _sz(n))
    return d.reshape(n, 128), e.reshape(n, 128), st


def fixed_base(which, s):
    s = _u8(s)
    n = s.size // 32
    out, st = np.zeros(n * 32, np.uint8), np.zeros(n, np.uint8)
    lib().oq_fixed_base_batch(int(which), _p(s), _p(out), _p(st), _sz(n))
    return out.reshape(n, 32), st


def msm(scalars, points):
    scalars, points = _u8(scalars), _u8(points)
    n = scalars.size // 32
    out, st = np.zeros(32, np.uint8), np.zeros(1, np.uint8)
    lib().oq_msm(_p(scalars), _p(points), _sz(n), _p(out), _p(st))
    return out, int(st[0])


def msm_segmented(scalars, points, offsets):
    scalars, points = _u8(scalars), _u8(points)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    m = offsets.size - 1
    out, st = np.zeros(m * 32, np.uint8), np.zeros(m, np.uint8)
    lib().oq_msm_segmented(_p(scalars), _p(points), _p(offsets), _sz(m), _p(out), _p(st))
    return out.reshape(m, 32), st


def delta_identity_check(acc):
    acc = _u8(acc)
    v = np.zeros(1, np.uint8)
    lib().oq_delta_identity_check(_p(acc), _sz(acc.size // 128), _p(v))
    return int(v[0])
