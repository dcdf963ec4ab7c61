"""ctypes binding of libqq_b200.so (include/qq_b200.h).  No CPU fallback: a missing library or GPU raises."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

QQ_OK = 0
ST_OK, ST_BAD_POINT, ST_BAD_SCALAR, ST_KEYPAIR, ST_COMMIT, ST_NOT_FOUND, ST_PROOF, ST_PANIC = 0, 1, 2, 3, 4, 5, 6, 7
BASE_B, BASE_H = 0, 1

EXPORTS = [
    "qq_init", "qq_destroy", "qq_last_error", "qq_device_sm_count", "qq_launch_count", "qq_last_kernel_ms",
    "qq_last_kernel_breakdown", "qq_dev_alloc", "qq_dev_free", "qq_dev_upload", "qq_dev_download",
    "qq_measure_imad_peak", "qq_event_record", "qq_event_elapsed_ms",
    "qq_update_public_key_batch", "qq_update_public_key_batch_dev", "qq_verify_public_key_update_batch",
    "qq_generate_commitment_batch", "qq_generate_commitment_batch_dev", "qq_add_commitments_batch",
    "qq_mul_commitment_batch", "qq_update_account_batch", "qq_update_account_batch_dev",
    "qq_verify_account_batch", "qq_verify_account_batch_dev", "qq_delta_epsilon_batch", "qq_delta_identity_check",
    "qq_fixed_base_batch", "qq_fixed_base_batch_dev", "qq_fixed_base_set_window", "qq_fixed_base_window", "qq_varbase_set_coop_limit",
    "qq_fixed_base_i64_batch", "qq_fixed_base_i64_batch_dev", "qq_msm", "qq_msm_dev", "qq_msm_partial", "qq_msm_partial_dev",
    "qq_points_sum", "qq_msm_segmented", "qq_msm_grouped", "qq_msm_points_prepare", "qq_msm_points_prepare_dev", "qq_msm_points_free",
    "qq_msm_points_count", "qq_msm_prepared", "qq_msm_prepared_dev",
    "qq_verify_ddh_batch", "qq_verify_svp_batch", "qq_verify_hadamard_batch", "qq_verify_product_batch", "qq_verify_shuffle_batch",
    "qq_verify_account_sigma_batch", "qq_verify_zero_balance_batch", "qq_verify_destroy_account_batch",
    "qq_verify_same_value_compact_batch", "qq_verify_update_account_dark_tx_batch",
    "qq_verify_update_account_dlog_batch", "qq_verify_delta_compact_batch", "qq_decommit_batch", "qq_decommit_value_batch", "qq_from_uniform_bytes_batch", "qq_vector_pedersen_gens", "qq_bulletproof_gens",
    "qq_verify_range_proof_batch", "qq_transcript_state_bytes", "qq_transcript_capture", "qq_msm_set_overlap",
    "qq_verify_set_transcripts", "qq_verify_set_aggregation",
    "qq_shuffle_proofs_from_bincode", "qq_shuffle_statements_from_bincode", "qq_shuffle_proofs_to_bincode",
    "qq_shuffle_statements_to_bincode", "qq_accounts_from_bincode", "qq_sigma_proof_from_bincode",
    "qq_msm_set_shifted", "qq_msm_points_shifted_bytes", "qq_set_secret_mode", "qq_secret_mode", "qq_sigma_commit_batch", "qq_sigma_commit_batch_dev",
    "qq_init_multi", "qq_destroy_multi", "qq_multi_device_count", "qq_multi_ctx", "qq_multi_last_error",
    "qq_multi_update_account_batch", "qq_multi_generate_commitment_batch", "qq_multi_verify_shuffle_batch",
    "qq_multi_verify_range_proof_batch", "qq_multi_msm", "qq_multi_msm_dev", "qq_points_sum_dev", "qq_warp_ops_selftest",
]


class QQError(RuntimeError):
    pass


def lib_path():
    return os.path.join(_HERE, "libqq_b200.so")


_lib = None


def load_library():
    """Load libqq_b200.so; raises QQError if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not os.path.exists(p):
        raise QQError("libqq_b200.so is not built at %s -- run __graft_entry__.build(); there is no CPU fallback" % p)
    lib = ctypes.CDLL(p)
    vp, sz, u8p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p
    lib.qq_init.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
    lib.qq_destroy.argtypes = [vp]
    lib.qq_destroy.restype = None
    lib.qq_last_error.argtypes = [vp]
    lib.qq_last_error.restype = ctypes.c_char_p
    lib.qq_device_sm_count.argtypes = [vp]
    lib.qq_launch_count.argtypes = [vp]
    lib.qq_launch_count.restype = ctypes.c_uint64
    lib.qq_last_kernel_ms.argtypes = [vp]
    lib.qq_last_kernel_ms.restype = ctypes.c_float
    lib.qq_last_kernel_breakdown.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.c_int]
    lib.qq_dev_alloc.argtypes = [vp, ctypes.POINTER(vp), sz]
    lib.qq_dev_free.argtypes = [vp, vp]
    lib.qq_dev_upload.argtypes = [vp, vp, vp, sz]
    lib.qq_dev_download.argtypes = [vp, vp, vp, sz]
    lib.qq_event_record.argtypes = [vp, ctypes.c_int]
    lib.qq_event_elapsed_ms.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    lib.qq_measure_imad_peak.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    for name in ("qq_update_public_key_batch", "qq_update_public_key_batch_dev", "qq_mul_commitment_batch"):
        getattr(lib, name).argtypes = [vp, u8p, u8p, u8p, u8p, sz]
    lib.qq_verify_public_key_update_batch.argtypes = [vp, u8p, u8p, u8p, u8p, sz]
    for name in ("qq_generate_commitment_batch", "qq_generate_commitment_batch_dev"):
        getattr(lib, name).argtypes = [vp, u8p, u8p, u8p, u8p, u8p, sz]
    lib.qq_add_commitments_batch.argtypes = [vp, u8p, u8p, ctypes.c_int, u8p, u8p, sz]
    for name in ("qq_update_account_batch", "qq_update_account_batch_dev"):
        getattr(lib, name).argtypes = [vp, u8p, u8p, u8p, u8p, u8p, u8p, sz]
    for name in ("qq_verify_account_batch", "qq_verify_account_batch_dev"):
        getattr(lib, name).argtypes = [vp, u8p, u8p, u8p, u8p, sz]
    lib.qq_delta_epsilon_batch.argtypes = [vp, u8p, u8p, u8p, u8p, u8p, u8p, u8p, sz]
    lib.qq_delta_identity_check.argtypes = [vp, u8p, sz, u8p]
    for name in ("qq_fixed_base_batch", "qq_fixed_base_batch_dev"):
        getattr(lib, name).argtypes = [vp, ctypes.c_int, u8p, u8p, u8p, sz]
    for name in ("qq_fixed_base_i64_batch", "qq_fixed_base_i64_batch_dev"):
        getattr(lib, name).argtypes = [vp, ctypes.c_int, vp, u8p, sz]
    lib.qq_fixed_base_set_window.argtypes = [vp, ctypes.c_int, ctypes.c_int]
    lib.qq_fixed_base_window.argtypes = [vp, ctypes.c_int]
    lib.qq_varbase_set_coop_limit.argtypes = [vp, ctypes.c_long]
    for name in ("qq_msm", "qq_msm_dev", "qq_msm_partial", "qq_msm_partial_dev"):
        getattr(lib, name).argtypes = [vp, u8p, u8p, sz, u8p, u8p]
    lib.qq_points_sum.argtypes = [vp, u8p, sz, u8p, u8p]
    lib.qq_warp_ops_selftest.argtypes = [vp, u8p, u8p, sz, u8p, u8p]
    for name in ("qq_msm_points_prepare", "qq_msm_points_prepare_dev"):
        getattr(lib, name).argtypes = [vp, u8p, sz, ctypes.POINTER(vp)]
    lib.qq_verify_update_account_dlog_batch.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, u8p, u8p, u8p, u8p, sz, sz, u8p]
    lib.qq_verify_delta_compact_batch.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, u8p, u8p, u8p, u8p, u8p, u8p, sz, sz, u8p]
    cs = ctypes.c_char_p
    lib.qq_verify_account_sigma_batch.argtypes = [vp, cs, cs, u8p, u8p, u8p, u8p, u8p, u8p, u8p, sz, sz, u8p]
    lib.qq_verify_zero_balance_batch.argtypes = [vp, cs, cs, u8p, u8p, u8p, sz, sz, ctypes.c_int, u8p]
    lib.qq_verify_destroy_account_batch.argtypes = [vp, cs, cs, u8p, u8p, u8p, sz, sz, u8p]
    lib.qq_verify_same_value_compact_batch.argtypes = [vp, u8p, u8p, u8p, u8p, u8p, sz, u8p]
    lib.qq_verify_update_account_dark_tx_batch.argtypes = [vp, cs, cs, u8p, u8p, u8p, u8p, sz, sz, u8p]
    lib.qq_verify_ddh_batch.argtypes = [vp, cs, cs, u8p, u8p, u8p, u8p, u8p, u8p, sz, u8p]
    lib.qq_verify_svp_batch.argtypes = [vp, cs, cs, u8p, u8p, u8p, sz, u8p]
    lib.qq_verify_hadamard_batch.argtypes = [vp, cs, cs, u8p, u8p, u8p, u8p, u8p, sz, u8p, u8p]
    lib.qq_verify_product_batch.argtypes = [vp, cs, cs, u8p, u8p, u8p, sz, u8p, u8p]
    lib.qq_verify_shuffle_batch.argtypes = [vp, cs, cs, u8p, u8p, u8p, u8p, sz, u8p, u8p, u8p]
    lib.qq_decommit_batch.argtypes = [vp, u8p, u8p, u8p, u8p, sz]
    lib.qq_decommit_value_batch.argtypes = [vp, u8p, u8p, ctypes.c_int, u8p, u8p, sz]
    lib.qq_from_uniform_bytes_batch.argtypes = [vp, u8p, u8p, sz]
    lib.qq_vector_pedersen_gens.argtypes = [vp, sz, u8p, u8p]
    lib.qq_bulletproof_gens.argtypes = [vp, sz, sz, u8p, u8p]
    lib.qq_verify_range_proof_batch.argtypes = [vp, cs, cs, u8p, cs, u8p, u8p, sz, sz, sz, sz, u8p]
    lib.qq_transcript_state_bytes.argtypes = []
    lib.qq_transcript_state_bytes.restype = ctypes.c_size_t
    lib.qq_transcript_capture.argtypes = [vp, u8p, ctypes.c_size_t]
    lib.qq_verify_set_transcripts.argtypes = [vp, ctypes.c_int]
    lib.qq_verify_set_aggregation.argtypes = [vp, ctypes.c_int]
    szp = ctypes.POINTER(ctypes.c_size_t)
    lib.qq_msm_set_shifted.argtypes = [vp, ctypes.c_size_t, ctypes.c_int]
    lib.qq_msm_points_shifted_bytes.argtypes = [vp]
    lib.qq_msm_points_shifted_bytes.restype = ctypes.c_size_t
    lib.qq_set_secret_mode.argtypes = [vp, ctypes.c_int]
    lib.qq_secret_mode.argtypes = [vp]
    for name in ("qq_sigma_commit_batch", "qq_sigma_commit_batch_dev"):
        getattr(lib, name).argtypes = [vp, u8p, u8p, u8p, u8p, u8p, sz]
    lib.qq_init_multi.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int), ctypes.c_int]
    lib.qq_destroy_multi.argtypes = [vp]
    lib.qq_destroy_multi.restype = None
    lib.qq_multi_device_count.argtypes = [vp]
    lib.qq_multi_ctx.argtypes = [vp, ctypes.c_int]
    lib.qq_multi_ctx.restype = vp
    lib.qq_multi_last_error.argtypes = [vp]
    lib.qq_multi_last_error.restype = ctypes.c_char_p
    lib.qq_multi_update_account_batch.argtypes = [vp, u8p, u8p, u8p, u8p, u8p, u8p, sz]
    lib.qq_multi_generate_commitment_batch.argtypes = [vp, u8p, u8p, u8p, u8p, u8p, sz]
    lib.qq_multi_verify_shuffle_batch.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, u8p, u8p, u8p, u8p, sz, u8p, u8p, u8p]
    lib.qq_multi_verify_range_proof_batch.argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p, u8p, ctypes.c_char_p, u8p, u8p, sz, sz, sz, sz, u8p]
    lib.qq_multi_msm.argtypes = [vp, u8p, u8p, sz, u8p, u8p]
    lib.qq_multi_msm_dev.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), szp, u8p, u8p]
    lib.qq_points_sum_dev.argtypes = [vp, u8p, sz, sz, u8p, u8p, u8p]
    lib.qq_shuffle_proofs_from_bincode.argtypes = [u8p, sz, sz, u8p, szp]
    lib.qq_shuffle_statements_from_bincode.argtypes = [u8p, sz, sz, u8p, szp]
    lib.qq_shuffle_proofs_to_bincode.argtypes = [u8p, sz, u8p, sz, szp]
    lib.qq_shuffle_statements_to_bincode.argtypes = [u8p, sz, u8p, sz, szp]
    lib.qq_accounts_from_bincode.argtypes = [u8p, sz, u8p, sz, szp, szp]
    lib.qq_sigma_proof_from_bincode.argtypes = [u8p, sz, ctypes.POINTER(ctypes.c_int), u8p, sz, szp, u8p, szp]
    lib.qq_msm_set_overlap.argtypes = [vp, ctypes.c_long, ctypes.c_int, ctypes.c_int]
    lib.qq_msm_points_free.argtypes = [vp, vp]
    lib.qq_msm_points_free.restype = None
    lib.qq_msm_points_count.argtypes = [vp]
    lib.qq_msm_points_count.restype = ctypes.c_size_t
    for name in ("qq_msm_prepared", "qq_msm_prepared_dev"):
        getattr(lib, name).argtypes = [vp, u8p, vp, sz, u8p, u8p]
    lib.qq_msm_segmented.argtypes = [vp, u8p, u8p, vp, sz, u8p, u8p]
    lib.qq_msm_grouped.argtypes = [vp, u8p, u8p, vp, sz, u8p, u8p]
    for name in EXPORTS:
        f = getattr(lib, name)
        if f.restype is ctypes.c_int and name not in ("qq_device_sm_count", "qq_last_kernel_breakdown"):
            pass
    _lib = lib
    return lib


def _u8(a, nbytes=None):
    if isinstance(a, (bytes, bytearray, memoryview)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
    if nbytes is not None and a.size != nbytes:
        raise ValueError("expected %d bytes, got %d" % (nbytes, a.size))
    return a


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


class MultiEngine:
    """qq_multi: several GPUs behind one handle (one context + worker thread per device, device 0 of the list is the root)."""

    def __init__(self, devices):
        self.lib = load_library()
        self.h = ctypes.c_void_p()
        arr = (ctypes.c_int * len(devices))(*devices)
        rc = self.lib.qq_init_multi(ctypes.byref(self.h), arr, len(devices))
        if rc != QQ_OK:
            raise QQError("qq_init_multi failed (%d); there is no CPU fallback" % rc)
        self.ndev = len(devices)

    def close(self):
        if self.h:
            self.lib.qq_destroy_multi(self.h)
            self.h = ctypes.c_void_p()

    def _ck(self, rc, what):
        if rc != QQ_OK:
            raise QQError("%s failed (%d): %s" % (what, rc, (self.lib.qq_multi_last_error(self.h) or b"").decode()))

    def ctx(self, index):
        return self.lib.qq_multi_ctx(self.h, index)

    def update_account(self, acc, bl, u, c):
        acc, bl, u, c = _u8(acc), _u8(bl), _u8(u), _u8(c)
        n = acc.size // 128
        out, st = np.zeros(n * 128, np.uint8), np.zeros(n, np.uint8)
        self._ck(self.lib.qq_multi_update_account_batch(self.h, _ptr(acc), _ptr(bl), _ptr(u), _ptr(c), _ptr(out), _ptr(st), n),
                 "qq_multi_update_account_batch")
        return out.reshape(n, 128), st

    def msm(self, scalars, points):
        s, p = _u8(scalars), _u8(points)
        n = s.size // 32
        out, st = np.zeros(32, np.uint8), np.zeros(1, np.uint8)
        self._ck(self.lib.qq_multi_msm(self.h, _ptr(s), _ptr(p), n, _ptr(out), _ptr(st)), "qq_multi_msm")
        return out, int(st[0])

    def msm_dev(self, scalar_ptrs, point_ptrs, counts):
        nd = self.ndev
        sp = (ctypes.c_void_p * nd)(*scalar_ptrs)
        pp = (ctypes.c_void_p * nd)(*point_ptrs)
        cn = (ctypes.c_size_t * nd)(*counts)
        out, st = np.zeros(32, np.uint8), np.zeros(1, np.uint8)
        self._ck(self.lib.qq_multi_msm_dev(self.h, sp, pp, cn, _ptr(out), _ptr(st)), "qq_multi_msm_dev")
        return out, int(st[0])

    def verify_shuffle(self, shuffle_input, shuffle_output, statement, proof, transcript_label=b"ShuffleProof", verifier_label=b"Shuffle"):
        si, so, stm, pr = (_u8(a) for a in (shuffle_input, shuffle_output, statement, proof))
        nproofs = pr.size // 3776
        st, sg, det = (np.zeros(nproofs, np.uint8) for _ in range(3))
        self._ck(self.lib.qq_multi_verify_shuffle_batch(self.h, transcript_label, verifier_label, _ptr(si), _ptr(so), _ptr(stm), _ptr(pr),
                                                        nproofs, _ptr(st), _ptr(sg), _ptr(det)), "qq_multi_verify_shuffle_batch")
        return st, sg, det

    def verify_range_proofs(self, commitments, proofs, m, chain=1, n_bits=64, transcript_label=b"SenderAccountProof",
                            verifier_label=b"BulletProof", domain_label=b"AggregateBulletProof"):
        cm, pr = _u8(commitments), _u8(proofs)
        per = (9 + 2 * ((n_bits * m).bit_length() - 1)) * 32 * chain
        nproofs = pr.size // per
        st = np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_multi_verify_range_proof_batch(self.h, transcript_label, verifier_label, None, domain_label, _ptr(cm), _ptr(pr),
                                                            n_bits, m, chain, nproofs, _ptr(st)), "qq_multi_verify_range_proof_batch")
        return st


# ---- wire format (host-only helpers of the library; no GPU context) ------------------------------------------------------
def shuffle_proofs_from_bincode(data, nproofs):
    """bincode(ShuffleProof) x nproofs, back to back -> (nproofs x 3776 flattened bytes, bytes consumed); ValueError when malformed."""
    lib = load_library()
    d = _u8(data)
    out = np.zeros(nproofs * 3776, np.uint8)
    used = ctypes.c_size_t()
    if lib.qq_shuffle_proofs_from_bincode(_ptr(d), d.size, nproofs, _ptr(out), ctypes.byref(used)) != QQ_OK:
        raise ValueError("malformed bincode ShuffleProof")
    return out.reshape(nproofs, 3776), used.value


def shuffle_statements_from_bincode(data, nproofs):
    lib = load_library()
    d = _u8(data)
    out = np.zeros(nproofs * 352, np.uint8)
    used = ctypes.c_size_t()
    if lib.qq_shuffle_statements_from_bincode(_ptr(d), d.size, nproofs, _ptr(out), ctypes.byref(used)) != QQ_OK:
        raise ValueError("malformed bincode ShuffleStatement")
    return out.reshape(nproofs, 352), used.value


def shuffle_proofs_to_bincode(proofs):
    lib = load_library()
    p = _u8(proofs)
    n = p.size // 3776
    out = np.zeros(n * 3920, np.uint8)
    wr = ctypes.c_size_t()
    if lib.qq_shuffle_proofs_to_bincode(_ptr(p), n, _ptr(out), out.size, ctypes.byref(wr)) != QQ_OK or wr.value != out.size:
        raise ValueError("qq_shuffle_proofs_to_bincode")
    return out


def shuffle_statements_to_bincode(statements):
    lib = load_library()
    p = _u8(statements)
    n = p.size // 352
    out = np.zeros(n * 360, np.uint8)
    wr = ctypes.c_size_t()
    if lib.qq_shuffle_statements_to_bincode(_ptr(p), n, _ptr(out), out.size, ctypes.byref(wr)) != QQ_OK or wr.value != out.size:
        raise ValueError("qq_shuffle_statements_to_bincode")
    return out


def accounts_from_bincode(data):
    """bincode(Vec<Account>) -> (n x 128 bytes, bytes consumed)"""
    lib = load_library()
    d = _u8(data)
    cap = max(d.size // 128, 1)
    out = np.zeros(cap * 128, np.uint8)
    n, used = ctypes.c_size_t(), ctypes.c_size_t()
    if lib.qq_accounts_from_bincode(_ptr(d), d.size, _ptr(out), cap, ctypes.byref(n), ctypes.byref(used)) != QQ_OK:
        raise ValueError("malformed bincode Vec<Account>")
    return out[:n.value * 128].reshape(n.value, 128), used.value


def sigma_proof_from_bincode(data):
    """bincode(SigmaProof) -> ("dlog", [z], x) or ("dleq", [zv, zr1, zr2], x) with the vectors as (k x 32) byte arrays."""
    lib = load_library()
    d = _u8(data)
    cap = max(d.size // 32, 1)
    out = np.zeros(cap * 32, np.uint8)
    lens = (ctypes.c_size_t * 3)()
    variant, used = ctypes.c_int(), ctypes.c_size_t()
    x = np.zeros(32, np.uint8)
    if lib.qq_sigma_proof_from_bincode(_ptr(d), d.size, ctypes.byref(variant), _ptr(out), cap, lens, _ptr(x), ctypes.byref(used)) != QQ_OK:
        raise ValueError("malformed bincode SigmaProof")
    vecs, o = [], 0
    for k in range(1 if variant.value == 0 else 3):
        vecs.append(out[32 * o:32 * (o + lens[k])].reshape(lens[k], 32).copy())
        o += lens[k]
    return ("dlog" if variant.value == 0 else "dleq"), vecs, x, used.value


class Engine:
    """One qq_ctx on one GPU.  Host-array methods take/return numpy uint8 arrays shaped (n, bytes)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = ctypes.c_void_p()
        rc = self.lib.qq_init(ctypes.byref(h), int(device))
        if rc != QQ_OK:
            raise QQError("qq_init(device=%d) failed with %d: no usable sm_100 GPU (there is no CPU fallback)" % (device, rc))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.qq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != QQ_OK:
            raise QQError("%s failed (%d): %s" % (what, rc, self.lib.qq_last_error(self.h).decode()))

    # ---- introspection -------------------------------------------------------------------------------------
    @property
    def sm_count(self):
        return self.lib.qq_device_sm_count(self.h)

    @property
    def launch_count(self):
        return int(self.lib.qq_launch_count(self.h))

    @property
    def last_kernel_ms(self):
        return float(self.lib.qq_last_kernel_ms(self.h))

    def last_kernel_breakdown(self):
        buf = (ctypes.c_float * 8)()
        k = self.lib.qq_last_kernel_breakdown(self.h, buf, 8)
        names = ["decompress", "varbase", "fixedbase", "finish", "msm_bucket", "msm_reduce", "transcripts"]
        return {names[i]: float(buf[i]) for i in range(k)}

    def measure_imad_peak(self):
        w, lo = ctypes.c_double(), ctypes.c_double()
        self._ck(self.lib.qq_measure_imad_peak(self.h, ctypes.byref(w), ctypes.byref(lo)), "qq_measure_imad_peak")
        return {"imad_wide_per_s": w.value, "imad_lo_per_s": lo.value}

    def event_record(self, slot):
        self._ck(self.lib.qq_event_record(self.h, slot), "qq_event_record")

    def event_elapsed_ms(self, a, b):
        ms = ctypes.c_float()
        self._ck(self.lib.qq_event_elapsed_ms(self.h, a, b, ctypes.byref(ms)), "qq_event_elapsed_ms")
        return float(ms.value)

    # ---- device memory ---------------------------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = ctypes.c_void_p()
        self._ck(self.lib.qq_dev_alloc(self.h, ctypes.byref(p), nbytes), "qq_dev_alloc")
        return p

    def dev_free(self, p):
        self._ck(self.lib.qq_dev_free(self.h, p), "qq_dev_free")

    def dev_upload(self, p, arr):
        a = _u8(arr)
        self._ck(self.lib.qq_dev_upload(self.h, p, _ptr(a), a.size), "qq_dev_upload")

    def dev_download(self, p, nbytes):
        out = np.empty(nbytes, dtype=np.uint8)
        self._ck(self.lib.qq_dev_download(self.h, _ptr(out), p, nbytes), "qq_dev_download")
        return out

    # ---- host-array batch calls --------------------------------------------------------------------------------
    def update_public_key(self, pk, r):
        pk, r = _u8(pk), _u8(r)
        n = r.size // 32
        _u8(pk, n * 64)
        out, st = np.zeros(n * 64, np.uint8), np.zeros(n, np.uint8)
        self._ck(self.lib.qq_update_public_key_batch(self.h, _ptr(pk), _ptr(r), _ptr(out), _ptr(st), n), "qq_update_public_key_batch")
        return out.reshape(n, 64), st

    def mul_commitment(self, comm, s):
        comm, s = _u8(comm), _u8(s)
        n = s.size // 32
        _u8(comm, n * 64)
        out, st = np.zeros(n * 64, np.uint8), np.zeros(n, np.uint8)
        self._ck(self.lib.qq_mul_commitment_batch(self.h, _ptr(comm), _ptr(s), _ptr(out), _ptr(st), n), "qq_mul_commitment_batch")
        return out.reshape(n, 64), st

    def verify_public_key_update(self, upd, pk, r):
        upd, pk, r = _u8(upd), _u8(pk), _u8(r)
        n = r.size // 32
        _u8(pk, n * 64), _u8(upd, n * 64)
        st = np.zeros(n, np.uint8)
        self._ck(self.lib.qq_verify_public_key_update_batch(self.h, _ptr(upd), _ptr(pk), _ptr(r), _ptr(st), n), "qq_verify_public_key_update_batch")
        return st

    def generate_commitment(self, pk, r, v):
        pk, r, v = _u8(pk), _u8(r), _u8(v)
        n = r.size // 32
        _u8(pk, n * 64), _u8(v, n * 32)
        out, st = np.zeros(n * 64, np.uint8), np.zeros(n, np.uint8)
        self._ck(self.lib.qq_generate_commitment_batch(self.h, _ptr(pk), _ptr(r), _ptr(v), _ptr(out), _ptr(st), n), "qq_generate_commitment_batch")
        return out.reshape(n, 64), st

    def add_commitments(self, a, b, negate_b=False):
        a, b = _u8(a), _u8(b)
        n = a.size // 64
        _u8(b, n * 64)
        out, st = np.zeros(n * 64, np.uint8), np.zeros(n, np.uint8)
        self._ck(self.lib.qq_add_commitments_batch(self.h, _ptr(a), _ptr(b), int(bool(negate_b)), _ptr(out), _ptr(st), n), "qq_add_commitments_batch")
        return out.reshape(n, 64), st

    def update_account(self, acc, bl, u, c):
        acc, bl, u, c = _u8(acc), _u8(bl), _u8(u), _u8(c)
        n = bl.size // 32
        _u8(acc, n * 128), _u8(u, n * 32), _u8(c, n * 32)
        out, st = np.zeros(n * 128, np.uint8), np.zeros(n, np.uint8)
        self._ck(self.lib.qq_update_account_batch(self.h, _ptr(acc), _ptr(bl), _ptr(u), _ptr(c), _ptr(out), _ptr(st), n), "qq_update_account_batch")
        return out.reshape(n, 128), st

    def verify_account(self, acc, sk, bl):
        acc, sk, bl = _u8(acc), _u8(sk), _u8(bl)
        n = sk.size // 32
        _u8(acc, n * 128), _u8(bl, n * 32)
        st = np.zeros(n, np.uint8)
        self._ck(self.lib.qq_verify_account_batch(self.h, _ptr(acc), _ptr(sk), _ptr(bl), _ptr(st), n), "qq_verify_account_batch")
        return st

    def delta_epsilon(self, acc, bl, r, base_pk):
        acc, bl, r, base_pk = _u8(acc), _u8(bl), _u8(r), _u8(base_pk, 64)
        n = bl.size // 32
        _u8(acc, n * 128), _u8(r, n * 32)
        d, e, st = np.zeros(n * 128, np.uint8), np.zeros(n * 128, np.uint8), np.zeros(n, np.uint8)
        self._ck(self.lib.qq_delta_epsilon_batch(self.h, _ptr(acc), _ptr(bl), _ptr(r), _ptr(base_pk), _ptr(d), _ptr(e), _ptr(st), n), "qq_delta_epsilon_batch")
        return d.reshape(n, 128), e.reshape(n, 128), st

    def delta_identity_check(self, acc):
        acc = _u8(acc)
        n = acc.size // 128
        v = np.zeros(1, np.uint8)
        self._ck(self.lib.qq_delta_identity_check(self.h, _ptr(acc), n, _ptr(v)), "qq_delta_identity_check")
        return int(v[0])

    def fixed_base(self, which, s):
        s = _u8(s)
        n = s.size // 32
        out, st = np.zeros(n * 32, np.uint8), np.zeros(n, np.uint8)
        self._ck(self.lib.qq_fixed_base_batch(self.h, int(which), _ptr(s), _ptr(out), _ptr(st), n), "qq_fixed_base_batch")
        return out.reshape(n, 32), st

    def fixed_base_i64(self, which, values):
        """enc(v * Base) for signed 64-bit values (balances)."""
        v = np.ascontiguousarray(values, dtype=np.int64).reshape(-1)
        out = np.zeros(v.size * 32, np.uint8)
        self._ck(self.lib.qq_fixed_base_i64_batch(self.h, int(which), v.ctypes.data_as(ctypes.c_void_p), _ptr(out), v.size),
                 "qq_fixed_base_i64_batch")
        return out.reshape(v.size, 32)

    def fixed_base_set_window(self, which, window_bits):
        """Rebuild the large fixed-base table of base `which` with `window_bits`-bit windows (0 frees it)."""
        self._ck(self.lib.qq_fixed_base_set_window(self.h, int(which), int(window_bits)), "qq_fixed_base_set_window")

    def varbase_set_coop_limit(self, max_scalar_mults):
        """Variable-base calls up to this many scalar mults use four lanes per multiplication (< 0: default, 0: off)."""
        self._ck(self.lib.qq_varbase_set_coop_limit(self.h, int(max_scalar_mults)), "qq_varbase_set_coop_limit")

    def fixed_base_window(self, which):
        return int(self.lib.qq_fixed_base_window(self.h, int(which)))

    def msm(self, scalars, points):
        scalars, points = _u8(scalars), _u8(points)
        n = scalars.size // 32
        _u8(points, n * 32)
        out, st = np.zeros(32, np.uint8), np.zeros(1, np.uint8)
        self._ck(self.lib.qq_msm(self.h, _ptr(scalars), _ptr(points), n, _ptr(out), _ptr(st)), "qq_msm")
        return out, int(st[0])

    def msm_partial(self, scalars, points):
        scalars, points = _u8(scalars), _u8(points)
        n = scalars.size // 32
        _u8(points, n * 32)
        out, st = np.zeros(128, np.uint8), np.zeros(1, np.uint8)
        self._ck(self.lib.qq_msm_partial(self.h, _ptr(scalars), _ptr(points), n, _ptr(out), _ptr(st)), "qq_msm_partial")
        return out, int(st[0])

    def msm_points_prepare(self, points):
        """Decompress a reusable point set once; returns an opaque handle for msm_prepared / msm_points_free."""
        points = _u8(points)
        n = points.size // 32
        h = ctypes.c_void_p()
        self._ck(self.lib.qq_msm_points_prepare(self.h, _ptr(points), n, ctypes.byref(h)), "qq_msm_points_prepare")
        return h

    def msm_points_free(self, handle):
        self.lib.qq_msm_points_free(self.h, handle)

    def msm_prepared(self, scalars, handle):
        scalars = _u8(scalars)
        n = scalars.size // 32
        out, st = np.zeros(32, np.uint8), np.zeros(1, np.uint8)
        self._ck(self.lib.qq_msm_prepared(self.h, _ptr(scalars), handle, n, _ptr(out), _ptr(st)), "qq_msm_prepared")
        return out, int(st[0])

    def verify_update_account_dlog(self, input_accounts, delta_accounts, z, x, n, transcript_label=b"UpdateAccount",
                                   verifier_label=b"DLOGProof"):
        """Verifier::verify_update_account_verifier for x.size // 32 proofs of n accounts each -> status per proof."""
        ia, da, z, x = _u8(input_accounts), _u8(delta_accounts), _u8(z), _u8(x)
        nproofs = x.size // 32
        _u8(ia, nproofs * n * 128), _u8(da, nproofs * n * 128), _u8(z, nproofs * n * 32)
        st = np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_verify_update_account_dlog_batch(self.h, transcript_label, verifier_label, _ptr(ia), _ptr(da),
                                                               _ptr(z), _ptr(x), n, nproofs, _ptr(st)),
                 "qq_verify_update_account_dlog_batch")
        return st

    def verify_delta_compact(self, delta_accounts, epsilon_accounts, zv, zr1, zr2, x, n, transcript_label=b"DeltaCompact",
                             verifier_label=b"DLEQProof"):
        """Verifier::verify_delta_compact_verifier for x.size // 32 proofs of n accounts each -> status per proof."""
        da, ea, zv, zr1, zr2, x = (_u8(a) for a in (delta_accounts, epsilon_accounts, zv, zr1, zr2, x))
        nproofs = x.size // 32
        _u8(da, nproofs * n * 128), _u8(ea, nproofs * n * 128)
        for a in (zv, zr1, zr2):
            _u8(a, nproofs * n * 32)
        st = np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_verify_delta_compact_batch(self.h, transcript_label, verifier_label, _ptr(da), _ptr(ea), _ptr(zv),
                                                         _ptr(zr1), _ptr(zr2), _ptr(x), n, nproofs, _ptr(st)),
                 "qq_verify_delta_compact_batch")
        return st

    def verify_account_sigma(self, delta_accounts, epsilon_accounts, base_pk, zv, zsk, zr, x, n,
                             transcript_label=b"SenderAccountProof", verifier_label=b"DLOGProof"):
        """Verifier::verify_account_verifier_bulletproof (= the sigma part of verify_account_verifier) for x.size // 32
        proofs of n sender accounts each -> status per proof."""
        da, ea, bp, zv, zsk, zr, x = (_u8(a) for a in (delta_accounts, epsilon_accounts, base_pk, zv, zsk, zr, x))
        nproofs = x.size // 32
        _u8(da, nproofs * n * 128), _u8(ea, nproofs * n * 128), _u8(bp, 64)
        for a in (zv, zsk, zr):
            _u8(a, nproofs * n * 32)
        st = np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_verify_account_sigma_batch(self.h, transcript_label, verifier_label, _ptr(da), _ptr(ea), _ptr(bp),
                                                         _ptr(zv), _ptr(zsk), _ptr(zr), _ptr(x), n, nproofs, _ptr(st)),
                 "qq_verify_account_sigma_batch")
        return st

    def verify_zero_balance(self, accounts, z, x, n, vector_form=True, transcript_label=b"ZeroBalanceAccount",
                            verifier_label=b"DLOGProof"):
        """Verifier::zero_balance_account_vector_verifier (vector_form) / zero_balance_account_verifier (n = 1)."""
        ac, z, x = _u8(accounts), _u8(z), _u8(x)
        nproofs = x.size // 32
        _u8(ac, nproofs * n * 128), _u8(z, nproofs * n * 32)
        st = np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_verify_zero_balance_batch(self.h, transcript_label, verifier_label, _ptr(ac), _ptr(z), _ptr(x), n,
                                                        nproofs, 1 if vector_form else 0, _ptr(st)),
                 "qq_verify_zero_balance_batch")
        return st

    def verify_destroy_account(self, accounts, z, x, n, transcript_label=b"DestroyAccount", verifier_label=b"DLOGProof"):
        """Verifier::destroy_account_verifier for x.size // 32 proofs of n accounts each -> status per proof."""
        ac, z, x = _u8(accounts), _u8(z), _u8(x)
        nproofs = x.size // 32
        _u8(ac, nproofs * n * 128), _u8(z, nproofs * n * 32)
        st = np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_verify_destroy_account_batch(self.h, transcript_label, verifier_label, _ptr(ac), _ptr(z), _ptr(x),
                                                           n, nproofs, _ptr(st)), "qq_verify_destroy_account_batch")
        return st

    def verify_same_value_compact(self, enc_accounts, commitments, zv, zr, x):
        """Verifier::verify_same_value_compact_verifier, one proof per element -> status per proof."""
        ac, cm, zv, zr, x = (_u8(a) for a in (enc_accounts, commitments, zv, zr, x))
        nproofs = x.size // 32
        _u8(ac, nproofs * 128), _u8(cm, nproofs * 32), _u8(zv, nproofs * 32), _u8(zr, nproofs * 32)
        st = np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_verify_same_value_compact_batch(self.h, _ptr(ac), _ptr(cm), _ptr(zv), _ptr(zr), _ptr(x), nproofs,
                                                              _ptr(st)), "qq_verify_same_value_compact_batch")
        return st

    def verify_update_account_dark_tx(self, delta_accounts, output_accounts, z, x, n, transcript_label=b"UpdateAccount",
                                      verifier_label=b"DLOGProof"):
        """Verifier::verify_update_account_dark_tx_verifier; z: two scalars per proof -> status per proof."""
        da, oa, z, x = (_u8(a) for a in (delta_accounts, output_accounts, z, x))
        nproofs = x.size // 32
        _u8(da, nproofs * n * 128), _u8(oa, nproofs * n * 128), _u8(z, nproofs * 64)
        st = np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_verify_update_account_dark_tx_batch(self.h, transcript_label, verifier_label, _ptr(da), _ptr(oa),
                                                                  _ptr(z), _ptr(x), n, nproofs, _ptr(st)),
                 "qq_verify_update_account_dark_tx_batch")
        return st

    def verify_ddh(self, g, h, g_dash, h_dash, challenge, z, transcript_label=b"ShuffleProof", verifier_label=b"DDHTuple"):
        """DDHProof::verify_ddh_proof, one proof per 32-byte element -> status per proof."""
        g, h, gd, hd, ch, z = (_u8(a) for a in (g, h, g_dash, h_dash, challenge, z))
        nproofs = z.size // 32
        for a in (g, h, gd, hd, ch):
            _u8(a, nproofs * 32)
        st = np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_verify_ddh_batch(self.h, transcript_label, verifier_label, _ptr(g), _ptr(h), _ptr(gd), _ptr(hd),
                                               _ptr(ch), _ptr(z), nproofs, _ptr(st)), "qq_verify_ddh_batch")
        return st

    def verify_svp(self, commitment_a, b, proof, transcript_label=b"SingleValue", verifier_label=b"Shuffle"):
        """SVPProof::verify; proof: 352 bytes per proof (see include/qq_b200.h) -> status per proof."""
        ca, b, pr = _u8(commitment_a), _u8(b), _u8(proof)
        nproofs = b.size // 32
        _u8(ca, nproofs * 32), _u8(pr, nproofs * 352)
        st = np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_verify_svp_batch(self.h, transcript_label, verifier_label, _ptr(ca), _ptr(b), _ptr(pr), nproofs,
                                               _ptr(st)), "qq_verify_svp_batch")
        return st

    def verify_hadamard(self, omega, commit_a, commit_b, commit_c, proof, transcript_label=b"Hadamard",
                        verifier_label=b"Shuffle"):
        """HadamardProof::verify; proof: 640 bytes per proof -> (status, detail) per proof."""
        om, ca, cb, cc, pr = (_u8(a) for a in (omega, commit_a, commit_b, commit_c, proof))
        nproofs = om.size // 96
        _u8(ca, nproofs * 96), _u8(cb, nproofs * 96), _u8(cc, nproofs * 96), _u8(pr, nproofs * 640)
        st, det = np.zeros(nproofs, np.uint8), np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_verify_hadamard_batch(self.h, transcript_label, verifier_label, _ptr(om), _ptr(ca), _ptr(cb),
                                                    _ptr(cc), _ptr(pr), nproofs, _ptr(st), _ptr(det)),
                 "qq_verify_hadamard_batch")
        return st, det

    def verify_product(self, c_prod_A, statement, proof, transcript_label=b"ShuffleProof", verifier_label=b"Shuffle"):
        """ProductProof::verify; statement 192 B, proof 1024 B per proof (include/qq_b200.h) -> (status, detail)."""
        ca, stm, pr = _u8(c_prod_A), _u8(statement), _u8(proof)
        nproofs = ca.size // 96
        _u8(stm, nproofs * 192), _u8(pr, nproofs * 1024)
        st, det = np.zeros(nproofs, np.uint8), np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_verify_product_batch(self.h, transcript_label, verifier_label, _ptr(ca), _ptr(stm), _ptr(pr),
                                                   nproofs, _ptr(st), _ptr(det)), "qq_verify_product_batch")
        return st, det

    def verify_shuffle(self, shuffle_input, shuffle_output, statement, proof, transcript_label=b"ShuffleProof",
                       verifier_label=b"Shuffle"):
        """ShuffleProof::verify for proof.size // 3776 proofs over 9 accounts each -> (status, stage, detail) per proof."""
        si, so, stm, pr = (_u8(a) for a in (shuffle_input, shuffle_output, statement, proof))
        nproofs = pr.size // 3776
        _u8(si, nproofs * 9 * 128), _u8(so, nproofs * 9 * 128), _u8(stm, nproofs * 352), _u8(pr, nproofs * 3776)
        st, sg, det = (np.zeros(nproofs, np.uint8) for _ in range(3))
        self._ck(self.lib.qq_verify_shuffle_batch(self.h, transcript_label, verifier_label, _ptr(si), _ptr(so), _ptr(stm), _ptr(pr),
                                                   nproofs, _ptr(st), _ptr(sg), _ptr(det)), "qq_verify_shuffle_batch")
        return st, sg, det

    def verify_set_transcripts(self, on_device=True):
        """Shuffle verifier: per-proof transcripts / scalar algebra in GPU transcript kernels (default) or on the host threads."""
        self._ck(self.lib.qq_verify_set_transcripts(self.h, 1 if on_device else 0), "qq_verify_set_transcripts")

    def msm_set_shifted(self, budget_bytes=112 << 20, use_it=True):
        """Shifted form of prepared point sets (see the header): memory budget per set, and whether qq_msm_prepared uses it."""
        self._ck(self.lib.qq_msm_set_shifted(self.h, budget_bytes, 1 if use_it else 0), "qq_msm_set_shifted")

    def set_secret_mode(self, on=True):
        """Constant-time table access for the scalar multiplications of the wallet / prover entry points (see the header)."""
        self._ck(self.lib.qq_set_secret_mode(self.h, 1 if on else 0), "qq_set_secret_mode")

    def sigma_commit(self, points, r, v=None):
        """out_i = enc(r_i * P_i [+ v_i * B]): the provers' e / f commitment maps (src/accounts/prover.rs)."""
        points, r = _u8(points), _u8(r)
        n = points.size // 32
        vv = _u8(v) if v is not None else None
        out, st = np.zeros(n * 32, np.uint8), np.zeros(n, np.uint8)
        self._ck(self.lib.qq_sigma_commit_batch(self.h, _ptr(points), _ptr(r), _ptr(vv) if vv is not None else None, _ptr(out), _ptr(st), n),
                 "qq_sigma_commit_batch")
        return out.reshape(n, 32), st

    def verify_set_aggregation(self, on=True):
        """Shuffle verifier: identity equations of all proofs in one weighted Pippenger MSM (default) or one MSM per equation."""
        self._ck(self.lib.qq_verify_set_aggregation(self.h, 1 if on else 0), "qq_verify_set_aggregation")

    def msm_set_overlap(self, split_min=1 << 17, tail_pct=30, sort_blocks_per_sm=3):
        """Large-MSM tuning (decompression of the last tail_pct % of the points under the counting sort); tail_pct 0 = off."""
        self._ck(self.lib.qq_msm_set_overlap(self.h, split_min, tail_pct, sort_blocks_per_sm), "qq_msm_set_overlap")

    def range_proof_bytes(self, m, n_bits=64):
        """Length of RangeProof::to_bytes() for m aggregated n_bits-bit values."""
        return (9 + 2 * ((n_bits * m).bit_length() - 1)) * 32

    def verify_range_proofs(self, commitments, proofs, m, chain=1, n_bits=64, transcript_label=b"SenderAccountProof",
                            verifier_label=b"BulletProof", domain_label=b"AggregateBulletProof", transcript_state=None):
        """RangeProof::verify_multiple (chain = 1, m values per proof) / a chain of verify_single calls on one transcript (m = 1),
        batched over independent transcripts -> status per transcript.  transcript_state: the array transcript_capture()
        filled during an earlier verification on the same transcripts (replaces the two labels)."""
        cm, pr = _u8(commitments), _u8(proofs)
        per = self.range_proof_bytes(m, n_bits) * chain
        nproofs = pr.size // per
        _u8(pr, nproofs * per), _u8(cm, nproofs * chain * m * 32)
        ts = None
        if transcript_state is not None:
            ts = _u8(transcript_state, nproofs * self.lib.qq_transcript_state_bytes())
        st = np.zeros(nproofs, np.uint8)
        self._ck(self.lib.qq_verify_range_proof_batch(self.h, transcript_label, verifier_label, _ptr(ts) if ts is not None else None,
                                                       domain_label, _ptr(cm), _ptr(pr), n_bits, m, chain, nproofs, _ptr(st)),
                 "qq_verify_range_proof_batch")
        return st

    def transcript_capture(self, nproofs):
        """Arms the one-shot capture: the next sigma verification call leaves its nproofs transcripts in the returned array."""
        buf = np.zeros(nproofs * self.lib.qq_transcript_state_bytes(), np.uint8)
        self._ck(self.lib.qq_transcript_capture(self.h, _ptr(buf), nproofs), "qq_transcript_capture")
        return buf

    def decommit(self, comm, sk):
        comm, sk = _u8(comm), _u8(sk)
        n = sk.size // 32
        _u8(comm, n * 64)
        out, st = np.zeros(n * 32, np.uint8), np.zeros(n, np.uint8)
        self._ck(self.lib.qq_decommit_batch(self.h, _ptr(comm), _ptr(sk), _ptr(out), _ptr(st), n), "qq_decommit_batch")
        return out.reshape(n, 32), st

    def decommit_value(self, comm, sk, search_bits=32):
        comm, sk = _u8(comm), _u8(sk)
        n = sk.size // 32
        _u8(comm, n * 64)
        vals, st = np.zeros(n, np.uint64), np.zeros(n, np.uint8)
        self._ck(self.lib.qq_decommit_value_batch(self.h, _ptr(comm), _ptr(sk), int(search_bits),
                                                   vals.ctypes.data_as(ctypes.c_void_p), _ptr(st), n), "qq_decommit_value_batch")
        return vals, st

    def from_uniform_bytes(self, uniform64):
        u = _u8(uniform64)
        n = u.size // 64
        out = np.zeros(n * 32, np.uint8)
        self._ck(self.lib.qq_from_uniform_bytes_batch(self.h, _ptr(u), _ptr(out), n), "qq_from_uniform_bytes_batch")
        return out.reshape(n, 32)

    def vector_pedersen_gens(self, capacity):
        """(H, G_vec) of VectorPedersenGens::new(capacity): 32 bytes and (capacity - 1) x 32 bytes."""
        h, g = np.zeros(32, np.uint8), np.zeros((capacity - 1) * 32, np.uint8)
        self._ck(self.lib.qq_vector_pedersen_gens(self.h, capacity, _ptr(h), _ptr(g)), "qq_vector_pedersen_gens")
        return h, g.reshape(capacity - 1, 32)

    def bulletproof_gens(self, gens_capacity, party_capacity):
        n = gens_capacity * party_capacity
        g, h = np.zeros(n * 32, np.uint8), np.zeros(n * 32, np.uint8)
        self._ck(self.lib.qq_bulletproof_gens(self.h, gens_capacity, party_capacity, _ptr(g), _ptr(h)), "qq_bulletproof_gens")
        return g.reshape(party_capacity, gens_capacity, 32), h.reshape(party_capacity, gens_capacity, 32)

    def points_sum(self, xyzt):
        xyzt = _u8(xyzt)
        k = xyzt.size // 128
        out, ident = np.zeros(32, np.uint8), np.zeros(1, np.uint8)
        self._ck(self.lib.qq_points_sum(self.h, _ptr(xyzt), k, _ptr(out), _ptr(ident)), "qq_points_sum")
        return out, bool(ident[0])

    def warp_ops_selftest(self, p_limbs, q_limbs):
        """parity hook of the warp-cooperative group operations: (doubling formula of p, addition formula of p and q), raw limbs"""
        p_limbs, q_limbs = _u8(p_limbs), _u8(q_limbs)
        n = p_limbs.size // 128
        if q_limbs.size != p_limbs.size or p_limbs.size % 128:
            raise ValueError("p and q must hold the same number of 128-byte points")
        d, a = np.zeros(n * 128, np.uint8), np.zeros(n * 128, np.uint8)
        self._ck(self.lib.qq_warp_ops_selftest(self.h, _ptr(p_limbs), _ptr(q_limbs), n, _ptr(d), _ptr(a)), "qq_warp_ops_selftest")
        return d.reshape(n, 128), a.reshape(n, 128)

    def points_sum_dev(self, records_ptr, k, stride=144):
        """sum of k partial MSM results resident in device memory (records of `stride` bytes) -> (compressed, identity?, status)"""
        out, ident, st = np.zeros(32, np.uint8), np.zeros(1, np.uint8), np.zeros(1, np.uint8)
        self._ck(self.lib.qq_points_sum_dev(self.h, ctypes.c_void_p(records_ptr), k, stride, _ptr(out), _ptr(ident), _ptr(st)),
                 "qq_points_sum_dev")
        return out, bool(ident[0]), int(st[0])

    def msm_segmented(self, scalars, points, offsets):
        scalars, points = _u8(scalars), _u8(points)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        m = offsets.size - 1
        out, st = np.zeros(m * 32, np.uint8), np.zeros(m, np.uint8)
        self._ck(self.lib.qq_msm_segmented(self.h, _ptr(scalars), _ptr(points), ctypes.c_void_p(offsets.ctypes.data), m, _ptr(out), _ptr(st)), "qq_msm_segmented")
        return out.reshape(m, 32), st

    def msm_grouped(self, scalars, points, offsets):
        """m independent large MSMs in one Pippenger pass (CSR offsets as msm_segmented) -> (m x 32 encodings, m status bytes)"""
        scalars, points = _u8(scalars), _u8(points)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        m = offsets.size - 1
        out, st = np.zeros(m * 32, np.uint8), np.zeros(m, np.uint8)
        self._ck(self.lib.qq_msm_grouped(self.h, _ptr(scalars), _ptr(points), ctypes.c_void_p(offsets.ctypes.data), m, _ptr(out), _ptr(st)), "qq_msm_grouped")
        return out.reshape(m, 32), st

    # ---- device-pointer calls (pointers are ints / c_void_p on this engine's GPU) ---------------------------------
    def call_dev(self, name, *args):
        """Call a `_dev` entry point; args are ctypes values (c_void_p device pointers, c_size_t counts, ints)."""
        self._ck(getattr(self.lib, name)(self.h, *args), name)
