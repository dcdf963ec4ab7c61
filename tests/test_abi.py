"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/qq_b200.h declares, and the product
path fails loudly (no CPU fallback) when no B200 is present."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "qq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qq_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(pkg):
    lib = ctypes.CDLL(pkg.lib_path())
    syms = declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), "libqq_b200.so does not export %s" % s


def test_binding_covers_header(pkg):
    from quisquis_rust_b200.binding import EXPORTS
    assert set(EXPORTS) == set(declared_symbols())


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; this test checks the no-GPU failure mode")
    with pytest.raises(pkg.QQError, match="no CPU fallback"):
        pkg.Engine(0)


def test_product_does_not_reference_oracle():
    """Nothing under quisquis-rust_b200/ may import, link or call the oracle."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "quisquis-rust_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".inc", ".hpp", ".h", ".cpp")):
                t = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"(import\s+ristretto_ref|import\s+c_oracle|libqq_oracle|qq_oracle\.c|#include\s+\"[^\"]*oracle)", t):
                    bad.append(f)
    assert not bad, bad


def _bincode_vec32(items):
    return len(items).to_bytes(8, "little") + b"".join(items)


def _bincode_shuffle_proof(flat):
    """bincode 1.x of a reference ShuffleProof (src/shuffle/shuffle.rs:164-184) from its 3 776 flattened bytes, written out
    field by field independently of the library's table (Vec<T> = u64 little-endian length + elements)."""
    o = [0]

    def take(n):
        b = flat[o[0]:o[0] + n]
        o[0] += n
        return b

    def vec(k):
        return _bincode_vec32([take(32) for _ in range(k)])
    out = vec(3) + vec(3) + vec(3) + vec(3)                              # c_A, c_tau, c_B, c_B_dash
    out += take(640)                                                     # HadamardProof: arrays and scalars only
    out += vec(3) + take(64) + vec(7) + vec(3) + vec(3) + take(96)       # MultiHadamardProof{c_B, ZeroProof}
    out += take(96) + vec(3) + vec(3) + take(64)                         # SVPProof
    for _ in range(2):                                                   # multi_exponen_pk, multi_exponen_commit
        out += take(32) + vec(6) + vec(6) + vec(6) + vec(3) + take(128)
    out += take(64)                                                      # DDHProof
    assert o[0] == 3776
    return out


def test_bincode_wire_format_round_trip(pkg):
    """bincode(ShuffleProof / ShuffleStatement / Vec<Account> / SigmaProof) <-> the flattened layouts.  The
    golden proofs are encoded by an independent Python bincode writer; the library's reader must return the committed flattened
    bytes, its writer the same bincode; truncated input and wrong Vec lengths are refused."""
    import numpy as np
    from quisquis_rust_b200 import binding as B
    raw = np.fromfile(os.path.join(os.path.dirname(__file__), "golden", "shuffle_proofs.bin"), dtype=np.uint8).reshape(-1, 6432)
    n = raw.shape[0]
    proofs = [raw[i, 2656:].tobytes() for i in range(n)]
    stms = [raw[i, 2304:2656].tobytes() for i in range(n)]
    enc = b"".join(_bincode_shuffle_proof(p) for p in proofs)
    assert len(enc) == n * 3920
    flat, used = B.shuffle_proofs_from_bincode(enc, n)
    assert used == len(enc) and flat.tobytes() == b"".join(proofs)
    assert B.shuffle_proofs_to_bincode(flat).tobytes() == enc
    senc = b"".join(s[:128] + _bincode_vec32([s[128 + 32 * k:160 + 32 * k] for k in range(3)]) + s[224:] for s in stms)
    assert len(senc) == n * 360
    sflat, used = B.shuffle_statements_from_bincode(senc, n)
    assert used == len(senc) and sflat.tobytes() == b"".join(stms)
    assert B.shuffle_statements_to_bincode(sflat).tobytes() == senc
    import pytest
    with pytest.raises(ValueError):
        B.shuffle_proofs_from_bincode(enc[:-1], n)                       # truncated
    bad = bytearray(enc)
    bad[0] = 4                                                           # c_A announced with 4 entries
    with pytest.raises(ValueError):
        B.shuffle_proofs_from_bincode(bytes(bad), n)
    # Vec<Account> and SigmaProof
    accs = [raw[0, 128 * i:128 * (i + 1)].tobytes() for i in range(9)]
    got, used = B.accounts_from_bincode(_bincode_vec32(accs) + b"tail")
    assert used == 8 + 9 * 128 and got.tobytes() == b"".join(accs)
    sc = [bytes([i + 1]) + bytes(31) for i in range(7)]
    kind, vecs, x, used = B.sigma_proof_from_bincode((0).to_bytes(4, "little") + _bincode_vec32(sc[:3]) + sc[6])
    assert kind == "dlog" and vecs[0].tobytes() == b"".join(sc[:3]) and x.tobytes() == sc[6] and used == 4 + 8 + 96 + 32
    kind, vecs, x, used = B.sigma_proof_from_bincode((1).to_bytes(4, "little") + _bincode_vec32(sc[:2]) + _bincode_vec32(sc[2:4]) +
                                                     _bincode_vec32(sc[4:6]) + sc[6])
    assert kind == "dleq" and [v.tobytes() for v in vecs] == [b"".join(sc[:2]), b"".join(sc[2:4]), b"".join(sc[4:6])]
    with pytest.raises(ValueError):
        B.sigma_proof_from_bincode((2).to_bytes(4, "little") + _bincode_vec32(sc[:3]) + sc[6])      # unknown variant
    with pytest.raises(ValueError):
        B.sigma_proof_from_bincode((0).to_bytes(4, "little") + (1 << 40).to_bytes(8, "little") + sc[6])   # absurd length
