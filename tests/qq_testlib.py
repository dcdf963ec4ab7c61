"""Shared helpers for the parity tests: seeded synthetic inputs and edge cases."""
import hashlib

import numpy as np

import ristretto_ref as R

SEED = b"QUISQUIS"


class Stream:
    """Counter-mode SHAKE256 stream keyed by the ASCII seed 'QUISQUIS' (0x5155495351554953)."""

    def __init__(self, label=b""):
        self.label = label
        self.ctr = 0

    def bytes(self, n):
        out = hashlib.shake_256(SEED + self.label + self.ctr.to_bytes(8, "little")).digest(n)
        self.ctr += 1
        return out

    def scalar(self):
        return int.from_bytes(self.bytes(64), "little") % R.L

    def scalar_bytes(self):
        return self.scalar().to_bytes(32, "little")


def sb(k):
    return (k % R.L).to_bytes(32, "little")


def cat(items):
    return np.frombuffer(b"".join(items), dtype=np.uint8).copy()


def make_account(st, value=0):
    """Account as in the reference tests (src/shuffle/shuffle.rs:762-768): pk = (rho*B, sk*rho*B), comm = commit(pk, k, v)."""
    sk, rho, k = st.scalar(), st.scalar(), st.scalar()
    gr = R.mul(rho, R.BASEPOINT)
    pk = R.compress(gr) + R.compress(R.mul(sk, gr))
    comm, s = R.generate_commitment(pk, sb(k), sb(value))
    assert s == 0
    return pk + comm, sk, k


def invalid_encodings():
    """One representative per reject class of RFC 9496 4.3.1 (each verified against the oracle in test_oracle.py)."""
    out = []
    out.append(("non_canonical_p", R.P.to_bytes(32, "little")))
    out.append(("non_canonical_p_plus_2", (R.P + 2).to_bytes(32, "little")))
    out.append(("bit255_set", (2 | (1 << 255)).to_bytes(32, "little")))
    out.append(("all_ff", b"\xff" * 32))
    out.append(("negative_s", (1).to_bytes(32, "little")))
    # search small even s for the remaining classes
    found = {}
    s = 2
    while len(found) < 3 and s < 4000:
        b = s.to_bytes(32, "little")
        if R.decompress(b) is None:
            ss = s * s % R.P
            u1, u2 = (1 - ss) % R.P, (1 + ss) % R.P
            v = (-(R.D * u1 * u1) - u2 * u2) % R.P
            ok, inv = R.sqrt_ratio_i(1, v * u2 * u2 % R.P)
            if not ok:
                found.setdefault("non_square", b)
            else:
                dx = inv * u2 % R.P
                x = R._abs(2 * s * dx % R.P)
                y = u1 * (inv * dx % R.P * v % R.P) % R.P
                if y == 0:
                    found.setdefault("y_zero", b)
                elif (x * y % R.P) & 1:
                    found.setdefault("negative_t", b)
        s += 2
    out.extend(found.items())
    return out
