"""Host-side mirror of the reference's operator interface for the hot path (same names, argument meaning and error
behaviour), routed through libqq_b200.so.  Scalar (n = 1) forms follow the reference signatures; `*_batch` forms are
the data-parallel entry points the Rust shim would add.

Reference: src/keys.rs:33-126 (trait PublicKey), src/ristretto/keys.rs, src/elgamal/elgamal.rs,
src/accounts/accounts.rs, src/accounts/verifier.rs:91-99,566-581.
"""
import numpy as np

from . import binding as B

_engine = None


def default_engine():
    global _engine
    if _engine is None:
        _engine = B.Engine(0)
    return _engine


def set_default_engine(e):
    global _engine
    _engine = e


class PanicError(RuntimeError):
    """The reference panics here (`.unwrap()` on a failed decompress)."""


def _b(x, n):
    a = np.frombuffer(bytes(x), dtype=np.uint8) if not isinstance(x, np.ndarray) else x
    if a.size != n:
        raise ValueError("expected %d bytes" % n)
    return a


class RistrettoPublicKey:
    """src/ristretto/keys.rs:75-79; 64 bytes gr||grsk."""

    def __init__(self, data):
        self.data = bytes(_b(data, 64))

    def as_bytes(self):
        return self.data

    @staticmethod
    def generate_base_pk():
        """src/ristretto/keys.rs:171-177 -> BASE_PK_BTC_COMPRESSED."""
        return RistrettoPublicKey(bytes.fromhex(
            "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76"
            "8c9240b456a9e6dc65c377a1048d745f94a08cdb7f44cbcd7b46f34048871134"))

    @staticmethod
    def update_public_key(p, rscalar):
        """src/ristretto/keys.rs:146-148."""
        out, st = default_engine().update_public_key(_b(p.data, 64), _b(rscalar, 32))
        if st[0] == B.ST_BAD_POINT:
            raise PanicError("called `Option::unwrap()` on a `None` value")
        if st[0]:
            raise ValueError("non-canonical scalar")
        return RistrettoPublicKey(out[0].tobytes())

    @staticmethod
    def update_public_key_batch(pks, rscalars, engine=None):
        return (engine or default_engine()).update_public_key(pks, rscalars)

    @staticmethod
    def verify_public_key_update(u, p, rscalar):
        """src/ristretto/keys.rs:161-169 -> bool."""
        st = default_engine().verify_public_key_update(_b(u.data, 64), _b(p.data, 64), _b(rscalar, 32))
        if st[0] == B.ST_BAD_POINT:
            raise PanicError("called `Option::unwrap()` on a `None` value")
        return st[0] == 0

    def verify_keypair(self, sk):
        """src/ristretto/keys.rs:187-195 -> None or raises ValueError(msg) like Err(msg)."""
        acc = self.data + bytes(64)
        st = default_engine().verify_account(_b(acc, 128), _b(sk, 32), np.zeros(32, np.uint8))
        if st[0] == B.ST_BAD_POINT:
            raise ValueError("Error::Decompression Failed")
        if st[0] == B.ST_KEYPAIR:
            raise ValueError("Invalid Account::Keypair Verification Failed")
        return None

    def __eq__(self, o):
        return self.data == o.data


class ElGamalCommitment:
    """src/elgamal/elgamal.rs:18-22; 64 bytes c||d."""

    def __init__(self, data):
        self.data = bytes(_b(data, 64))

    def to_bytes(self):
        return self.data

    @staticmethod
    def generate_commitment(p, rscalar, bl_scalar):
        """src/elgamal/elgamal.rs:41-53."""
        out, st = default_engine().generate_commitment(_b(p.data, 64), _b(rscalar, 32), _b(bl_scalar, 32))
        if st[0] == B.ST_BAD_POINT:
            raise PanicError("called `Option::unwrap()` on a `None` value")
        if st[0]:
            raise ValueError("non-canonical scalar")
        return ElGamalCommitment(out[0].tobytes())

    @staticmethod
    def generate_commitment_batch(pks, rscalars, bl_scalars, engine=None):
        return (engine or default_engine()).generate_commitment(pks, rscalars, bl_scalars)

    @staticmethod
    def add_commitments(a, b):
        """src/elgamal/elgamal.rs:65-69."""
        out, st = default_engine().add_commitments(_b(a.data, 64), _b(b.data, 64))
        if st[0]:
            raise PanicError("called `Option::unwrap()` on a `None` value")
        return ElGamalCommitment(out[0].tobytes())

    def __sub__(self, other):
        """src/elgamal/elgamal.rs:201-218."""
        out, st = default_engine().add_commitments(_b(self.data, 64), _b(other.data, 64), negate_b=True)
        if st[0]:
            raise PanicError("called `Option::unwrap()` on a `None` value")
        return ElGamalCommitment(out[0].tobytes())

    def __mul__(self, scalar):
        """src/elgamal/elgamal.rs:220-236."""
        out, st = default_engine().mul_commitment(_b(self.data, 64), _b(scalar, 32))
        if st[0] == B.ST_BAD_POINT:
            raise PanicError("called `Option::unwrap()` on a `None` value")
        return ElGamalCommitment(out[0].tobytes())

    def decommit(self, pr):
        """src/elgamal/elgamal.rs:106-108 -> compressed G*v (32 bytes)."""
        out, st = default_engine().decommit(_b(self.data, 64), _b(pr, 32))
        if st[0] == B.ST_BAD_POINT:
            raise PanicError("called `Option::unwrap()` on a `None` value")
        return out[0].tobytes()

    def decommit_value(self, pr, search_bits=40):
        """src/elgamal/elgamal.rs:119-122 -> Some(v) as int, or None.  The reference searches every u64 in ascending order;
        here the search space is 2^search_bits (baby-step / giant-step on the GPU)."""
        vals, st = default_engine().decommit_value(_b(self.data, 64), _b(pr, 32), search_bits)
        if st[0] == B.ST_BAD_POINT:
            raise PanicError("called `Option::unwrap()` on a `None` value")
        return int(vals[0]) if st[0] == 0 else None

    def __eq__(self, o):
        return self.data == o.data


class Account:
    """src/accounts/accounts.rs:47-53; 128 bytes pk||comm."""

    def __init__(self, data):
        self.data = bytes(_b(data, 128))

    @staticmethod
    def set_account(pk, comm):
        return Account(pk.data + comm.data)

    @property
    def pk(self):
        return RistrettoPublicKey(self.data[:64])

    @property
    def comm(self):
        return ElGamalCommitment(self.data[64:])

    @staticmethod
    def update_account(a, bl, update_key_scalar, generate_commitment_scalar):
        """src/accounts/accounts.rs:143-154."""
        out, st = default_engine().update_account(_b(a.data, 128), _b(bl, 32), _b(update_key_scalar, 32),
                                                  _b(generate_commitment_scalar, 32))
        if st[0] == B.ST_BAD_POINT:
            raise PanicError("called `Option::unwrap()` on a `None` value")
        if st[0]:
            raise ValueError("non-canonical scalar")
        return Account(out[0].tobytes())

    @staticmethod
    def update_account_batch(accounts, bl, u, c, engine=None):
        return (engine or default_engine()).update_account(accounts, bl, u, c)

    def verify_account(self, sk, bl):
        """src/accounts/accounts.rs:81-84 -> None or raises ValueError(msg) like Err(msg)."""
        st = default_engine().verify_account(_b(self.data, 128), _b(sk, 32), _b(bl, 32))
        msg = {B.ST_BAD_POINT: "Error::Decompression Failed",
               B.ST_KEYPAIR: "Invalid Account::Keypair Verification Failed",
               B.ST_COMMIT: "Invalid Account::Commitment Verification Failed",
               B.ST_BAD_SCALAR: "non-canonical scalar"}
        if st[0]:
            raise ValueError(msg[int(st[0])])
        return None

    def decrypt_account_balance(self, sk, bl):
        """src/accounts/accounts.rs:103-110."""
        self.verify_account(sk, bl)
        return self.comm.decommit(sk)

    def decrypt_account_balance_value(self, sk, search_bits=40):
        """src/accounts/accounts.rs:119-128."""
        self.pk.verify_keypair(sk)
        v = self.comm.decommit_value(sk, search_bits)
        if v is None:
            raise ValueError("Decryption value failed.")
        return v

    @staticmethod
    def verify_account_batch(accounts, sks, bls, engine=None):
        return (engine or default_engine()).verify_account(accounts, sks, bls)

    @staticmethod
    def verify_account_update(updated_input_accounts, accounts, updated_keys_scalar, generate_commitment_scalar):
        """src/accounts/accounts.rs:173-193: recompute update_account with bl = 0 for exactly 9 accounts (quirk kept:
        shorter input raises, like the reference's index panic) and compare all 128 bytes."""
        if min(len(accounts), len(updated_keys_scalar), len(generate_commitment_scalar)) < 9:
            raise PanicError("index out of bounds")
        acc = np.frombuffer(b"".join(a.data for a in accounts[:9]), np.uint8)
        u = np.frombuffer(b"".join(bytes(x) for x in updated_keys_scalar[:9]), np.uint8)
        c = np.frombuffer(b"".join(bytes(x) for x in generate_commitment_scalar[:9]), np.uint8)
        out, st = default_engine().update_account(acc, np.zeros(9 * 32, np.uint8), u, c)
        if (st == B.ST_BAD_POINT).any():
            raise PanicError("called `Option::unwrap()` on a `None` value")
        return all(out[i].tobytes() == ua.data for i, ua in zip(range(9), updated_input_accounts))

    @staticmethod
    def create_delta_and_epsilon_accounts(a, bl, base_pk, rscalar):
        """src/accounts/accounts.rs:198-220 with the randomness `rscalar` supplied (the reference draws it inside)."""
        n = len(a)
        acc = np.frombuffer(b"".join(x.data for x in a), np.uint8)
        blb = np.frombuffer(b"".join(bytes(x) for x in bl), np.uint8)
        rb = np.frombuffer(b"".join(bytes(x) for x in rscalar), np.uint8)
        d, e, st = default_engine().delta_epsilon(acc, blb, rb, _b(base_pk.data, 64))
        if (st == B.ST_BAD_POINT).any():
            raise PanicError("called `Option::unwrap()` on a `None` value")
        return ([Account(d[i].tobytes()) for i in range(n)], [Account(e[i].tobytes()) for i in range(n)], list(rscalar))

    def __eq__(self, o):
        return self.data == o.data


def _same_len(*seqs):
    """The batched verifiers take n accounts and n responses per proof.  The reference zips its vectors (the shortest wins for
    the commitments while the transcript absorbs every account), so lists of different lengths can only fail there; here they
    are refused up front instead of silently building a different transcript."""
    n = len(seqs[0])
    if any(len(q) != n for q in seqs):
        raise ValueError("account and response vectors must have the same length (%s)" % ", ".join(str(len(q)) for q in seqs))
    return n


class Verifier:
    @staticmethod
    def multiscalar_multiplication(combined_scalars, point):
        """src/accounts/verifier.rs:91-99 -> 32-byte compressed point, or None if any point fails to decompress."""
        s = np.frombuffer(b"".join(bytes(x) for x in combined_scalars), np.uint8)
        p = np.frombuffer(b"".join(bytes(x) for x in point), np.uint8)
        out, st = default_engine().msm(s, p)
        return None if st else out.tobytes()

    @staticmethod
    def verify_delta_identity_check(epsilon_accounts):
        """src/accounts/verifier.rs:566-581."""
        acc = np.frombuffer(b"".join(x.data for x in epsilon_accounts), np.uint8)
        v = default_engine().delta_identity_check(acc)
        if v == B.ST_BAD_POINT:
            raise PanicError("called `Option::unwrap()` on a `None` value")
        if v:
            raise ValueError("Identity sum verify: Failed")
        return None

    @staticmethod
    def verify_update_account_verifier(updated_input_accounts, updated_delta_accounts, z_vector, x,
                                       transcript_label=b"UpdateAccount", verifier_label=b"DLOGProof"):
        """src/accounts/verifier.rs:223-292.  The reference receives a `Verifier` carrying a transcript; here the two labels
        that built it (Transcript::new / Verifier::new) are passed instead.  Returns None or raises ValueError(msg)."""
        n = _same_len(updated_input_accounts, updated_delta_accounts, z_vector)
        ia = b"".join(a.data for a in updated_input_accounts)
        da = b"".join(a.data for a in updated_delta_accounts)
        st = default_engine().verify_update_account_dlog(ia, da, b"".join(bytes(v) for v in z_vector), bytes(x), n,
                                                          transcript_label, verifier_label)
        if st[0] == B.ST_BAD_POINT:
            raise PanicError("called `Option::unwrap()` on a `None` value")
        if st[0]:
            raise ValueError("DLOG Proof Verify: Failed")
        return None

    @staticmethod
    def verify_delta_compact_verifier(delta_accounts, epsilon_accounts, zv_vector, zr1_vector, zr2_vector, x,
                                      transcript_label=b"DeltaCompact", verifier_label=b"DLEQProof"):
        """src/accounts/verifier.rs:138-209.  Returns None or raises ValueError with the reference's message."""
        n = _same_len(delta_accounts, epsilon_accounts, zv_vector, zr1_vector, zr2_vector)
        da = b"".join(a.data for a in delta_accounts)
        ea = b"".join(a.data for a in epsilon_accounts)
        j = lambda v: b"".join(bytes(s) for s in v)  # noqa: E731
        st = default_engine().verify_delta_compact(da, ea, j(zv_vector), j(zr1_vector), j(zr2_vector), bytes(x), n,
                                                    transcript_label, verifier_label)
        if st[0] == B.ST_BAD_POINT:
            raise ValueError("Delta Compact Proof Verify: Failed")
        if st[0]:
            raise ValueError("Dleq Proof Verify: Failed")
        return None

    @staticmethod
    def verify_account_verifier_bulletproof(updated_delta_account_sender, account_epsilon_sender, base_pk, zv, zsk, zr, x,
                                            transcript_label=b"SenderAccountProof", verifier_label=b"DLOGProof",
                                            keep_transcript=False):
        """src/accounts/verifier.rs:396-470 (and the sigma-protocol part of verify_account_verifier, :305-381).
        keep_transcript: return the verifier's running transcript (opaque bytes) for the range-proof verification that
        follows on the same Verifier in the reference (verifier.rs:1603-1628)."""
        n = len(zv)
        j = lambda v: b"".join(bytes(s) for s in v)  # noqa: E731
        state = default_engine().transcript_capture(1) if keep_transcript else None
        st = default_engine().verify_account_sigma(b"".join(a.data for a in updated_delta_account_sender),
                                                   b"".join(a.data for a in account_epsilon_sender), base_pk.data, j(zv),
                                                   j(zsk), j(zr), bytes(x), n, transcript_label, verifier_label)
        if st[0] == B.ST_BAD_POINT:
            raise ValueError("Account Verify: Failed")
        if st[0]:
            raise ValueError("sender account verification failed")
        return state

    @staticmethod
    def verify_non_negative_sender_receiver_bulletproof_batch_verifier(epsilon_account, proof, transcript_label=b"SenderAccountProof",
                                                                       verifier_label=b"BulletProof", transcript=None):
        """src/accounts/verifier.rs:504-523: one aggregated 64-bit range proof over the d components of the epsilon accounts.
        proof: RangeProof::to_bytes().  transcript: what verify_account_verifier_bulletproof(keep_transcript=True) returned."""
        eng = default_engine()
        m = len(epsilon_account)
        proof = bytes(proof)
        if m == 0 or m > 16 or m & (m - 1) or len(proof) != eng.range_proof_bytes(m):
            raise ValueError("Bulletproof verification failed")
        st = eng.verify_range_proofs(b"".join(a.data[96:128] for a in epsilon_account), proof, m, 1, 64, transcript_label,
                                     verifier_label, b"AggregateBulletProof", transcript)
        if st[0]:
            raise ValueError("Bulletproof verification failed")
        return None

    @staticmethod
    def verify_non_negative_sender_receiver_bulletproof_vector_verifier(epsilon_account, proof_vector,
                                                                        transcript_label=b"SenderAccountProof",
                                                                        verifier_label=b"BulletProof", transcript=None):
        """src/accounts/verifier.rs:534-555: one single-value range proof per epsilon account, all on one transcript."""
        eng = default_engine()
        k = min(len(epsilon_account), len(proof_vector))          # zip() stops at the shorter one
        if k == 0:
            return None
        proofs = [bytes(p) for p in proof_vector[:k]]
        if any(len(p) != eng.range_proof_bytes(1) for p in proofs):
            raise ValueError("Bulletproof verification failed")
        st = eng.verify_range_proofs(b"".join(a.data[96:128] for a in epsilon_account[:k]), b"".join(proofs), 1, k, 64,
                                     transcript_label, verifier_label, b"AggregateBulletProof", transcript)
        if st[0]:
            raise ValueError("Bulletproof verification failed")
        return None

    @staticmethod
    def zero_balance_account_vector_verifier(anonymity_accounts, z, x, transcript_label=b"ZeroBalanceAccount",
                                             verifier_label=b"DLOGProof"):
        """src/accounts/verifier.rs:593-634 (domain separator spelled as the reference's verifier spells it)."""
        assert len(anonymity_accounts) == len(z)
        st = default_engine().verify_zero_balance(b"".join(a.data for a in anonymity_accounts),
                                                  b"".join(bytes(s) for s in z), bytes(x), len(z), True, transcript_label,
                                                  verifier_label)
        if st[0] == B.ST_BAD_POINT:
            raise ValueError("Zero balance Account Verify: Failed")
        if st[0]:
            raise ValueError("Zero balance account verification failed")
        return None

    @staticmethod
    def zero_balance_account_verifier(account, z, x, transcript_label=b"ZeroBalanceAccount", verifier_label=b"DLOGProof"):
        """src/accounts/verifier.rs:647-680."""
        st = default_engine().verify_zero_balance(account.data, bytes(z), bytes(x), 1, False, transcript_label, verifier_label)
        if st[0] == B.ST_BAD_POINT:
            raise ValueError("Zero balance Account Verify: Failed")
        if st[0]:
            raise ValueError("Zero balance account verification failed")
        return None

    @staticmethod
    def destroy_account_verifier(accounts, z, x, transcript_label=b"DestroyAccount", verifier_label=b"DLOGProof"):
        """src/accounts/verifier.rs:693-735."""
        assert len(accounts) == len(z)
        st = default_engine().verify_destroy_account(b"".join(a.data for a in accounts), b"".join(bytes(s) for s in z),
                                                     bytes(x), len(z), transcript_label, verifier_label)
        if st[0] == B.ST_BAD_POINT:
            raise ValueError("Destroy Account Verify: Failed")
        if st[0]:
            raise ValueError("Destroy account verification failed")
        return None

    @staticmethod
    def verify_same_value_compact_verifier(enc_account, commitment, proof):
        """src/accounts/verifier.rs:747-806.  proof = SigmaProof::Dleq as the tuple (zv[], zr[], _, x) of get_dleq()."""
        zv, zr, _, x = proof
        st = default_engine().verify_same_value_compact(enc_account.data, bytes(commitment), bytes(zv[0]), bytes(zr[0]),
                                                        bytes(x))
        if st[0] == B.ST_BAD_POINT:
            raise ValueError("Delta Compact Proof Verify: Failed")
        if st[0]:
            raise ValueError("Same Value Proof Verify: Failed")
        return None

    @staticmethod
    def verify_update_account_dark_tx_verifier(delta_updated_accounts, output_accounts, z_vector, x,
                                               transcript_label=b"UpdateAccount", verifier_label=b"DLOGProof"):
        """src/accounts/verifier.rs:818-917."""
        if len(delta_updated_accounts) != len(output_accounts):
            raise ValueError("Length of delta_updated_accounts and output_accounts is not same")
        st = default_engine().verify_update_account_dark_tx(b"".join(a.data for a in delta_updated_accounts),
                                                            b"".join(a.data for a in output_accounts),
                                                            bytes(z_vector[0]) + bytes(z_vector[1]), bytes(x),
                                                            len(output_accounts), transcript_label, verifier_label)
        if st[0] == B.ST_PANIC:
            raise PanicError("called `Option::unwrap()` on a `None` value")
        if st[0] == B.ST_BAD_POINT:
            raise ValueError("Update Account: DLOG Proof Verify: Failed")
        if st[0]:
            raise ValueError("Update Output Challenge : DLOG Proof Verify: Failed")
        return None


# ---- leaf arguments of the shuffle proof (src/shuffle/{ddh,singlevalueproduct,hadamard}.rs) ---------------------------------
# The reference's verify methods receive a `Verifier` carrying the running Merlin transcript; here the two labels that
# built it are keyword arguments (defaults: the labels of the reference's own unit tests).
class DDHProof:
    def __init__(self, challenge, z):
        self.challenge, self.z = _b(challenge, 32), _b(z, 32)

    def verify_ddh_proof(self, statement, G, H, transcript_label=b"ShuffleProof", verifier_label=b"DDHTuple"):
        """src/shuffle/ddh.rs:109-142; statement = (G_dash, H_dash).  Returns None or raises ValueError."""
        st = default_engine().verify_ddh(bytes(G), bytes(H), bytes(statement[0]), bytes(statement[1]), self.challenge, self.z,
                                         transcript_label, verifier_label)
        if st[0]:
            raise ValueError("DDH Proof Verify: Failed")
        return None


class SVPProof:
    FIELDS = ("commitment_d", "commitment_delta_small", "commitment_delta_capital", "a_twildle", "b_twildle", "r_twildle",
              "s_twildle")

    def __init__(self, **kw):
        for f in self.FIELDS:
            setattr(self, f, kw[f])

    def verify(self, svparg, transcript_label=b"SingleValue", verifier_label=b"Shuffle"):
        """src/shuffle/singlevalueproduct.rs:175-257; svparg = (commitment_a, b)."""
        if len(self.a_twildle) != 3 or len(self.b_twildle) != 3:
            raise ValueError("SingleValue Product Proof Verify: Size check failed")
        blob = (bytes(self.commitment_d) + bytes(self.commitment_delta_small) + bytes(self.commitment_delta_capital) +
                b"".join(bytes(s) for s in self.a_twildle) + b"".join(bytes(s) for s in self.b_twildle) +
                bytes(self.r_twildle) + bytes(self.s_twildle))
        st = default_engine().verify_svp(bytes(svparg[0]), bytes(svparg[1]), blob, transcript_label, verifier_label)
        if st[0] == B.ST_BAD_POINT:
            raise ValueError("SingleValue Product Proof Verify: Decompression Failed")
        if st[0]:
            raise ValueError("SingleValue Product Proof Verify: Failed")
        return None


class HadamardProof:
    FIELDS = ("commitment_a_0", "commitment_b_0", "commitment_c_0", "commitment_delta", "a_bar", "b_bar", "c_bar", "r_bar",
              "s_bar", "t_bar", "rho_bar")
    MESSAGES = {1: "Hadamard Proof Verify: Omega values are not unique",
                2: "Hadamard Proof Verify: A_bar , B_bar, C_bar check failed",
                3: "Hadamard Proof Verify: Delta Commitment check failed"}

    def __init__(self, **kw):
        for f in self.FIELDS:
            setattr(self, f, kw[f])

    def verify(self, hstatement, commit_a, commit_b, commit_c, transcript_label=b"Hadamard", verifier_label=b"Shuffle"):
        """src/shuffle/hadamard.rs:249-389; hstatement = omega (3 scalars)."""
        j = lambda v: b"".join(bytes(s) for s in v)  # noqa: E731
        blob = (bytes(self.commitment_a_0) + bytes(self.commitment_b_0) + bytes(self.commitment_c_0) + j(self.commitment_delta) +
                j(self.a_bar) + j(self.b_bar) + j(self.c_bar) + bytes(self.r_bar) + bytes(self.s_bar) + bytes(self.t_bar) +
                bytes(self.rho_bar))
        st, det = default_engine().verify_hadamard(j(hstatement), j(commit_a), j(commit_b), j(commit_c), blob, transcript_label,
                                                   verifier_label)
        if st[0] == B.ST_BAD_POINT:
            raise ValueError("HadamardProof Verify: Decompression Failed")
        if st[0]:
            raise ValueError(self.MESSAGES.get(int(det[0]), "Hadamard Proof Verify: failed"))
        return None
