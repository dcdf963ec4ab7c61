"""Generates tests/golden/range_proofs_m{1,4,16}.bin: valid Bulletproofs range proofs made by the oracle's prover restatement
(oracle/rangeproof_ref.py) the way the reference's prover makes them (Prover::verify_non_negative_sender_receiver_prover,
src/accounts/prover.rs:544-590): transcript Transcript::new(b"SenderAccountProof"), Prover::new(b"BulletProof"),
domain_sep(b"AggregateBulletProof"), 64-bit values, PedersenGens::default(), BulletproofGens::new(64, 16).
Record = m commitments (32 B each) | proof ((9 + 2 lg(64 m)) x 32 B), the C ABI's layout (include/qq_b200.h,
qq_verify_range_proof_batch, chain = 1).  Ready-made workload for bench.py / tools (the product never imports the oracle).
Run: python tests/golden/gen_range_proofs.py [count]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))


def records(m, count):
    import rangeproof_ref as RP
    from merlin_ref import Transcript
    from qq_testlib import Stream
    st = Stream(b"range-golden-%d" % m)
    out = b""
    for _ in range(count):
        vals = [int.from_bytes(st.bytes(8), "little") for _ in range(m)]
        bl = [st.scalar() for _ in range(m)]

        def tr():
            t = Transcript(b"SenderAccountProof")
            t.domain_sep(b"BulletProof")
            t.domain_sep(b"AggregateBulletProof")
            return t
        proof, V = RP.prove_multiple(tr(), vals, bl, 64, st.scalar, RP.BulletproofGens(64, 16))
        assert RP.verify_multiple(tr(), proof, V, 64, bp_gens=RP.BulletproofGens(64, 16)) is True
        out += b"".join(V) + proof
    return out


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    for m in (1, 4, 16):
        data = records(m, count)
        with open(os.path.join(HERE, "range_proofs_m%d.bin" % m), "wb") as f:
            f.write(data)
        print("m =", m, ":", count, "proofs,", len(data), "bytes")


if __name__ == "__main__":
    main()
