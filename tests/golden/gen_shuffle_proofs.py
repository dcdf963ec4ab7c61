"""Generates tests/golden/shuffle_proofs.bin: valid shuffle proofs made by the oracle's restatement of the reference prover
(Shuffle::input_shuffle + ShuffleProof::create_shuffle_proof, oracle/shuffle_ref.py) on the reference's shuffle_proof_test
scenario, serialised in the C ABI's layout (include/qq_b200.h, qq_verify_shuffle_batch).  Record = shuffle_input (9 x 128 B) |
shuffle_output (9 x 128 B) | statement (352 B) | proof (3776 B) = 6432 B.  Used by tests and by bench.py / tools as a
ready-made workload (the product never imports the oracle).  Run: python tests/golden/gen_shuffle_proofs.py [count]"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

RECORD = 9 * 128 * 2 + 352 + 3776


def main():
    import __graft_entry__ as g
    g.load_package()
    import shuffle_ref as F
    import test_gpu_parity as T
    from qq_testlib import Stream, scenario_shuffle
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    st = Stream(b"shuffle-golden")
    xpc = F.XpcGens(4)
    out = b""
    for i in range(count):
        inp, outp, proof, state = scenario_shuffle(st)
        assert F.shuffle_verify(F.new_transcript(b"ShuffleProof", b"Shuffle"), proof, state, inp, outp, xpc) == (True, None)
        pr, stm = T._shuffle_blobs(proof, state)
        out += b"".join(inp) + b"".join(outp) + stm + pr
    assert len(out) == count * RECORD
    with open(os.path.join(HERE, "shuffle_proofs.bin"), "wb") as f:
        f.write(out)
    print("wrote", count, "proofs,", len(out), "bytes")


if __name__ == "__main__":
    main()
