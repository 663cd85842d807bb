"""Generates tests/golden/vectors.json from the pure-Python restatement (oracle/pyref.py).

The reference is Rust and cannot run here (no cargo/rustc; winter-math / sha2 not vendored), and none of its tests
pins a concrete value, so these vectors are DERIVED (parity unpinned): they freeze the agreed output of two
independent restatements (pyref.py and oracle.c) on the reference's own test inputs so that any later
drift in the oracle or the CUDA path is caught.  Run:  python tests/golden/gen_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pyref as P  # noqa: E402

M = P.M


def h(b):
    return hashlib.sha256(b).hexdigest()


def main():
    v = {"modulus": str(M)}
    v["pow_2_generator"] = {str(k): str(P.pow_2_generator(k)) for k in (1, 2, 3, 10, 21, 25, 40)}
    v["from_i64"] = {str(x): str(P.from_i64(x)) for x in (-1, -7, 0, 1, 2**40, -2**63)}
    v["half"] = str(P.div(1, 2))
    v["challenge_empty"] = str(P.Transcript().next_challenge())
    v["ntt8"] = [str(x) for x in P.ntt(list(range(8)), P.pow_2_generator(3))]

    # src/ntt/mod.rs:192-201 intt_test scaled to 2^10
    coeffs = [P.from_i64(i) for i in range(1 << 10)]
    g = P.pow_2_generator(10)
    ev = P.ntt(coeffs, g)
    assert P.intt(ev, g) == coeffs
    v["ntt_1024_sha"] = h(b"".join(P.to_bytes(x) for x in ev))
    rs = P.reed_solomon(coeffs, P.pow_2_generator(11))
    v["rs_1024_sha"] = h(b"".join(P.to_bytes(x) for x in rs))

    # src/merkle_tree/mod.rs:301-438 test data
    d0 = [bytes([x]) for x in [0, 8, 4, 1, 5, 7, 6, 1]]
    v["merkle_test_root"] = P.Merkle.commit(d0).root().hex()
    d1 = [bytes([x]) for x in [1, 3, 2, 3, 2, 1, 2, 3]]
    m = P.Merkle.batch_commit([d0, d1])
    v["batched_merkle_test_root"] = m.root().hex()
    assert m.open(5)[0] == bytes([7, 1]) and m.open(2)[0] == bytes([4, 2])
    vec = [[[0, 4], [8, 2], [4, 9], [1, 3], [5, 7], [7, 2], [6, 8], [1, 5]],
           [[9, 3], [2, 7], [6, 1], [3, 8], [4, 2], [8, 5], [1, 9], [7, 4]],
           [[3, 6], [5, 1], [8, 3], [2, 9], [7, 5], [1, 8], [4, 3], [6, 2]],
           [[7, 1], [3, 9], [5, 2], [8, 6], [1, 4], [9, 7], [2, 5], [4, 8]]]
    v["batched_merkle_with_vectors_test_root"] = P.Merkle.batch_commit([[bytes(x) for x in b] for b in vec]).root().hex()

    # src/polynomials.rs:207-214 (non power of two length)
    e6 = [0, 1, 4, 8, 9, 3]
    v["mle_conv6"] = [str(x) for x in P.to_coefficient(e6)]
    assert P.to_evaluation(P.to_coefficient(e6)) == e6

    # src/fri/mod.rs:350-363 prove_and_verify_test
    log_n = 10
    vals = [P.from_i64(7 * i + 3) for i in range(1 << log_n)]
    gp = P.pow_2_generator_powers(log_n + 1)
    code = P.reed_solomon(vals, gp[1])
    t = P.Transcript()
    proof = P.fri_prove(code, gp, t)
    assert P.fri_verify(proof)
    blob = P.fri_proof_serialize(proof)
    v["fri_log10"] = {"commitments": [c.hex() for c in proof["commitments"]], "last_elem": str(proof["last_elem"]),
                      "last_random": proof["last_random"].hex(), "blob_len": len(blob), "blob_sha": h(blob),
                      "r0": str(P.new(int.from_bytes(hashlib.sha256(proof["commitments"][0]).digest()[:16], "little")))}

    # src/fri/multilinear_pcs.rs:211-228 scaled to n_vars = 8
    nv = 8
    evals = [P.from_i64(7 * i + 3) for i in range(1 << nv)]
    inputs = [P.from_i64(i) for i in range(nv)]
    out = P.mle_evals_evaluate(evals, inputs)
    t = P.Transcript()
    pp = P.pcs_prove(inputs, out, evals, t)
    assert P.pcs_verify(pp, P.Transcript())
    blob = P.fri_proof_serialize(pp["fri"])
    v["pcs_nv8"] = {"output": str(out), "root0": pp["fri"]["commitments"][0].hex(), "last_elem": str(pp["fri"]["last_elem"]),
                    "sumcheck": [[str(c) for c in nz] for nz in pp["sumcheck"]], "challenges": [str(r) for r in pp["challenges"]],
                    "blob_sha": h(blob), "final_random": t.random().hex()}

    # src/fri/batched_fri.rs:441-479 batched_fri_benchmark
    log_n, B = 6, 4
    gp = P.pow_2_generator_powers(log_n + 1)
    codes = [P.reed_solomon([P.from_i64(7 * i + 3 + 100 * j) for i in range(1 << log_n)], gp[1]) for j in range(B)]
    t = P.Transcript()
    bp = P.batched_fri_prove(codes, gp, t)
    blob = P.bfri_proof_serialize(bp)
    v["bfri_log6_b4"] = {"batch_commitment": bp["batch_commitment"].hex(), "commitments": [c.hex() for c in bp["commitments"]],
                         "last_elem": str(bp["last_elem"]), "blob_sha": h(blob)}

    # src/fri/batched_pcs.rs:262-306 scaled to n_vars = 6, 10 polys
    nv, B = 6, 10
    polys = [[(j * 3 + i * 5) % 100 for j in range(1 << nv)] for i in range(B)]
    inputs = [P.from_i64(i) for i in range(nv)]
    outputs = [P.mle_evals_evaluate(p, inputs) for p in polys]
    t = P.Transcript()
    bpp = P.batched_pcs_prove(inputs, outputs, polys, t)
    blob = P.bfri_proof_serialize(bpp["fri"])
    v["bpcs_nv6_b10"] = {"outputs": [str(o) for o in outputs], "batch_commitment": bpp["fri"]["batch_commitment"].hex(),
                         "last_elem": str(bpp["fri"]["last_elem"]), "sumcheck": [[str(c) for c in nz] for nz in bpp["sumcheck"]],
                         "blob_sha": h(blob)}

    with open(os.path.join(HERE, "vectors.json"), "w") as f:
        json.dump(v, f, indent=1, sort_keys=True)
    print("wrote vectors.json")


if __name__ == "__main__":
    main()
