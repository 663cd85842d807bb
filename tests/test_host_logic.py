"""Host-side logic of the product that needs no GPU: transcript, scalar field helpers, path verification."""
import hashlib
import os
import random

import numpy as np

from oracle import pyref as P

M = P.M


def test_transcript_matches_reference_semantics(ml, oracle, golden):
    t = ml.Transcript()
    assert t.next_challenge() == int(golden["challenge_empty"])
    assert t.next_challenge() == int(golden["challenge_empty"])  # idempotent until the next absorb (transcript.rs:35-38)
    ot, pt = oracle.transcript(), P.Transcript()
    rng = random.Random(11)
    for n in (0, 1, 16, 32, 55, 56, 63, 64, 65, 100, 200):
        b = bytes(rng.randrange(256) for _ in range(n))
        t.absorb(b), ot.absorb(b), pt.absorb(b)
        assert t.random() == ot.random() == pt.random()
        assert t.next_challenge() == ot.next_challenge() == pt.next_challenge()
    c = t.clone()
    c.absorb(b"fork")
    assert c.random() != t.random() and t.random() == pt.random()


def test_pow2_generator(ml, golden):
    for k, g in golden["pow_2_generator"].items():
        assert ml.pow_2_generator(int(k)) == int(g)
    assert ml.pow_2_generator(41) is None
    assert ml.pow_2_generator_powers(41) is None


def test_fingerprint_and_delta(ml, oracle):
    rng = random.Random(12)
    r = rng.randrange(M)
    cs = [rng.randrange(M) for _ in range(9)]
    assert ml.fingerprint(r, cs) == P.fingerprint(r, cs)
    a, b = [rng.randrange(M) for _ in range(7)], [rng.randrange(M) for _ in range(7)]
    assert ml.delta_evaluate(a, b) == P.delta_evaluate(a, b)


def test_path_verify_against_oracle_tree(ml, oracle):
    data = np.frombuffer(os.urandom(16 * 5), dtype=np.uint8).reshape(16, 5)
    m = oracle.merkle_commit(data)
    for idx in (0, 5, 15):
        value, path = m.open(idx)
        assert ml.path_verify(value, path, m.root(), idx) == 0
        assert ml.path_verify(value, path, m.root(), idx ^ 1) == 105   # IncompatibleIndex
        assert ml.path_verify(value + b"x", path, m.root(), idx) == 104  # IncompatibleHash


def test_top_from_roots(ml):
    from multilinear_b200 import load
    import ctypes as C
    roots = [hashlib.sha256(bytes([i])).digest() for i in range(8)]
    layer = roots
    while len(layer) > 1:
        layer = [hashlib.sha256(layer[i] + layer[i + 1]).digest() for i in range(0, len(layer), 2)]
    buf = np.frombuffer(b"".join(roots), dtype=np.uint8)
    out = np.empty(32, dtype=np.uint8)
    assert load().ml_merkle_top_from_roots(C.c_void_p(buf.ctypes.data), C.c_size_t(8), C.c_void_p(out.ctypes.data)) == 0
    assert out.tobytes() == layer[0]
    assert load().ml_merkle_top_from_roots(C.c_void_p(buf.ctypes.data), C.c_size_t(6), C.c_void_p(out.ctypes.data)) == 1


def test_compute_entry_points_fail_loudly_without_gpu(ml):
    import torch
    if torch.cuda.is_available():
        return
    try:
        ml.ntt(list(range(8)), ml.pow_2_generator(3))
    except ml.MlError as e:
        assert e.code == 6  # ML_ERR_CUDA: no CPU fallback
    else:
        raise AssertionError("ntt ran without a GPU")


def test_host_verifier_accepts_oracle_proof_and_rejects_tampering(ml, oracle):
    """the C++ verifier behind the ABI (ml_fri_verify, src/fri/mod.rs:287-340) is host code: a proof made by the CPU oracle,
    handed over as its bincode bytes (ml_fri_proof_deserialize), must verify; tampered copies must fail with the FriProofError code"""
    from oracle.binding import fe_ints
    log_n = 6
    coeffs = oracle.synthetic(17, 1 << log_n)
    gp = oracle.pow2_generator_powers(log_n + 1)
    code = oracle.reed_solomon(coeffs, fe_ints(gp[1:2])[0])
    oproof, st = oracle.fri_prove(code, gp, oracle.transcript())
    assert st == 0
    blob = bytes(oproof.blob)
    p = ml.FriProof.deserialize(blob)
    assert p.verify() == 0 and p.serialize() == blob
    assert [c.hex() for c in p.commitments] == [blob[8 + 32 * i:8 + 32 * (i + 1)].hex() for i in range(log_n)]
    b = bytearray(blob); b[-1] ^= 1
    assert ml.FriProof.deserialize(b).verify() == 106                      # IncompatibleLastRandom
    q0 = 8 + 32 * log_n + 8
    b = bytearray(blob); b[q0 + 8 + 8] ^= 1                                 # first opened value of the first query
    assert ml.FriProof.deserialize(b).verify() == 104                      # IncompatibleHash
    b = bytearray(blob[:q0 - 8] + (5).to_bytes(8, "little") + blob[q0:])    # claims 5 queries
    try:
        r = ml.FriProof.deserialize(b).verify()
    except ml.MlError:
        r = "malformed"
    assert r in (102, "malformed")
    for bad in (blob[:-1], blob + b"\0", b""):
        try:
            ml.FriProof.deserialize(bad)
        except ml.MlError as e:
            assert e.code == 8
        else:
            raise AssertionError("malformed proof accepted")


def test_interpolate_is_host_code_and_matches_oracle(ml, oracle):
    """PolynomialEvals::interpolate (src/polynomials.rs:51-86) runs on host scalars: it must work without a GPU"""
    from oracle.binding import fe_ints
    for n in (1, 2, 3, 4, 9, 33):
        e = oracle.synthetic(800 + n, n)
        c = ml.PolynomialEvals(e).interpolate()
        assert np.array_equal(c.coeffs, oracle.interpolate(e)), n
        xs = fe_ints(c.coeffs)
        for i in (0, n - 1):
            assert P.poly_eval(xs, i) == fe_ints(e)[i]


def test_sharded_prover_argument_checks_without_gpu(ml):
    import torch
    for world, n_polys, n_vars in ((3, 6, 8), (2, 3, 8), (4, 4, 1), (32, 32, 8)):   # not a power of two / uneven split / < 2 rows / too many ranks
        try:
            ml.ShardedBatchedProver.single_process([0] * world, n_polys, n_vars)
        except ml.MlError as e:
            assert e.code in (2, 8), (world, e.code)
        else:
            raise AssertionError("bad shape accepted")
    if not torch.cuda.is_available():
        try:
            ml.ShardedBatchedProver.single_process([0, 0], 4, 6)
        except ml.MlError as e:
            assert e.code in (6, 7)  # no device: fails loudly, no CPU fallback
        else:
            raise AssertionError("sharded prover created without a GPU")
