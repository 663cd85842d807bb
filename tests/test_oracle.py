"""The CPU oracle against the golden vectors and the independent Python restatement (no GPU)."""
import hashlib
import os
import random

import numpy as np
import pytest

from oracle import pyref as P
from oracle.binding import fe_arr, fe_ints

M = P.M


def sha(b):
    return hashlib.sha256(b).hexdigest()


def test_field_vectors(oracle, golden):
    assert str(M) == golden["modulus"]
    for k, g in golden["pow_2_generator"].items():
        assert oracle.pow2_generator(int(k)) == int(g) == P.pow_2_generator(int(k))
    assert oracle.pow2_generator(41) is None and P.pow_2_generator(41) is None
    for x, v in golden["from_i64"].items():
        assert oracle.from_i64(int(x)) == int(v)
    assert oracle.binop("div", 1, 2) == int(golden["half"])
    assert oracle.binop("div", 5, 0) == 0  # winter-math: inv(0) = 0
    assert oracle.transcript().next_challenge() == int(golden["challenge_empty"])


def test_field_random_vs_python(oracle):
    rng = random.Random(1)
    edge = [0, 1, 2, M - 1, M - 2, 2**64 - 1, 2**64, 2**64 + 1, 2**127, 2**96 - 1, 45 * 2**40 - 1]
    xs = [rng.randrange(M) for _ in range(3000)] + edge
    ys = [rng.randrange(M) for _ in range(3000)] + edge[::-1]
    a, b = fe_arr(xs), fe_arr(ys)
    assert fe_ints(oracle.vec("mul", a, b)) == [x * y % M for x, y in zip(xs, ys)]
    assert fe_ints(oracle.vec("add", a, b)) == [(x + y) % M for x, y in zip(xs, ys)]
    assert fe_ints(oracle.vec("sub", a, b)) == [(x - y) % M for x, y in zip(xs, ys)]
    for x in edge:
        for y in edge:
            assert oracle.binop("mul", x, y) == x * y % M


def test_sha256_kat(oracle):
    # FIPS 180-4 known answers + hashlib cross-check on both code paths (SHA-NI and portable)
    kat = {b"abc": "ba7816bf8f01cfea414140de5dae2223b00361a396177a9cb410ff61f20015ad",
           b"": "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855",
           b"abcdbcdecdefdefgefghfghighijhijkijkljklmklmnlmnomnopnopq": "248d6a61d20638b8e5c026930c3e6039a33ce45964ff2167f6ecedd419db06c1"}
    for force in (0, 1):
        oracle.L.or_sha256_force_portable(force)
        for msg, dg in kat.items():
            assert oracle.sha256(msg).hex() == dg
        for n in [1, 55, 56, 63, 64, 65, 119, 120, 128, 1000, 2048]:
            d = os.urandom(n)
            assert oracle.sha256(d) == hashlib.sha256(d).digest()
    oracle.L.or_sha256_force_portable(0)


def test_ntt_golden(oracle, golden):
    e = oracle.ntt(fe_arr(range(8)), oracle.pow2_generator(3))
    assert [str(x) for x in fe_ints(e)] == golden["ntt8"]
    coeffs = fe_arr([P.from_i64(i) for i in range(1 << 10)])
    g = oracle.pow2_generator(10)
    ev = oracle.ntt(coeffs, g)
    assert sha(ev.tobytes()) == golden["ntt_1024_sha"]
    assert np.array_equal(oracle.intt(ev, g), coeffs)  # intt_test, src/ntt/mod.rs:192-201
    rs = oracle.reed_solomon(coeffs, oracle.pow2_generator(11))
    assert sha(rs.tobytes()) == golden["rs_1024_sha"]


def test_ntt_matches_pyref_and_threads(oracle):
    rng = random.Random(2)
    for log_n in (1, 2, 5, 9):
        x = [rng.randrange(M) for _ in range(1 << log_n)]
        g = P.pow_2_generator(log_n)
        for th in (1, 4):
            oracle.set_threads(th)
            assert fe_ints(oracle.ntt(fe_arr(x), g)) == P.ntt(x, g)
            assert fe_ints(oracle.reed_solomon(fe_arr(x), P.pow_2_generator(log_n + 1))) == P.reed_solomon(x, P.pow_2_generator(log_n + 1))
    # definition check: X[k] = sum_j x[j] gen^(jk)
    x = [rng.randrange(M) for _ in range(16)]
    g = P.pow_2_generator(4)
    assert fe_ints(oracle.ntt(fe_arr(x), g)) == [sum(x[j] * pow(g, j * k, M) for j in range(16)) % M for k in range(16)]


def test_merkle_golden(oracle, golden):
    d0 = np.array([[0], [8], [4], [1], [5], [7], [6], [1]], dtype=np.uint8)
    d1 = np.array([[1], [3], [2], [3], [2], [1], [2], [3]], dtype=np.uint8)
    m = oracle.merkle_commit(d0)
    assert m.root().hex() == golden["merkle_test_root"]
    value, path = m.open(5)
    assert value == bytes([7]) and len(path) == 3 and P.path_verify(value, path, m.root(), 5)
    assert m.open(8) is None
    assert oracle.merkle_commit(d0[:6]) is None  # not a power of two: the reference panics
    mb = oracle.merkle_batch_commit([d0, d1])
    assert mb.root().hex() == golden["batched_merkle_test_root"]
    assert mb.open(5)[0] == bytes([7, 1]) and mb.open(2)[0] == bytes([4, 2])
    value, path = mb.open(2)
    assert P.path_verify(value, path, mb.root(), 2) and not P.path_verify(value, path, mb.root(), 1)


def test_mle(oracle, golden):
    e6 = fe_arr([0, 1, 4, 8, 9, 3])
    c = oracle.to_coefficient(e6)
    assert [str(x) for x in fe_ints(c)] == golden["mle_conv6"]
    assert np.array_equal(oracle.to_evaluation(c), e6)  # multilinear_conversion_test, polynomials.rs:207-214
    rng = random.Random(3)
    ev = [rng.randrange(M) for _ in range(64)]
    args = [rng.randrange(M) for _ in range(6)]
    assert oracle.mle_evals_evaluate(fe_arr(ev), fe_arr(args)) == P.mle_evals_evaluate(ev, args)
    co = P.to_coefficient(ev)
    assert fe_ints(oracle.to_coefficient(fe_arr(ev))) == co
    assert oracle.mle_coeffs_evaluate(fe_arr(co), fe_arr(args)) == P.mle_coeffs_evaluate(co, args) == P.mle_evals_evaluate(ev, args)


def test_fri_prove_golden(oracle, golden):
    g = golden["fri_log10"]
    log_n = 10
    vals = fe_arr([P.from_i64(7 * i + 3) for i in range(1 << log_n)])
    gp = oracle.pow2_generator_powers(log_n + 1)
    code = oracle.reed_solomon(vals, fe_ints(gp[1:2])[0])
    t = oracle.transcript()
    proof, st = oracle.fri_prove(code, gp, t)
    assert st == 0
    assert [c.hex() for c in proof.commitments] == g["commitments"]
    assert str(proof.last_elem) == g["last_elem"] and proof.last_random.hex() == g["last_random"]
    assert len(proof.blob) == g["blob_len"] and sha(proof.blob) == g["blob_sha"]
    assert proof.verify() == 0  # prove_and_verify_test, src/fri/mod.rs:350-363


def test_fri_rejects_non_rs_code(oracle):
    rng = random.Random(5)
    gp = oracle.pow2_generator_powers(6)
    code = fe_arr([rng.randrange(M) for _ in range(64)])  # not a low-degree codeword
    f, st = oracle.fri_fold(gp, code, oracle.transcript())
    assert f is None and st == 4


def test_sumcheck_matches_pyref(oracle):
    rng = random.Random(7)
    nv = 5
    ev = [rng.randrange(M) for _ in range(1 << nv)]
    inp = [rng.randrange(M) for _ in range(nv)]
    s = oracle.sumcheck_build(fe_arr(inp), fe_arr(ev))
    ps = P.SumcheckTables(inp, ev)
    m, d = s.tables()
    assert fe_ints(m) == ps.matrix and fe_ints(d) == ps.delta
    for r in (1, 2, 3, rng.randrange(M)):
        assert s.partial_sum(r) == ps.partial_sum(r)
    claim = P.mle_evals_evaluate(ev, inp)
    assert (ps.partial_sum(1) + ps.partial_sum(0)) % M == claim  # sum_x f(x) eq(inputs, x) = f(inputs)
    t, pt = oracle.transcript(), P.Transcript()
    coeffs, rs = s.compute_sumcheck_polynomials(1, t, claim)
    prev, pc, pr = claim, [], []
    for _ in range(nv):
        nz, r, prev = ps.compute_sumcheck_polynomial(2, prev, pt)
        pc += nz
        pr.append(r)
    assert coeffs == pc and rs == pr and t.random() == pt.random()


def test_pcs_golden(oracle, golden):
    g = golden["pcs_nv8"]
    nv = 8
    evals = fe_arr([P.from_i64(7 * i + 3) for i in range(1 << nv)])
    inputs = fe_arr([P.from_i64(i) for i in range(nv)])
    out = oracle.mle_evals_evaluate(evals, inputs)
    assert str(out) == g["output"]
    t = oracle.transcript()
    proof, st = oracle.pcs_prove(inputs, out, evals, t)
    assert st == 0
    assert proof.fri.commitments[0].hex() == g["root0"] and str(proof.fri.last_elem) == g["last_elem"]
    assert [str(c) for c in proof.sumcheck] == [c for nz in g["sumcheck"] for c in nz]
    assert sha(proof.fri.blob) == g["blob_sha"] and t.random().hex() == g["final_random"]
    assert proof.verify(oracle.transcript()) == 0  # multilinear_pcs_bench_test, multilinear_pcs.rs:211-228
    bad = oracle.transcript()
    bad.absorb(b"x")
    assert proof.verify(bad) != 0


def test_batched_golden(oracle, golden):
    g = golden["bfri_log6_b4"]
    log_n, B = 6, 4
    gp = oracle.pow2_generator_powers(log_n + 1)
    gen = fe_ints(gp[1:2])[0]
    codes = [oracle.reed_solomon(fe_arr([P.from_i64(7 * i + 3 + 100 * j) for i in range(1 << log_n)]), gen) for j in range(B)]
    proof, st = oracle.batched_fri_prove(codes, gp, oracle.transcript())
    assert st == 0 and proof.batch_commitment.hex() == g["batch_commitment"]
    assert [c.hex() for c in proof.commitments] == g["commitments"] and str(proof.last_elem) == g["last_elem"]
    assert sha(proof.blob) == g["blob_sha"] and proof.verify() == 0

    g = golden["bpcs_nv6_b10"]
    nv, B = 6, 10
    polys = [fe_arr([(j * 3 + i * 5) % 100 for j in range(1 << nv)]) for i in range(B)]
    inputs = fe_arr([P.from_i64(i) for i in range(nv)])
    outputs = [oracle.mle_evals_evaluate(p, inputs) for p in polys]
    assert [str(o) for o in outputs] == g["outputs"]
    proof, st = oracle.batched_pcs_prove(inputs, fe_arr(outputs), polys, oracle.transcript())
    assert st == 0 and proof.fri.batch_commitment.hex() == g["batch_commitment"]
    assert [str(c) for c in proof.sumcheck] == [c for nz in g["sumcheck"] for c in nz]
    assert sha(proof.fri.blob) == g["blob_sha"]
    assert proof.verify(oracle.transcript()) == 0  # batched_pcs_verify_test, batched_pcs.rs:262-306


def test_synthetic_generator(oracle):
    a = oracle.synthetic(0xB200, 1000)
    assert np.array_equal(a, oracle.synthetic(0xB200, 1000)) and all(x < M for x in fe_ints(a))
    assert not np.array_equal(a, oracle.synthetic(0xB201, 1000))


# ------------------------------------------------------------------ width-w sumcheck tables (System path)
PYTHAGOREAN = [3, 4, 5, 7, 5, 12, 13, 17, 8, 15, 17, 23, 7, 24, 25, 31, 20, 21, 29, 41, 12, 35, 37, 47, 9, 40, 41, 49, 28, 45, 53, 73,
               11, 60, 61, 71, 16, 63, 65, 79, 33, 56, 65, 89, 48, 55, 73, 103, 13, 84, 85, 97, 36, 77, 85, 113, 39, 80, 89, 119,
               65, 72, 97, 137]  # src/constraint_system/sumcheck.rs:302-325


def pythagorean_terms(m0, m1):
    """System::evaluate_composition for pythagorean_set (:333-339): mask0 * (x0^2 + x1^2 - x2^2) + mask1 * (x0 + x1 - x3)"""
    return [(m0, [0, 0]), (m0, [1, 1]), ((-m0) % M, [2, 2]), (m1, [0]), (m1, [1]), ((-m1) % M, [3])]


def eval_terms(terms, x):
    acc = 0
    for c, cols in terms:
        for j in cols:
            c = c * x[j] % M
        acc = (acc + c) % M
    return acc


def verify_sumcheck_debug(oracle, t, pols, total_degree, s, matrix, width, row_point, terms):
    """System::verify_sumcheck_debug (sumcheck.rs:55-90) restated on the oracle's primitives"""
    def to_poly(nz, sm):  # SumcheckPolynomial::to_polynomial (:269-276)
        return [(sm - sum(nz)) * pow(2, -1, M) % M] + list(nz)

    def ev(p, x):
        a = 0
        for c in reversed(p):
            a = (a * x + c) % M
        return a
    rounds = [pols[total_degree * k:total_degree * (k + 1)] for k in range(len(pols) // total_degree)]
    for c in rounds[0]:
        t.absorb(int(c).to_bytes(16, "little"))
    pol, rs = to_poly(rounds[0], s), []
    for p in rounds[1:]:
        r = t.next_challenge()
        for c in p:
            t.absorb(int(c).to_bytes(16, "little"))
        pol = to_poly(p, ev(pol, r))
        rs.append(r)
    r = t.next_challenge()
    rs.append(r)
    out = oracle.trace_evaluate(matrix, width, fe_arr(rs))
    delta = oracle.delta_evaluate(row_point, fe_arr(rs))
    assert delta * eval_terms(terms, out) % M == ev(pol, r), "Does not match polynomial evaluation"
    return rs


def pythagorean_system(oracle, log_height=4):
    """the reference's sumcheck_test / sumcheck_high_bench set-up (:342-398): trace, ChallengeSet, constraint mask"""
    rows = list(PYTHAGOREAN)
    while len(rows) < 4 << log_height:
        rows = rows + rows
    matrix = fe_arr(rows)
    t = oracle.transcript()
    c = t.next_challenge()  # ChallengeSet::new (system.rs:132-147): the transcript is not mutated, every challenge is c
    row_point, cons = fe_arr([c] * log_height), fe_arr([c])
    m0, m1 = oracle.mask_evaluate(0, cons), oracle.mask_evaluate(1, cons)  # system.rs:91-93
    assert m0 == (1 - c) % M and m1 == c
    return matrix, row_point, pythagorean_terms(m0, m1), t


@pytest.mark.parametrize("log_height", [4, 7])
def test_wide_sumcheck_reference_sumcheck_test(oracle, log_height):
    matrix, row_point, terms, t = pythagorean_system(oracle, log_height)
    s = oracle.wsumcheck_build(row_point, matrix, 4)
    s.set_composition(terms)
    m, d = s.tables()
    assert np.array_equal(m, matrix)
    assert fe_ints(d)[5] == oracle.mask_evaluate(5, row_point)
    vt = oracle.transcript()
    pols, rs = s.compute_sumcheck_polynomials(2, t, 0)  # constraints.degree() = 2 (:338), sum = 0 (:350)
    assert len(pols) == 3 * log_height and s.height() == 1
    assert verify_sumcheck_debug(oracle, vt, pols, 3, 0, matrix, 4, row_point, terms) == rs


def test_wide_sumcheck_width1_equals_pcs_tables(oracle):
    """width 1 with the composition x[0] is exactly the PCS specialisation (multilinear_pcs.rs:56-57)"""
    nv = 6
    evals = oracle.synthetic(77, 1 << nv)
    inputs = fe_arr([3 * i + 2 for i in range(nv)])
    claim = oracle.mle_evals_evaluate(evals, inputs)
    a = oracle.sumcheck_build(inputs, evals)
    b = oracle.wsumcheck_build(inputs, evals, 1)
    b.set_composition([(1, [0])])
    assert a.partial_sum(2) == b.partial_sum(2) and a.partial_sum(1) == b.partial_sum(1)
    assert a.compute_sumcheck_polynomials(1, oracle.transcript(), claim) == b.compute_sumcheck_polynomials(1, oracle.transcript(), claim)


def test_wide_sumcheck_matches_python_restatement(oracle):
    """independent big-int restatement of partial_sum / fold (sumcheck.rs:204-247) for a width-3 trace, degree-3 composition"""
    rng = random.Random(4)
    nv, w = 4, 3
    h = 1 << nv
    mat = [rng.randrange(M) for _ in range(h * w)]
    pts = [rng.randrange(M) for _ in range(nv)]
    terms = [(rng.randrange(M), [0, 1, 2]), (rng.randrange(M), [2, 2]), (rng.randrange(M), []), (M - 1, [1])]
    delta = []
    for idx in range(h):  # Mask::evaluate (evaluation.rs:56-73), big-endian variable order
        prod = 1
        for i in range(nv):
            p = pts[nv - 1 - i]
            prod = prod * (p if (idx >> i) & 1 else (1 - p)) % M
        delta.append(prod)
    s = oracle.wsumcheck_build(fe_arr(pts), fe_arr(mat), w)
    s.set_composition(terms)
    assert fe_ints(s.tables()[1]) == delta
    height = h
    for r in (1, 2, 3, rng.randrange(M)):
        off = height >> 1
        want = 0
        for i in range(off):
            if r == 1:
                d = delta[i + off]
                row = [mat[(i + off) * w + j] for j in range(w)]
            else:
                d = ((1 - r) * delta[i] + r * delta[i + off]) % M
                row = [((1 - r) * mat[i * w + j] + r * mat[(i + off) * w + j]) % M for j in range(w)]
            want = (want + eval_terms(terms, row) * d) % M
        assert s.partial_sum(r) == want
        # fold with the same r
        s.fold(r)
        for i in range(off):
            delta[i] = ((1 - r) * delta[i] + r * delta[i + off]) % M
            for j in range(w):
                mat[i * w + j] = ((1 - r) * mat[i * w + j] + r * mat[(i + off) * w + j]) % M
        height = off
        m_now, d_now = s.tables()
        assert fe_ints(d_now) == delta[:height] and fe_ints(m_now) == mat[:height * w]
