// The reference's own unit tests (hot-path subset), re-expressed against the C++ mirror of its interface
// (include/multilinear_b200.hpp).  Each test cites the Rust test it follows; they are round-trip / prove->verify
// checks exactly like the originals, plus the golden roots pinned in tests/golden/vectors.json.
//   build: g++ -std=c++17 -Iinclude tests/cpp/reference_tests.cpp -Lmultilinear_b200 -lmultilinear_b200
#include <cstdio>
#include <string>

#include "multilinear_b200.hpp"

using namespace ml;

static int failures = 0;
#define EXPECT(cond)                                                            \
    do {                                                                        \
        if (!(cond)) { printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

static std::string hex(const HashDigest& d) {
    static const char* k = "0123456789abcdef";
    std::string s;
    for (uint8_t b : d) { s += k[b >> 4]; s += k[b & 15]; }
    return s;
}
static std::vector<std::vector<uint8_t>> bytes1(std::initializer_list<int> v) {
    std::vector<std::vector<uint8_t>> out;
    for (int x : v) out.push_back({(uint8_t)x});
    return out;
}

// src/ntt/mod.rs:192-201
static void intt_test() {
    const int log_n = 18;
    std::vector<F> coeffs;
    for (long long i = 0; i < (1 << log_n); i++) coeffs.push_back(F::from(i));
    Polynomial pol{coeffs};
    F gen = *pow_2_generator(log_n);
    LagrangePolynomial ntt = pol.ntt(gen);
    Polynomial intt = ntt.intt();
    EXPECT(pol == intt);
}
// src/merkle_tree/mod.rs:301-309
static void merkle_test() {
    Merkle tree = Merkle::commit(bytes1({0, 8, 4, 1, 5, 7, 6, 1}));
    auto proof = tree.open(5);
    EXPECT(proof.has_value());
    EXPECT(proof->verify(tree.root(), 5) == 0);
    EXPECT(hex(tree.root()) == "5363442da24fb12a0ccc1648a93f5ffcebfc51f486b4690d52b2c237a8540573");
    EXPECT(!tree.open(8).has_value());
    bool panicked = false;
    try { Merkle::commit(bytes1({1, 2, 3})); } catch (const Panic& p) { panicked = p.code == ML_ERR_NOT_POW2; }
    EXPECT(panicked);  // assert!(data.len().is_power_of_two())
}
// src/merkle_tree/mod.rs:312-351
static void batched_merkle_test() {
    Merkle tree = Merkle::batch_commit({bytes1({0, 8, 4, 1, 5, 7, 6, 1}), bytes1({1, 3, 2, 3, 2, 1, 2, 3})});
    auto proof = tree.batch_open(5);
    EXPECT(proof->value.size() == 2 && proof->value[0] == 7 && proof->value[1] == 1);
    EXPECT(proof->verify(tree.root(), 5) == 0);
    proof = tree.batch_open(2);
    EXPECT(proof->value[0] == 4 && proof->value[1] == 2);
    EXPECT(proof->verify(tree.root(), 2) == 0);
    EXPECT(proof->verify(tree.root(), 1) != 0);  // incorrect index fails
    EXPECT(hex(tree.root()) == "1abbfb63f271dd7335a2763a6ec1d931745c6a44f9dd659f01081f848e74f44a");
}
// src/polynomials.rs:207-214 (length 6: only the first 2^trailing_zeros entries are transformed)
static void multilinear_conversion_test() {
    MultilinearPolynomialEvals evals{{F::from(0), F::from(1), F::from(4), F::from(8), F::from(9), F::from(3)}};
    MultilinearPolynomial pol = evals.to_coefficient();
    EXPECT(evals == pol.to_evaluation());
}
// src/fri/mod.rs:350-363
static void prove_and_verify_test() {
    const int log_n = 10;
    std::vector<F> values;
    for (long long i = 0; i < (1 << log_n); i++) values.push_back(F::from(i * 7 + 3));
    std::vector<F> gen_pows = *pow_2_generator_powers(log_n + LOG_BLOWUP);
    std::vector<F> code = reed_solomon(values, gen_pows[1]);
    Transcript transcript;
    FriProof proof = FriProof::prove(code, gen_pows, transcript);
    EXPECT(proof.verify() == 0);
    EXPECT(proof.commitments().size() == 10);
    EXPECT(hex(proof.commitments()[0]) == "f6ea9e052de7f712e802f82ff7ec0f818d484dc56a6e1bc43305acb4026cadf0");
}
// src/fri/multilinear_pcs.rs:211-228 — BASELINE config 1
static void multilinear_pcs_bench_test() {
    const int n_vars = 20;
    std::vector<F> evals, inputs;
    for (long long i = 0; i < (1 << n_vars); i++) evals.push_back(F::from(i * 7 + 3));
    MultilinearPolynomialEvals multilinear{evals};
    for (long long i = 0; i < n_vars; i++) inputs.push_back(F::from(i));
    F output = multilinear.evaluate(inputs);
    Transcript transcript;
    PCSProof proof = PCSProof::prove(inputs, output, multilinear, transcript);
    Transcript vt;
    EXPECT(proof.verify(vt) == 0);
    EXPECT(proof.sumcheck_polynomials().size() == (size_t)n_vars);
}
// src/fri/batched_pcs.rs:262-306 (n_vars reduced from 20 to 14 to keep the host-side setup short)
static void batched_pcs_verify_test() {
    const int n_vars = 14, num_polys = 10;
    const size_t height = (size_t)1 << n_vars;
    std::vector<F> inputs;
    for (long long i = 0; i < n_vars; i++) inputs.push_back(F::from(i));
    std::vector<MultilinearPolynomialEvals> polys;
    std::vector<F> outputs;
    for (int i = 0; i < num_polys; i++) {
        std::vector<F> evals;
        for (size_t j = 0; j < height; j++) evals.push_back(F::from((long long)((j * 3 + (size_t)i * 5) % 100)));
        MultilinearPolynomialEvals m{evals};
        outputs.push_back(m.evaluate(inputs));
        polys.push_back(std::move(m));
    }
    BatchedPCSClaim claim{inputs, outputs};
    Transcript transcript;
    BatchedPCSProof proof = BatchedPCSProof::prove(claim, polys, transcript);
    Transcript vt;
    EXPECT(proof.verify(vt) == 0);
    // the same claim through the sharded prover (two virtual ranks on device 0): identical proof bytes, identical transcript
    ShardedBatchedProver sharded({0, 0}, num_polys, n_vars);
    Transcript st;
    BatchedPCSProof sproof = sharded.prove(claim, polys, st);
    Transcript svt;
    EXPECT(sproof.verify(svt) == 0);
    EXPECT(sproof.fri_proof_bytes() == proof.fri_proof_bytes());
    EXPECT(st.random() == transcript.random());
}
// src/polynomials.rs:197-204
static void interpolation_test() {
    UnivariatePolynomialEvals evals{{F::from(0), F::from(1), F::from(4), F::from(8), F::from(9), F::from(3)}};
    UnivariatePolynomial pol = evals.interpolate();
    EXPECT(evals == pol.evaluate_over_domain());
}

// ---- field helpers for the verifier-side equations (one-element calls into the library; the host mirror has no arithmetic)
static F fadd(F a, F b) { F o; check(ml_fe_add_vec(a.bytes(), b.bytes(), 1, o.bytes())); return o; }
static F fsub(F a, F b) { F o; check(ml_fe_sub_vec(a.bytes(), b.bytes(), 1, o.bytes())); return o; }
static F fmul(F a, F b) { F o; check(ml_fe_mul_vec(a.bytes(), b.bytes(), 1, o.bytes())); return o; }
static F finv(F a) { F o; check(ml_fe_inv_vec(a.bytes(), 1, o.bytes())); return o; }
static F poly_eval(const std::vector<F>& c, F x) {
    F acc = F::from(0);
    for (size_t i = c.size(); i-- > 0;) acc = fadd(fmul(acc, x), c[i]);
    return acc;
}
// SumcheckPolynomial::to_polynomial (sumcheck.rs:269-276)
static std::vector<F> to_polynomial(const SumcheckPolynomial& p, F sum) {
    F s = F::from(0);
    for (F c : p.nonzero_coeffs) s = fadd(s, c);
    std::vector<F> coeffs{fmul(fsub(sum, s), finv(F::from(2)))};
    coeffs.insert(coeffs.end(), p.nonzero_coeffs.begin(), p.nonzero_coeffs.end());
    return coeffs;
}
// src/constraint_system/sumcheck.rs:342-365 — sumcheck_test: pythagorean trace (16 x 4), two constraints, sum 0,
// checked with verify_sumcheck_debug (:55-90)
static void sumcheck_test() {
    const long long rows[64] = {3, 4, 5, 7, 5, 12, 13, 17, 8, 15, 17, 23, 7, 24, 25, 31, 20, 21, 29, 41, 12, 35, 37, 47, 9, 40, 41, 49, 28, 45, 53, 73,
                                11, 60, 61, 71, 16, 63, 65, 79, 33, 56, 65, 89, 48, 55, 73, 103, 13, 84, 85, 97, 36, 77, 85, 113, 39, 80, 89, 119,
                                65, 72, 97, 137};
    const size_t width = 4, height = 16, n_vars = 4;
    std::vector<F> matrix;
    for (long long v : rows) matrix.push_back(F::from(v));
    Transcript transcript;
    // System::prover -> ChallengeSet::new (system.rs:132-147): the transcript is not mutated, so every challenge is the same value
    F c = transcript.next_challenge();
    std::vector<F> row_point(n_vars, c);
    // constraint_mask (system.rs:91-93) over one constraint variable: [1 - c, c]
    F m0 = fsub(F::from(1), c), m1 = c, zero = F::from(0);
    // pythagorean_set (:333-339): mask0 * (x0^2 + x1^2 - x2^2) + mask1 * (x0 + x1 - x3), degree 2
    std::vector<CompositionTerm> terms = {{m0, {0, 0}}, {m0, {1, 1}}, {fsub(zero, m0), {2, 2}}, {m1, {0}}, {m1, {1}}, {fsub(zero, m1), {3}}};
    auto comp = [&](const std::vector<F>& x) {
        F acc = F::from(0);
        for (auto& t : terms) {
            F p = t.coef;
            for (uint32_t j : t.cols) p = fmul(p, x[j]);
            acc = fadd(acc, p);
        }
        return acc;
    };
    Transcript verifier_transcript(transcript);
    SumcheckTables tables = SumcheckTables::build(row_point, matrix, width);
    tables.set_composition(terms);
    F sum = F::from(0);
    auto [pols, randoms] = tables.compute_sumcheck_polynomials(2, transcript, sum);
    EXPECT(pols.size() == n_vars && pols[0].nonzero_coeffs.size() == 3 && tables.height() == 1);
    // verify_sumcheck_debug
    std::vector<F> rs;
    for (F cf : pols[0].nonzero_coeffs) verifier_transcript.absorb(cf.bytes(), 16);
    std::vector<F> pol = to_polynomial(pols[0], sum);
    for (size_t k = 1; k < pols.size(); k++) {
        F r = verifier_transcript.next_challenge();
        for (F cf : pols[k].nonzero_coeffs) verifier_transcript.absorb(cf.bytes(), 16);
        pol = to_polynomial(pols[k], poly_eval(pol, r));
        rs.push_back(r);
    }
    F r = verifier_transcript.next_challenge();
    rs.push_back(r);
    EXPECT(rs == randoms);
    std::vector<F> output(width);  // Trace::evaluate (evaluation.rs:33-48): the multilinear extension of every column at rs
    for (size_t j = 0; j < width; j++) {
        std::vector<F> col(height);
        for (size_t i = 0; i < height; i++) col[i] = matrix[i * width + j];
        output[j] = MultilinearPolynomialEvals{col}.evaluate(rs);
    }
    F delta;
    check(ml_delta_evaluate(raw(row_point), raw(rs), n_vars, delta.bytes()));
    EXPECT(fmul(delta, comp(output)) == poly_eval(pol, r));  // "Does not match polynomial evaluation"
}

int main() {
    struct { const char* name; void (*fn)(); } tests[] = {
        {"intt_test", intt_test}, {"merkle_test", merkle_test}, {"batched_merkle_test", batched_merkle_test},
        {"multilinear_conversion_test", multilinear_conversion_test}, {"interpolation_test", interpolation_test},
        {"prove_and_verify_test", prove_and_verify_test},
        {"multilinear_pcs_bench_test", multilinear_pcs_bench_test}, {"batched_pcs_verify_test", batched_pcs_verify_test},
        {"sumcheck_test", sumcheck_test}};
    for (auto& t : tests) {
        int before = failures;
        try {
            t.fn();
        } catch (const Panic& p) {
            printf("  PANIC in %s: %s\n", t.name, p.what());
            failures++;
        }
        printf("test %s ... %s\n", t.name, failures == before ? "ok" : "FAILED");
    }
    printf("%s\n", failures ? "FAILED" : "all reference tests passed");
    return failures ? 1 : 0;
}
