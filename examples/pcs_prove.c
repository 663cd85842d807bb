/* Plain-C caller of the drop-in boundary (include/multilinear_b200.h): the reference's multilinear_pcs_bench_test
 * (src/fri/multilinear_pcs.rs:211-228) — evals 7i+3, inputs 0..n_vars, prove, verify — with host buffers only.
 *   gcc -std=c11 -Iinclude examples/pcs_prove.c -Lmultilinear_b200 -lmultilinear_b200 -Wl,-rpath,'$ORIGIN/../multilinear_b200' -o examples/pcs_prove
 *   ./examples/pcs_prove [n_vars]          (needs a B200; there is no CPU fallback: every call fails with ML_ERR_CUDA without one) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "multilinear_b200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int st_ = (call);                                                                \
        if (st_ != ML_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, st_, ml_last_error()); return 1; } \
    } while (0)

int main(int argc, char **argv) {
    const size_t n_vars = argc > 1 ? (size_t)atoi(argv[1]) : 20, n = (size_t)1 << n_vars;
    int64_t *iv = (int64_t *)malloc(n * sizeof(int64_t));
    uint8_t *evals = (uint8_t *)aligned_alloc(16, 16 * n), *inputs = (uint8_t *)aligned_alloc(16, 16 * n_vars), output[16];
    for (size_t i = 0; i < n; i++) iv[i] = (int64_t)(7 * i + 3);
    CHECK(ml_fe_from_i64_vec(iv, n, evals));                       /* F::from(i * 7 + 3) */
    for (size_t i = 0; i < n_vars; i++) iv[i] = (int64_t)i;
    CHECK(ml_fe_from_i64_vec(iv, n_vars, inputs));                 /* F::from(i) */
    CHECK(ml_mle_evals_evaluate(evals, n, inputs, n_vars, output));

    ml_transcript *t = NULL, *vt = NULL;
    ml_pcs_proof *proof = NULL;
    CHECK(ml_transcript_new(&t));
    CHECK(ml_pcs_prove(inputs, n_vars, output, evals, n, t, &proof));          /* PCSProof::prove */
    CHECK(ml_transcript_new(&vt));
    const int verdict = ml_pcs_verify(proof, vt);                                /* PCSProof::verify */

    const ml_fri_proof *fri = ml_pcs_proof_fri(proof);
    const size_t nc = ml_fri_proof_num_commitments(fri);
    uint8_t *roots = (uint8_t *)malloc(32 * nc), last[16], last_random[32];
    CHECK(ml_fri_proof_commitments(fri, roots));
    CHECK(ml_fri_proof_last(fri, last, last_random));
    printf("n_vars %zu  commitments %zu  root_0 ", n_vars, nc);
    for (int i = 0; i < 32; i++) printf("%02x", roots[i]);
    printf("\nlast_elem (LE) ");
    for (int i = 0; i < 16; i++) printf("%02x", last[i]);
    printf("\nverify -> %d (%s)\n", verdict, verdict == 0 ? "accepted" : "REJECTED");

    ml_pcs_proof_free(proof);
    ml_transcript_free(t);
    ml_transcript_free(vt);
    free(roots); free(iv); free(evals); free(inputs);
    return verdict == 0 ? 0 : 1;
}
