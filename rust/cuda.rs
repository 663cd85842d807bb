//! `src/cuda.rs` for fr34za/multilinear: extern "C" declarations of libmultilinear_b200.so and the thin wrappers
//! that keep the crate's existing signatures.  Source only — this image has no Rust toolchain;
//! the same symbols are exercised through ctypes (multilinear_b200/api.py) and C++ (include/multilinear_b200.hpp).
#![allow(dead_code)]
use std::os::raw::{c_char, c_int};

use crate::field::Field128;
use crate::fri::LOG_BLOWUP;

#[repr(C)] pub struct MlTranscript { _p: [u8; 0] }
#[repr(C)] pub struct MlMerkle { _p: [u8; 0] }
#[repr(C)] pub struct MlFri { _p: [u8; 0] }
#[repr(C)] pub struct MlFriProof { _p: [u8; 0] }
#[repr(C)] pub struct MlSumcheck { _p: [u8; 0] }
#[repr(C)] pub struct MlWSumcheck { _p: [u8; 0] }
#[repr(C)] pub struct MlPcsProof { _p: [u8; 0] }

const _: () = assert!(std::mem::size_of::<Field128>() == 16 && std::mem::align_of::<Field128>() == 16);
const _: () = assert!(cfg!(target_endian = "little"));

#[link(name = "multilinear_b200")]
extern "C" {
    pub fn ml_last_error() -> *const c_char;
    pub fn ml_pow2_generator(log_size: u64, out: *mut u8) -> c_int;
    pub fn ml_pow2_generator_powers(log_size: u64, out: *mut u8) -> c_int;
    pub fn ml_bit_reverse_permutation(values: *mut u8, n: usize, elem_bytes: usize) -> c_int;
    pub fn ml_ntt(coeffs: *const u8, n: usize, gen: *const u8, evals: *mut u8) -> c_int;
    pub fn ml_intt(evals: *const u8, n: usize, gen: *const u8, coeffs: *mut u8) -> c_int;
    pub fn ml_reed_solomon(coeffs: *const u8, n: usize, gen: *const u8, code: *mut u8) -> c_int;
    pub fn ml_mle_to_coefficient(evals: *const u8, len: usize, coeffs: *mut u8) -> c_int;
    pub fn ml_mle_to_evaluation(coeffs: *const u8, len: usize, evals: *mut u8) -> c_int;
    pub fn ml_mle_evals_evaluate(evals: *const u8, len: usize, args: *const u8, n_args: usize, out: *mut u8) -> c_int;
    pub fn ml_transcript_new(out: *mut *mut MlTranscript) -> c_int;
    pub fn ml_transcript_clone(t: *const MlTranscript, out: *mut *mut MlTranscript) -> c_int;
    pub fn ml_transcript_free(t: *mut MlTranscript);
    pub fn ml_transcript_absorb(t: *mut MlTranscript, bytes: *const u8, len: usize) -> c_int;
    pub fn ml_transcript_random(t: *const MlTranscript, out: *mut u8) -> c_int;
    pub fn ml_transcript_next_challenge(t: *mut MlTranscript, out: *mut u8) -> c_int;
    pub fn ml_merkle_commit(data: *const u8, item_bytes: usize, n_items: usize, out: *mut *mut MlMerkle) -> c_int;
    pub fn ml_merkle_batch_commit(data: *const *const u8, n_batches: usize, item_bytes: usize, n_items: usize, out: *mut *mut MlMerkle) -> c_int;
    pub fn ml_merkle_root(m: *const MlMerkle, out: *mut u8) -> c_int;
    pub fn ml_merkle_open(m: *const MlMerkle, index: usize, value: *mut u8, digests: *mut u8, dirs: *mut u8, path_len: *mut usize) -> c_int;
    pub fn ml_merkle_free(m: *mut MlMerkle);
    pub fn ml_fri_init(code: *const u8, n: usize, t: *mut MlTranscript, out: *mut *mut MlFri) -> c_int;
    pub fn ml_fri_fold_step(f: *mut MlFri, gen_pows: *const u8, gen_pows_len: usize, k: usize, r: *const u8, t: *mut MlTranscript) -> c_int;
    pub fn ml_fri_fold(gen_pows: *const u8, gen_pows_len: usize, code: *const u8, n: usize, t: *mut MlTranscript, out: *mut *mut MlFri) -> c_int;
    pub fn ml_fri_free(f: *mut MlFri);
    pub fn ml_fri_prove(code: *const u8, n: usize, gen_pows: *const u8, gen_pows_len: usize, t: *mut MlTranscript, out: *mut *mut MlFriProof) -> c_int;
    pub fn ml_fri_proof_serialized_len(p: *const MlFriProof) -> usize;
    pub fn ml_fri_proof_serialize(p: *const MlFriProof, out: *mut u8) -> c_int;
    pub fn ml_fri_proof_free(p: *mut MlFriProof);
    pub fn ml_sumcheck_build_tables_for_pcs(inputs: *const u8, n_vars: usize, evals: *const u8, height: usize, out: *mut *mut MlSumcheck) -> c_int;
    pub fn ml_sumcheck_compute_polynomial(s: *mut MlSumcheck, total_degree: usize, previous_sum: *mut u8, t: *mut MlTranscript, nonzero: *mut u8, r: *mut u8) -> c_int;
    pub fn ml_sumcheck_fold(s: *mut MlSumcheck, r: *const u8) -> c_int;
    pub fn ml_sumcheck_free(s: *mut MlSumcheck);
    // width-w tables (System path): the composition closure (sumcheck.rs:176) is passed as a sparse polynomial over the row
    pub fn ml_wsumcheck_build(row_point: *const u8, n_vars: usize, matrix: *const u8, width: usize, height: usize, out: *mut *mut MlWSumcheck) -> c_int;
    pub fn ml_wsumcheck_set_composition(w: *mut MlWSumcheck, n_terms: usize, coefs: *const u8, term_lens: *const u32, term_cols: *const u32) -> c_int;
    pub fn ml_wsumcheck_partial_sum(w: *mut MlWSumcheck, r: *const u8, out: *mut u8) -> c_int;
    pub fn ml_wsumcheck_fold(w: *mut MlWSumcheck, r: *const u8) -> c_int;
    pub fn ml_wsumcheck_compute_polynomials(w: *mut MlWSumcheck, composition_degree: usize, t: *mut MlTranscript, sum: *const u8, coeffs: *mut u8, randoms: *mut u8) -> c_int;
    pub fn ml_wsumcheck_free(w: *mut MlWSumcheck);
    pub fn ml_pcs_prove(inputs: *const u8, n_vars: usize, output: *const u8, evals: *const u8, n: usize, t: *mut MlTranscript, out: *mut *mut MlPcsProof) -> c_int;
    pub fn ml_pcs_proof_fri(p: *const MlPcsProof) -> *const MlFriProof;
    pub fn ml_pcs_proof_num_rounds(p: *const MlPcsProof) -> usize;
    pub fn ml_pcs_proof_sumcheck_coeffs(p: *const MlPcsProof, out: *mut u8) -> c_int;
    pub fn ml_pcs_proof_free(p: *mut MlPcsProof);
}

/// status -> the reference's behaviour: non-zero statuses 1, 2, 4 are its assert!/panic! sites
pub fn check(st: c_int) {
    if st != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(ml_last_error()) }.to_string_lossy().into_owned();
        panic!("{msg}");
    }
}

/// body of `reed_solomon` (src/fri/mod.rs:19-28)
pub fn reed_solomon(coeffs: Vec<Field128>, gen: Field128) -> Vec<Field128> {
    let n = coeffs.len();
    let mut code = vec![Field128::from(0); n << LOG_BLOWUP];
    check(unsafe { ml_reed_solomon(coeffs.as_ptr().cast(), n, gen.as_ref().as_ptr(), code.as_mut_ptr().cast()) });
    code
}

/// body of `Polynomial::ntt` (src/ntt/mod.rs:69-110)
pub fn ntt(coeffs: &[Field128], gen: Field128) -> Vec<Field128> {
    let mut evals = vec![Field128::from(0); coeffs.len()];
    check(unsafe { ml_ntt(coeffs.as_ptr().cast(), coeffs.len(), gen.as_ref().as_ptr(), evals.as_mut_ptr().cast()) });
    evals
}

/// body of `LagrangePolynomial::intt` (src/ntt/mod.rs:132-173)
pub fn intt(evals: &[Field128], gen: Field128) -> Vec<Field128> {
    let mut coeffs = vec![Field128::from(0); evals.len()];
    check(unsafe { ml_intt(evals.as_ptr().cast(), evals.len(), gen.as_ref().as_ptr(), coeffs.as_mut_ptr().cast()) });
    coeffs
}

/// `FriProof::prove` (src/fri/mod.rs:261-285): returns the bincode blob the unmodified serde derives decode
pub fn fri_prove_blob(code: &[Field128], gen_pows: &[Field128], transcript: *mut MlTranscript) -> Vec<u8> {
    let mut h = std::ptr::null_mut();
    check(unsafe { ml_fri_prove(code.as_ptr().cast(), code.len(), gen_pows.as_ptr().cast(), gen_pows.len(), transcript, &mut h) });
    let mut blob = vec![0u8; unsafe { ml_fri_proof_serialized_len(h) }];
    check(unsafe { ml_fri_proof_serialize(h, blob.as_mut_ptr()) });
    unsafe { ml_fri_proof_free(h) };
    blob
}
