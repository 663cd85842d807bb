//! Pinning kit for the CPU oracle (oracle/oracle.c, oracle/pyref.py) — to be run by a maintainer who has `cargo`.
//!
//! The graft image has no Rust toolchain, so every golden value under tests/golden/ descends from the two restatements of the
//! reference ("parity unpinned").  This file turns that into a check against the REAL crate:
//!
//!   1. copy this file to  <fr34za/multilinear checkout>/src/golden_dump.rs
//!   2. add to src/lib.rs:   #[cfg(test)] mod golden_dump;
//!   3. cargo test --release golden_dump -- --nocapture | grep '^GOLDEN ' | sed 's/^GOLDEN //' > rust_golden.json
//!   4. python tests/golden/check_against_rust.py rust_golden.json        (in this repository)
//!
//! It prints ONE line `GOLDEN {json}` with exactly the schema of tests/golden/vectors.json (generated there by
//! tests/golden/gen_golden.py) from the crate's own functions on the crate's own test inputs:
//!   prove_and_verify_test (src/fri/mod.rs:350-363), multilinear_pcs_bench_test scaled to n_vars = 8
//!   (src/fri/multilinear_pcs.rs:211-228), batched_fri_benchmark (src/fri/batched_fri.rs:441-479), batched_pcs_verify_test scaled
//!   to n_vars = 6 (src/fri/batched_pcs.rs:262-306), the three Merkle tests (src/merkle_tree/mod.rs:301-438), ntt([0..8)),
//!   multilinear_conversion_test (src/polynomials.rs:207-214), From<i64>, pow_2_generator, the bincode blob of fri_benchmark_test's
//!   configuration (src/fri/mod.rs:367-369).
//! BatchedFriProof does not derive Serialize in the crate; its blob is assembled below from the bincode encodings of its
//! (serialisable) parts in field order — the layout `#[derive(Serialize)]` would produce, and what ml_bfri_proof_serialize writes.
use sha2::{Digest, Sha256};

use crate::field::Field128;
use crate::fri::batched_fri::BatchedFriProof;
use crate::fri::batched_pcs::{BatchedPCSClaim, BatchedPCSProof};
use crate::fri::multilinear_pcs::PCSProof;
use crate::fri::{reed_solomon, FriProof, LOG_BLOWUP};
use crate::merkle_tree::Merkle;
use crate::ntt::{NttField, Polynomial};
use crate::polynomials::MultilinearPolynomialEvals;
use crate::transcript::Transcript;

type F = Field128;

fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}
fn sha_hex(b: &[u8]) -> String {
    hex(&Sha256::digest(b))
}
fn dec(x: F) -> String {
    // canonical value as a decimal string (the 16 little-endian bytes AsRef exposes, src/field.rs:33-38)
    let mut le = [0u8; 16];
    le.copy_from_slice(x.as_ref());
    format!("\"{}\"", u128::from_le_bytes(le))
}
fn dec_list(xs: &[F]) -> String {
    format!("[{}]", xs.iter().map(|x| dec(*x)).collect::<Vec<_>>().join(", "))
}
fn bytes_of(xs: &[F]) -> Vec<u8> {
    xs.iter().flat_map(|x| x.as_ref().to_vec()).collect()
}
fn cfg() -> impl bincode::config::Config {
    bincode::config::standard().with_little_endian().with_fixed_int_encoding() // src/fri/mod.rs:367-369
}
fn enc<T: serde::Serialize>(v: &T) -> Vec<u8> {
    bincode::serde::encode_to_vec(v, cfg()).expect("serialization failed")
}

#[test]
fn golden_dump() {
    let mut kv: Vec<(String, String)> = Vec::new();
    let mut put = |k: &str, v: String| kv.push((k.to_string(), v));

    put("modulus", format!("\"{}\"", F::modulus()));
    let gens: Vec<String> = [1u64, 2, 3, 10, 21, 25, 40]
        .iter()
        .map(|&k| format!("\"{}\": {}", k, dec(F::pow_2_generator(k).unwrap())))
        .collect();
    put("pow_2_generator", format!("{{{}}}", gens.join(", ")));
    let fi: Vec<String> = [-1i64, -7, 0, 1, 1 << 40, i64::MIN]
        .iter()
        .map(|&x| format!("\"{}\": {}", x, dec(F::from(x))))
        .collect();
    put("from_i64", format!("{{{}}}", fi.join(", ")));
    put("half", dec(F::from(1) / F::from(2)));
    put("challenge_empty", dec(Transcript::new().next_challenge::<F>()));
    let c8: Vec<F> = (0..8).map(|i| F::from(i as i64)).collect();
    put("ntt8", dec_list(&Polynomial { coeffs: c8 }.ntt(F::pow_2_generator(3).unwrap()).evals));

    // intt_test (src/ntt/mod.rs:192-201) scaled to 2^10
    let coeffs: Vec<F> = (0..1 << 10).map(|i| F::from(i as i64)).collect();
    let ev = Polynomial { coeffs: coeffs.clone() }.ntt(F::pow_2_generator(10).unwrap());
    assert_eq!(ev.intt().coeffs, coeffs);
    put("ntt_1024_sha", format!("\"{}\"", sha_hex(&bytes_of(&ev.evals))));
    let rs = reed_solomon(coeffs, F::pow_2_generator(11).unwrap());
    put("rs_1024_sha", format!("\"{}\"", sha_hex(&bytes_of(&rs))));

    // Merkle tests (src/merkle_tree/mod.rs:301-438)
    let d0: Vec<[u8; 1]> = vec![[0], [8], [4], [1], [5], [7], [6], [1]];
    let d1: Vec<[u8; 1]> = vec![[1], [3], [2], [3], [2], [1], [2], [3]];
    put("merkle_test_root", format!("\"{}\"", hex(&Merkle::commit(d0.clone()).root())));
    put("batched_merkle_test_root", format!("\"{}\"", hex(&Merkle::batch_commit(vec![d0, d1]).root())));
    let vecs: Vec<Vec<[u8; 2]>> = vec![
        vec![[0, 4], [8, 2], [4, 9], [1, 3], [5, 7], [7, 2], [6, 8], [1, 5]],
        vec![[9, 3], [2, 7], [6, 1], [3, 8], [4, 2], [8, 5], [1, 9], [7, 4]],
        vec![[3, 6], [5, 1], [8, 3], [2, 9], [7, 5], [1, 8], [4, 3], [6, 2]],
        vec![[7, 1], [3, 9], [5, 2], [8, 6], [1, 4], [9, 7], [2, 5], [4, 8]],
    ];
    put("batched_merkle_with_vectors_test_root", format!("\"{}\"", hex(&Merkle::batch_commit(vecs).root())));

    // multilinear_conversion_test (src/polynomials.rs:207-214): length 6, not a power of two
    let e6 = MultilinearPolynomialEvals { evals: [0, 1, 4, 8, 9, 3].iter().map(|&x| F::from(x as i64)).collect() };
    put("mle_conv6", dec_list(&e6.to_coefficient().coeffs));

    // prove_and_verify_test (src/fri/mod.rs:350-363)
    {
        let log_n = 10;
        let values: Vec<F> = (0..1 << log_n).map(|i| F::from(i as i64 * 7 + 3)).collect();
        let gen_pows = F::pow_2_generator_powers((log_n + LOG_BLOWUP) as u64).unwrap();
        let code = reed_solomon(values, gen_pows[1]);
        let mut t = Transcript::new();
        let proof = FriProof::prove(&code, &gen_pows, &mut t);
        proof.verify().unwrap();
        let blob = enc(&proof);
        let mut t0 = Transcript::new();
        t0.absorb(&proof.commitments[0]);
        let r0: F = t0.next_challenge();
        let comms: Vec<String> = proof.commitments.iter().map(|c| format!("\"{}\"", hex(c))).collect();
        put(
            "fri_log10",
            format!(
                "{{\"commitments\": [{}], \"last_elem\": {}, \"last_random\": \"{}\", \"blob_len\": {}, \"blob_sha\": \"{}\", \"r0\": {}}}",
                comms.join(", "), dec(proof.last_elem), hex(&proof.last_random), blob.len(), sha_hex(&blob), dec(r0)
            ),
        );
    }

    // multilinear_pcs_bench_test (src/fri/multilinear_pcs.rs:211-228) scaled to n_vars = 8
    {
        let n_vars = 8;
        let evals: Vec<F> = (0..1 << n_vars).map(|i| F::from(i as i64 * 7 + 3)).collect();
        let multilinear = MultilinearPolynomialEvals { evals };
        let inputs: Vec<F> = (0..n_vars).map(|i| F::from(i as i64)).collect();
        let output = multilinear.evaluate(&inputs);
        let mut t = Transcript::new();
        let proof = PCSProof::prove(inputs, output, multilinear, &mut t);
        proof.verify(&mut Transcript::new()).unwrap();
        // the round challenges, replayed as the verifier does (multilinear_pcs.rs:150-166)
        let mut vt = Transcript::new();
        let mut challenges = Vec::new();
        for (c, sp) in proof.fri_proof.commitments.iter().zip(proof.sumcheck_polynomials.iter()) {
            vt.absorb(c);
            for x in sp.nonzero_coeffs.iter() {
                vt.absorb(x.as_ref());
            }
            challenges.push(vt.next_challenge::<F>());
        }
        let sc: Vec<String> = proof.sumcheck_polynomials.iter().map(|sp| dec_list(&sp.nonzero_coeffs)).collect();
        put(
            "pcs_nv8",
            format!(
                "{{\"output\": {}, \"root0\": \"{}\", \"last_elem\": {}, \"sumcheck\": [{}], \"challenges\": {}, \"blob_sha\": \"{}\", \"final_random\": \"{}\"}}",
                dec(output), hex(&proof.fri_proof.commitments[0]), dec(proof.fri_proof.last_elem), sc.join(", "), dec_list(&challenges),
                sha_hex(&enc(&proof.fri_proof)), hex(&t.random())
            ),
        );
    }

    // batched_fri_benchmark (src/fri/batched_fri.rs:441-479)
    let bfri_blob = |p: &BatchedFriProof<F>| -> Vec<u8> {
        let mut b = Vec::new();
        b.extend_from_slice(&p.batch_commitment);
        b.extend(enc(&p.commitments));
        b.extend((p.queries.len() as u64).to_le_bytes());
        for q in p.queries.iter() {
            b.extend(enc(&q.batch_path));
            b.extend(enc(&q.query_proof));
        }
        b.extend(enc(&p.last_elem));
        b.extend_from_slice(&p.last_random);
        b
    };
    {
        let log_n = 6;
        let gen_pows = F::pow_2_generator_powers((log_n + LOG_BLOWUP) as u64).unwrap();
        let codes: Vec<Vec<F>> = (0..4i64)
            .map(|j| reed_solomon((0..1 << log_n).map(|i| F::from((i as i64 * 7 + 3) + j * 100)).collect(), gen_pows[1]))
            .collect();
        let mut t = Transcript::new();
        let proof = BatchedFriProof::prove(&codes, &gen_pows, &mut t);
        proof.verify().unwrap();
        let comms: Vec<String> = proof.commitments.iter().map(|c| format!("\"{}\"", hex(c))).collect();
        put(
            "bfri_log6_b4",
            format!(
                "{{\"batch_commitment\": \"{}\", \"commitments\": [{}], \"last_elem\": {}, \"blob_sha\": \"{}\"}}",
                hex(&proof.batch_commitment), comms.join(", "), dec(proof.last_elem), sha_hex(&bfri_blob(&proof))
            ),
        );
    }

    // batched_pcs_verify_test (src/fri/batched_pcs.rs:262-306) scaled to n_vars = 6, 10 polynomials
    {
        let n_vars = 6;
        let height = 1 << n_vars;
        let num_polys = 10;
        let inputs: Vec<F> = (0..n_vars).map(|i| F::from(i as i64)).collect();
        let mut polys = Vec::new();
        let mut outputs = Vec::new();
        for i in 0..num_polys {
            let evals: Vec<F> = (0..height).map(|j| F::from(((j as u64 * 3 + i as u64 * 5) % 100) as u128)).collect();
            let m = MultilinearPolynomialEvals { evals };
            outputs.push(m.evaluate(&inputs));
            polys.push(m);
        }
        let outs = dec_list(&outputs);
        let claim = BatchedPCSClaim { inputs, outputs };
        let mut t = Transcript::new();
        let proof = BatchedPCSProof::prove(claim, &polys, &mut t);
        proof.verify(&mut Transcript::new()).unwrap();
        let sc: Vec<String> = proof.sumcheck_polynomials.iter().map(|sp| dec_list(&sp.nonzero_coeffs)).collect();
        put(
            "bpcs_nv6_b10",
            format!(
                "{{\"outputs\": {}, \"batch_commitment\": \"{}\", \"last_elem\": {}, \"sumcheck\": [{}], \"blob_sha\": \"{}\"}}",
                outs, hex(&proof.fri_proof.batch_commitment), dec(proof.fri_proof.last_elem), sc.join(", "), sha_hex(&bfri_blob(&proof.fri_proof))
            ),
        );
    }

    let body: Vec<String> = kv.iter().map(|(k, v)| format!("\"{}\": {}", k, v)).collect();
    println!("GOLDEN {{{}}}", body.join(", "));
}
