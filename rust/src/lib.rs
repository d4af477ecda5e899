//! The reference crate's public surface (`/root/reference/src/lib.rs:1-4`) over the B200 engine.
//!
//! ```text
//! scalar::Scalar                                   src/scalar.rs
//! curves::{G1Point, G2Point, bilinear_map}         src/curves.rs
//! trusted_setup::{SetupArtifactsGenerator, SetupArtifact}   src/trusted_setup.rs
//! polynomial::{Polynomial, Evaluation}             src/polynomial.rs
//! ```
//! Values are kept in their canonical encodings (Fr: 32 bytes LE; G1: 48 bytes compressed; G2: 96 bytes
//! compressed) — exactly what the reference's serde impls emit, so JSON artifacts are byte-identical.
//! Bulk operations (`Polynomial::commit`, `evaluate`, `Evaluation::generate_proof`, the setup iterator) are one
//! FFI call each; per-element methods go to the host helpers of the same library.  [`engine`] owns the
//! process-wide context: ALL visible GPUs by default, driven from this one process (`kzg_ctx_create_multi`).
pub mod curves;
pub mod engine;
pub mod ffi;
pub mod polynomial;
pub mod scalar;
pub mod trusted_setup;

#[cfg(test)]
mod tests {
    //! The reference's acceptance test (`src/lib.rs:16-94`): commit -> evaluate -> prove -> pairing check,
    //! plus the known-answer vectors of its deterministic bench inputs.
    use crate::{polynomial::Polynomial, scalar::Scalar, trusted_setup::{SetupArtifact, SetupArtifactsGenerator}};

    fn bench_secret() -> [u8; 32] {
        let mut s = [0u8; 32];
        for (i, b) in s.iter_mut().enumerate() { *b = i as u8; }
        s
    }

    fn run_kate_proof_test(polynomial: &Polynomial, input_point: Scalar, setup_artifacts: &[SetupArtifact]) {
        let commitment = polynomial.commit(setup_artifacts).unwrap();
        let evaluation = polynomial.evaluate(input_point).unwrap();
        let proof = evaluation.generate_proof(polynomial, setup_artifacts).unwrap();
        assert!(evaluation.verify_proof(&proof, &commitment, setup_artifacts).unwrap());
        let mut tampered = evaluation.clone();
        tampered.result = tampered.result.add(&Scalar::from_i128(1));
        assert!(!tampered.verify_proof(&proof, &commitment, setup_artifacts).unwrap());
    }

    #[test]
    fn kate_proof_small_degrees() {
        let setup: Vec<SetupArtifact> = SetupArtifactsGenerator::new(bench_secret()).take(12).collect();
        for coefficients in [vec![5i128, 10, 0, -3], vec![0, 0, 0], vec![7], vec![1, 2, 3, 4, 5, 6, 7, 8, 9, 10]] {
            run_kate_proof_test(&Polynomial::try_from(coefficients).unwrap(), Scalar::from_i128(21), &setup);
        }
    }

    #[test]
    fn known_answer_degree_1() {
        // c_i = 5^i + 10, z = 5^d + 20 (benches/polynomial_commitment.rs:10-23, benches/evaluation_proof.rs:25-27)
        let setup: Vec<SetupArtifact> = SetupArtifactsGenerator::new(bench_secret()).take(2).collect();
        let p = Polynomial::try_from(vec![11i128, 15]).unwrap();
        let c = serde_json::to_string(&p.commit(&setup).unwrap()).unwrap();
        let want = hex::decode("add892451b1c924699eaadfa1566680ac35623081bd4c71db9aa61019f80283e7ab2f173a9d68b11d9defeb998b988bd").unwrap();
        assert_eq!(c, serde_json::to_string(&want.iter().map(|b| *b as u32).collect::<Vec<_>>()).unwrap());
        let e = p.evaluate(Scalar::from_i128(25)).unwrap();
        assert_eq!(e.result, Scalar::from_i128(386));
        assert!(e.generate_proof(&p, &setup).is_ok());
        let wrong = crate::polynomial::Evaluation { point: Scalar::from_i128(25), result: Scalar::from_i128(387) };
        assert_eq!(wrong.generate_proof(&p, &setup).unwrap_err().to_string(),
                   "[divide_by_root] Fail to divide the polynomial by a root, constant terms do not add up");
    }
}
