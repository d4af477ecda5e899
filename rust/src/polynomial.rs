//! `Polynomial`, `Evaluation` (`/root/reference/src/polynomial.rs`): same constructors, trimming rules, error
//! messages and serde shape; the loops run on the engine.
//!
//! | reference (one blst call per term)                     | here (one FFI call)   |
//! |---|---|
//! | `commit`: n x (`blst_p1_mult` + add), `:200-215`       | `kzg_commit`          |
//! | `evaluate`: one bespoke `pow` per term, `:112-123`     | `kzg_evaluate`        |
//! | `divide_by_root`: sequential synthetic division, `:150-195` | `kzg_quotient`   |
//! | `generate_proof` = sub + divide_by_root + commit, `:260-269` | `kzg_open`      |
//! | `verify_proof`: two pairings, `:276-294`               | `kzg_verify_proof`    |
use crate::curves::{G1Point, G2Point};
use crate::engine::engine;
use crate::ffi::*;
use crate::scalar::Scalar;
use crate::trusted_setup::SetupArtifact;
use serde::{Deserialize, Serialize};

#[derive(Debug, Deserialize, Serialize)]
pub struct Polynomial {
    coefficients: Vec<Scalar>,
}

impl TryFrom<Vec<i128>> for Polynomial {
    type Error = anyhow::Error;

    /// coefficients in ascending degree: `vec![2, 1, 3]` is 2 + x + 3x^2
    fn try_from(value: Vec<i128>) -> Result<Self, Self::Error> {
        Self::try_from(value.into_iter().map(Scalar::from_i128).collect::<Vec<Scalar>>())
    }
}

impl TryFrom<Vec<Scalar>> for Polynomial {
    type Error = anyhow::Error;

    /// Trailing zero coefficients are dropped, but an all-zero input keeps ONE zero coefficient
    /// (`src/polynomial.rs:63-72`: the index of the last non-zero coefficient starts at 0).
    fn try_from(mut value: Vec<Scalar>) -> Result<Self, Self::Error> {
        if value.len() > u32::MAX as usize {
            return Err(anyhow::anyhow!(
                "Too many coefficients for polynomial, only 2**32 - 1 coefficients is supported. Got {}",
                value.len()
            ));
        }
        let keep = value.iter().rposition(|c| !c.is_zero()).unwrap_or(0) + 1;
        value.truncate(keep);
        Ok(Polynomial { coefficients: value })
    }
}

impl From<Scalar> for Polynomial {
    /// the constant polynomial; zero becomes the EMPTY polynomial (`src/polynomial.rs:83-87`)
    fn from(value: Scalar) -> Self {
        Polynomial { coefficients: if value.is_zero() { vec![] } else { vec![value] } }
    }
}

impl Polynomial {
    pub fn degree(&self) -> u32 {
        self.coefficients.len().saturating_sub(1) as u32
    }

    /// the coefficients as the engine takes them: n x 32 bytes LE
    fn le_bytes(&self) -> Vec<u8> {
        let mut out = Vec::with_capacity(self.coefficients.len() * 32);
        for c in &self.coefficients {
            out.extend_from_slice(&c.to_le_bytes());
        }
        out
    }

    fn from_le_bytes(bytes: &[u8]) -> Vec<Scalar> {
        bytes.chunks_exact(32).map(|c| Scalar::from_le_bytes(c.try_into().unwrap())).collect()
    }

    /// p(x) (`src/polynomial.rs:112-123`)
    pub fn evaluate(&self, x: Scalar) -> Result<Evaluation, anyhow::Error> {
        let coeffs = self.le_bytes();
        let mut y = [0u8; 32];
        let eng = engine();
        eng.check(unsafe {
            kzg_evaluate(eng.ctx(), coeffs.as_ptr(), self.coefficients.len() as u64, x.to_le_bytes().as_ptr(), y.as_mut_ptr())
        })?;
        Ok(Evaluation { point: x, result: Scalar::from_le_bytes(y) })
    }

    /// self - other, re-trimmed like every constructed polynomial (`src/polynomial.rs:128-145`)
    pub fn sub(&self, other: &Self) -> Result<Self, anyhow::Error> {
        let n = self.coefficients.len().max(other.coefficients.len());
        let zero = Scalar::default();
        let mut out = Vec::with_capacity(n);
        for i in 0..n {
            let a = self.coefficients.get(i).unwrap_or(&zero);
            let b = other.coefficients.get(i).unwrap_or(&zero);
            out.push(a.sub(b));
        }
        Polynomial::try_from(out)
    }

    /// self / (x - root) (`src/polynomial.rs:150-195`): empty quotient for the empty / zero polynomial, an error for a
    /// non-zero constant and for a remainder (`root` is not a root)
    pub fn divide_by_root(&self, root: &Scalar) -> Result<Self, anyhow::Error> {
        let n = self.coefficients.len();
        if n == 0 || (n == 1 && self.coefficients[0].is_zero()) {
            return Ok(Polynomial { coefficients: vec![] });
        }
        // only reachable through a polynomial deserialised WITH trailing zeros (the derive does not re-trim): the
        // reference then divides [c0, 0, ..] as a non-constant polynomial and trips over its constant-term check
        if n >= 2 && !self.coefficients[0].is_zero() && self.coefficients[1..].iter().all(|c| c.is_zero()) {
            return Err(anyhow::anyhow!(
                "[divide_by_root] Fail to divide the polynomial by a root, constant terms do not add up"
            ));
        }
        let coeffs = self.le_bytes();
        let zero = [0u8; 32];                      // the remainder the division must leave: p(root) = 0
        let mut q = vec![0u8; n * 32];
        let mut q_len = 0u64;
        let eng = engine();
        eng.check(unsafe {
            kzg_quotient(eng.ctx(), coeffs.as_ptr(), n as u64, root.to_le_bytes().as_ptr(), zero.as_ptr(), q.as_mut_ptr(), &mut q_len)
        })?;
        Polynomial::try_from(Self::from_le_bytes(&q[..q_len as usize * 32]))
    }

    /// C = sum_i c_i * setup[i].g1 (`src/polynomial.rs:200-215`); only the first `degree + 1` artifacts are used
    pub fn commit(&self, setup_artifacts: &[SetupArtifact]) -> Result<G1Point, anyhow::Error> {
        if (self.degree() + 1) as usize > setup_artifacts.len() {
            return Err(anyhow::anyhow!(
                "Setup does not allow for commitment generation of the polynomial. The polynomial degree is too high."
            ));
        }
        let coeffs = self.le_bytes();
        let mut out = [0u8; 48];
        let mut eng = engine();
        let setup = eng.resident(setup_artifacts)?;
        eng.check(unsafe { kzg_commit(eng.ctx(), setup, coeffs.as_ptr(), self.coefficients.len() as u64, out.as_mut_ptr()) })?;
        Ok(G1Point::from_engine(out))
    }

    /// Many polynomials against one setup in a single call (rows split over the GPUs): the EIP-4844 blob shape.
    /// Not in the reference, whose callers loop over `commit`.
    pub fn commit_batch(polynomials: &[Polynomial], setup_artifacts: &[SetupArtifact]) -> Result<Vec<G1Point>, anyhow::Error> {
        let n = polynomials.iter().map(|p| p.coefficients.len()).max().unwrap_or(0);
        let mut rows = vec![0u8; polynomials.len() * n * 32];         // shorter rows are zero padded: trimmed again inside
        for (b, p) in polynomials.iter().enumerate() {
            rows[b * n * 32..b * n * 32 + p.coefficients.len() * 32].copy_from_slice(&p.le_bytes());
        }
        let mut out = vec![0u8; polynomials.len() * 48];
        let mut eng = engine();
        let setup = eng.resident(setup_artifacts)?;
        eng.check(unsafe { kzg_commit_batch(eng.ctx(), setup, rows.as_ptr(), polynomials.len() as u64, n as u64, out.as_mut_ptr()) })?;
        Ok(out.chunks_exact(48).map(|c| G1Point::from_engine(c.try_into().unwrap())).collect())
    }
}

impl std::fmt::Display for Polynomial {
    /// highest degree first, zero coefficients skipped, "0" for the empty polynomial (`src/polynomial.rs:218-247`)
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        let Some((top, rest)) = self.coefficients.split_last() else {
            return write!(f, "0");
        };
        let term = |c: &Scalar, degree: usize| match degree {
            0 => format!("{c}"),
            1 => format!("{c}x"),
            d => format!("{c}x^{d}"),
        };
        let mut shown = term(top, rest.len());
        for (i, c) in rest.iter().enumerate().rev() {
            if !c.is_zero() {
                shown += &format!(" + {}", term(c, i));
            }
        }
        write!(f, "{shown}")
    }
}

#[derive(Debug, Clone, Deserialize, Serialize)]
pub struct Evaluation {
    pub point: Scalar,
    pub result: Scalar,
}

impl Evaluation {
    /// proof = commit((p - result) / (x - point)) (`src/polynomial.rs:260-269`).  A `result` that is not p(point)
    /// fails like the reference's `divide_by_root`: "constant terms do not add up", or "Unable to divide a constant
    /// polynomial" when p is a constant.
    pub fn generate_proof(&self, polynomial: &Polynomial, setup_artifacts: &[SetupArtifact]) -> Result<G1Point, anyhow::Error> {
        // kzg_open trims trailing zero coefficients itself, as the reference's `sub` does through try_from
        let coeffs = polynomial.le_bytes();
        let (mut y, mut proof) = ([0u8; 32], [0u8; 48]);
        let mut eng = engine();
        let setup = eng.resident(setup_artifacts)?;
        eng.check(unsafe {
            kzg_open(eng.ctx(), setup, coeffs.as_ptr(), polynomial.coefficients.len() as u64, self.point.to_le_bytes().as_ptr(),
                     self.result.to_le_bytes().as_ptr(), y.as_mut_ptr(), proof.as_mut_ptr())
        })?;
        Ok(G1Point::from_engine(proof))
    }

    /// e(proof, [s]G2 - [point]G2) == e(commitment - [result]G1, G2) (`src/polynomial.rs:276-294`); like the reference
    /// this indexes `setup_artifacts[1]` unguarded and panics on a shorter setup
    pub fn verify_proof(&self, proof: &G1Point, commitment: &G1Point, setup_artifacts: &[SetupArtifact]) -> Result<bool, anyhow::Error> {
        let tau: &G2Point = &setup_artifacts[1].g2;
        let mut ok = 0i32;
        let st = unsafe {
            kzg_verify_proof(proof.as_compressed().as_ptr(), commitment.as_compressed().as_ptr(), self.point.to_le_bytes().as_ptr(),
                             self.result.to_le_bytes().as_ptr(), tau.as_compressed().as_ptr(), &mut ok)
        };
        if st != KZG_OK {
            return Err(anyhow::anyhow!("pairing check failed with status {st}"));
        }
        Ok(ok == 1)
    }
}

#[cfg(test)]
mod tests {
    use super::*;

    #[test]
    fn trailing_zeros_are_trimmed_but_one_zero_stays() {
        assert_eq!(Polynomial::try_from(vec![0i128, 0, 0, 0, 0]).unwrap().degree(), 0);
        assert_eq!(Polynomial::try_from(vec![0i128, 0, 0]).unwrap().coefficients.len(), 1);
        assert_eq!(Polynomial::try_from(vec![1i128, 0, 1, 0, 0]).unwrap().degree(), 2);
        assert_eq!(Polynomial::try_from(vec![1i128, 0, 1, 0, 0, 5]).unwrap().degree(), 5);
        assert!(Polynomial::from(Scalar::from_i128(0)).coefficients.is_empty());
    }

    #[test]
    fn display() {
        assert_eq!(Polynomial::try_from(vec![2i128, 1, 0, 3]).unwrap().to_string(), "3x^3 + 1x + 2");
        assert_eq!(Polynomial::from(Scalar::from_i128(0)).to_string(), "0");
    }
}
