//! Drop-in for the hot path of the reference crate: same names, one FFI call per bulk operation.
//! `Scalar`, `G1Point`, `SetupArtifact` are the reference's own types (src/scalar.rs, src/curves.rs,
//! src/trusted_setup.rs); only the loops of `Polynomial::commit` (src/polynomial.rs:200-215),
//! `Polynomial::evaluate` (:112-123), `Evaluation::generate_proof` (:260-269) and
//! `SetupArtifactsGenerator` (src/trusted_setup.rs:40-78) move to the GPU.
use crate::ffi::*;
use std::ffi::CStr;

pub struct Engine { ctx: *mut kzg_ctx }
pub struct DeviceSetup { ctx: *mut kzg_ctx, h: *mut kzg_setup }

fn err(ctx: *mut kzg_ctx, st: i32) -> anyhow::Error {
    // messages equal the reference's anyhow strings for statuses 1..=3 (kzg_status_string)
    let msg = unsafe { CStr::from_ptr(kzg_last_error(ctx)) }.to_string_lossy().into_owned();
    anyhow::anyhow!("{msg} (kzg status {st})")
}

impl Engine {
    pub fn new(device: i32) -> Result<Self, anyhow::Error> {
        let mut ctx = std::ptr::null_mut();
        let st = unsafe { kzg_ctx_create(&mut ctx, device) };
        if st != KZG_OK { return Err(anyhow::anyhow!("no CUDA device: the engine has no CPU fallback")); }
        Ok(Self { ctx })
    }
    /// SetupArtifactsGenerator::new(secret).take(n) — G1 half, device resident
    pub fn setup_generate(&self, secret_be: [u8; 32], n: usize) -> Result<DeviceSetup, anyhow::Error> {
        let mut h = std::ptr::null_mut();
        let st = unsafe { kzg_setup_generate(self.ctx, secret_be.as_ptr(), 0, n as u64, &mut h) };
        if st != KZG_OK { return Err(err(self.ctx, st)); }
        Ok(DeviceSetup { ctx: self.ctx, h })
    }
    /// upload an existing `&[SetupArtifact]` (compressed g1 bytes, as Serialize emits them)
    pub fn setup_import(&self, g1_compressed: &[u8]) -> Result<DeviceSetup, anyhow::Error> {
        let mut h = std::ptr::null_mut();
        let st = unsafe { kzg_setup_import_g1(self.ctx, g1_compressed.as_ptr(), (g1_compressed.len() / 48) as u64, &mut h) };
        if st != KZG_OK { return Err(err(self.ctx, st)); }
        Ok(DeviceSetup { ctx: self.ctx, h })
    }
    /// Polynomial::evaluate: `coeffs` = concatenated Scalar::to_le_bytes()
    pub fn evaluate(&self, coeffs: &[u8], z_le: [u8; 32]) -> Result<[u8; 32], anyhow::Error> {
        let mut y = [0u8; 32];
        let st = unsafe { kzg_evaluate(self.ctx, coeffs.as_ptr(), (coeffs.len() / 32) as u64, z_le.as_ptr(), y.as_mut_ptr()) };
        if st != KZG_OK { return Err(err(self.ctx, st)); }
        Ok(y)
    }
}
impl Drop for Engine { fn drop(&mut self) { unsafe { kzg_ctx_destroy(self.ctx) } } }

impl DeviceSetup {
    pub fn len(&self) -> usize { unsafe { kzg_setup_len(self.h) as usize } }
    /// Polynomial::commit -> 48 compressed bytes (feed to G1Point's Deserialize)
    pub fn commit(&self, coeffs: &[u8]) -> Result<[u8; 48], anyhow::Error> {
        let mut out = [0u8; 48];
        let st = unsafe { kzg_commit(self.ctx, self.h, coeffs.as_ptr(), (coeffs.len() / 32) as u64, out.as_mut_ptr()) };
        if st != KZG_OK { return Err(err(self.ctx, st)); }
        Ok(out)
    }
    /// Evaluation::generate_proof: `result_le` is Evaluation.result; a wrong value reproduces divide_by_root's errors
    pub fn generate_proof(&self, coeffs: &[u8], point_le: [u8; 32], result_le: [u8; 32]) -> Result<[u8; 48], anyhow::Error> {
        let (mut y, mut out) = ([0u8; 32], [0u8; 48]);
        let st = unsafe { kzg_open(self.ctx, self.h, coeffs.as_ptr(), (coeffs.len() / 32) as u64, point_le.as_ptr(),
                                   result_le.as_ptr(), y.as_mut_ptr(), out.as_mut_ptr()) };
        if st != KZG_OK { return Err(err(self.ctx, st)); }
        Ok(out)
    }
}
impl Drop for DeviceSetup { fn drop(&mut self) { unsafe { kzg_setup_destroy(self.h) } } }
