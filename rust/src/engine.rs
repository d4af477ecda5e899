//! The process-wide engine behind the reference-shaped API.
//!
//! The reference passes `&[SetupArtifact]` to every `commit` / `generate_proof` call
//! (`src/polynomial.rs:200`, `:260`); the engine keeps the matching DEVICE-resident setup (affine bases + the
//! precomputed window tables) in a small cache keyed by a fingerprint of the slice, so the first call with a
//! given setup uploads it (one `kzg_setup_import_g1`) and every later call is a single `kzg_commit` / `kzg_open`.
//!
//! One context drives every GPU of the node from this process (`kzg_ctx_create_multi`): a setup above 2^16 points is
//! sharded by point range and each commitment is summed on all devices at once; smaller setups are replicated and
//! batches split by row.  `KZG_B200_DEVICES=0,2,3` narrows the device list.  There is no CPU fallback: without a CUDA
//! device the first bulk call panics with the message below (the reference has no failure mode to map it to).
use crate::ffi::*;
use crate::trusted_setup::SetupArtifact;
use std::ffi::CStr;
use std::sync::{Mutex, MutexGuard, OnceLock};

/// A device-resident trusted setup (bases + tables).  Dropped before the context by construction: setups only live
/// inside [`Engine`], which owns both.
struct ResidentSetup {
    fingerprint: (usize, u64),
    handle: *mut kzg_setup,
}

pub struct Engine {
    ctx: *mut kzg_ctx,
    setups: Vec<ResidentSetup>,
}

// the C library wants one caller at a time per context: the Mutex in `engine()` provides exactly that
unsafe impl Send for Engine {}

const MAX_RESIDENT_SETUPS: usize = 4;

static ENGINE: OnceLock<Mutex<Engine>> = OnceLock::new();

/// Lock the process-wide engine (created on first use).
pub fn engine() -> MutexGuard<'static, Engine> {
    ENGINE
        .get_or_init(|| Mutex::new(Engine::new().expect("no usable CUDA device: the B200 engine has no CPU fallback")))
        .lock()
        .unwrap_or_else(|poisoned| poisoned.into_inner())
}

fn devices_from_env() -> Vec<i32> {
    if let Ok(list) = std::env::var("KZG_B200_DEVICES") {
        let ids: Vec<i32> = list.split(',').filter_map(|t| t.trim().parse().ok()).collect();
        if !ids.is_empty() {
            return ids;
        }
    }
    // every device that accepts a context (the C ABI carries no CUDA types, so counting is probing)
    let mut ids = vec![];
    for d in 0..32 {
        let mut probe = std::ptr::null_mut();
        if unsafe { kzg_ctx_create(&mut probe, d) } != KZG_OK {
            break;
        }
        unsafe { kzg_ctx_destroy(probe) };
        ids.push(d);
    }
    ids
}

/// FNV-1a over the first, middle and last G1 encodings and the length: cheap, and a different setup of the same length
/// at the same address (a reused Vec) still gets its own entry.
fn fingerprint(setup: &[SetupArtifact]) -> (usize, u64) {
    let mut h: u64 = 0xcbf29ce484222325;
    let mut eat = |bytes: &[u8]| {
        for b in bytes {
            h = (h ^ *b as u64).wrapping_mul(0x100000001b3);
        }
    };
    if !setup.is_empty() {
        for i in [0, setup.len() / 2, setup.len() - 1] {
            eat(setup[i].g1.as_compressed());
        }
    }
    (setup.len(), h)
}

impl Engine {
    fn new() -> Result<Self, anyhow::Error> {
        let ids = devices_from_env();
        if ids.is_empty() {
            return Err(anyhow::anyhow!("no CUDA device"));
        }
        let mut ctx = std::ptr::null_mut();
        let st = unsafe { kzg_ctx_create_multi(&mut ctx, ids.as_ptr(), ids.len() as i32) };
        if st != KZG_OK {
            return Err(anyhow::anyhow!("kzg_ctx_create_multi failed with status {st}"));
        }
        Ok(Self { ctx, setups: vec![] })
    }

    pub fn ctx(&self) -> *mut kzg_ctx {
        self.ctx
    }

    /// The reference's `anyhow` message for a failed call: statuses 1..=3 carry its exact strings
    /// (`kzg_status_string`), so callers and tests that match on the text keep working.
    pub fn error(&self, st: i32) -> anyhow::Error {
        let msg = unsafe { CStr::from_ptr(kzg_last_error(self.ctx)) }.to_string_lossy().into_owned();
        if msg.is_empty() {
            let generic = unsafe { CStr::from_ptr(kzg_status_string(st)) }.to_string_lossy().into_owned();
            return anyhow::anyhow!("{generic}");
        }
        anyhow::anyhow!("{msg}")
    }

    pub fn check(&self, st: i32) -> Result<(), anyhow::Error> {
        if st == KZG_OK { Ok(()) } else { Err(self.error(st)) }
    }

    /// Device-resident twin of `setup` (uploaded on first sight, least recently used entry evicted).
    pub fn resident(&mut self, setup: &[SetupArtifact]) -> Result<*mut kzg_setup, anyhow::Error> {
        let key = fingerprint(setup);
        if let Some(pos) = self.setups.iter().position(|s| s.fingerprint == key) {
            let hit = self.setups.remove(pos);
            let handle = hit.handle;
            self.setups.push(hit);
            return Ok(handle);
        }
        let mut g1 = Vec::with_capacity(setup.len() * 48);
        for item in setup {
            g1.extend_from_slice(item.g1.as_compressed());
        }
        let mut handle = std::ptr::null_mut();
        self.check(unsafe { kzg_setup_import_g1(self.ctx, g1.as_ptr(), setup.len() as u64, &mut handle) })?;
        if self.setups.len() == MAX_RESIDENT_SETUPS {
            let old = self.setups.remove(0);
            unsafe { kzg_setup_destroy(old.handle) };
        }
        self.setups.push(ResidentSetup { fingerprint: key, handle });
        Ok(handle)
    }

    /// Adopt a setup that already lives on the device (the generator's fast path) under the fingerprint of its host twin.
    pub fn adopt(&mut self, setup: &[SetupArtifact], handle: *mut kzg_setup) {
        let key = fingerprint(setup);
        if self.setups.iter().any(|s| s.fingerprint == key) {
            unsafe { kzg_setup_destroy(handle) };
            return;
        }
        if self.setups.len() == MAX_RESIDENT_SETUPS {
            let old = self.setups.remove(0);
            unsafe { kzg_setup_destroy(old.handle) };
        }
        self.setups.push(ResidentSetup { fingerprint: key, handle });
    }
}

impl Drop for Engine {
    fn drop(&mut self) {
        for s in self.setups.drain(..) {
            unsafe { kzg_setup_destroy(s.handle) };      // setups go before their context (kzg_b200.h)
        }
        unsafe { kzg_ctx_destroy(self.ctx) };
    }
}
