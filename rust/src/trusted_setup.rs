//! `SetupArtifactsGenerator`, `SetupArtifact` (`/root/reference/src/trusted_setup.rs`).
//!
//! The reference's iterator multiplies the generators by s^i one item at a time (a `blst_p1_mult` and a
//! `blst_p2_mult` per `next()`).  Here `next()` hands out items from a buffer that the engine refills in bulk: one
//! `kzg_setup_generate` + `kzg_setup_generate_g2` per refill (fixed-base kernels over all GPUs, index-range sharded),
//! so `SetupArtifactsGenerator::new(secret).take(n).collect()` costs O(log n) FFI calls instead of 2n group
//! multiplications.  Items are identical: item 0 is the generator pair whatever the secret, item i = ([s^i]G1, [s^i]G2).
use crate::curves::{G1Point, G2Point};
use crate::engine::engine;
use crate::ffi::*;
use serde::{Deserialize, Serialize};
use std::collections::VecDeque;

#[derive(Debug)]
pub struct SetupArtifactsGenerator {
    secret_be: [u8; 32],
    next_index: u64,
    ready: VecDeque<SetupArtifact>,
    refill: u64,
}

#[derive(Debug, Serialize, Deserialize)]
pub struct SetupArtifact {
    pub g1: G1Point,
    pub g2: G2Point,
}

const FIRST_REFILL: u64 = 16;
const LARGEST_REFILL: u64 = 1 << 20;

impl SetupArtifactsGenerator {
    /// * `secret` - Secret used to generate artifacts, in big endian bytes (reduced mod r, `src/trusted_setup.rs:20-28`)
    pub fn new(secret: [u8; 32]) -> Self {
        Self { secret_be: secret, next_index: 0, ready: VecDeque::new(), refill: FIRST_REFILL }
    }

    /// items `first .. first + count` straight from the engine
    fn generate(&self, first: u64, count: u64) -> Result<Vec<SetupArtifact>, anyhow::Error> {
        let eng = engine();
        let mut handle = std::ptr::null_mut();
        eng.check(unsafe { kzg_setup_generate(eng.ctx(), self.secret_be.as_ptr(), first, count, &mut handle) })?;
        let mut g1 = vec![0u8; count as usize * 48];
        let mut g2 = vec![0u8; count as usize * 96];
        let st = unsafe {
            let mut st = kzg_setup_generate_g2(eng.ctx(), handle, self.secret_be.as_ptr(), count);
            if st == KZG_OK { st = kzg_setup_export_g1(eng.ctx(), handle, 0, count, g1.as_mut_ptr()); }
            if st == KZG_OK { st = kzg_setup_export_g2(eng.ctx(), handle, 0, count, g2.as_mut_ptr()); }
            kzg_setup_destroy(handle);
            st
        };
        eng.check(st)?;
        Ok((0..count as usize)
            .map(|i| SetupArtifact {
                g1: G1Point::from_engine(g1[48 * i..48 * i + 48].try_into().unwrap()),
                g2: G2Point::from_engine(g2[96 * i..96 * i + 96].try_into().unwrap()),
            })
            .collect())
    }
}

impl Iterator for SetupArtifactsGenerator {
    type Item = SetupArtifact;

    fn next(&mut self) -> Option<Self::Item> {
        if self.ready.is_empty() {
            // refills double (16, 32, ... 2^20 items): short setups stay cheap, long ones amortise the launches
            let items = self.generate(self.next_index, self.refill).expect("trusted-setup generation failed on the device");
            self.next_index += self.refill;
            self.refill = (self.refill * 2).min(LARGEST_REFILL);
            self.ready.extend(items);
        }
        self.ready.pop_front()
    }
}
