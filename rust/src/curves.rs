//! `G1Point`, `G2Point`, `bilinear_map` — group elements kept as their canonical compressed encodings (ZCash
//! format: 48 / 96 bytes), which is what the reference serialises (`/root/reference/src/curves.rs:99-110, 259-271`)
//! and all that ever leaves its opaque blst structs.  Per-element group operations go to the engine's host helpers.
//!
//! Not carried over: the three places where the reference leaks blst types — `From<blst_p1> for G1Point`
//! (`src/curves.rs:13-17`), `Deref<Target = blst_p2> for G2Point` (`:252-257`) and `bilinear_map`'s `blst_fp12`
//! return value (`:355`).  `bilinear_map` returns a [`PairingValue`] instead, and the only thing the reference does
//! with the result — comparing two of them (`src/polynomial.rs:293`) — works unchanged.
use crate::ffi::*;
use crate::scalar::{ByteSeq, Scalar};
use serde::{de, Deserialize, Serialize};

#[derive(Debug, Clone, PartialEq)]
pub struct G1Point([u8; 48]);

#[derive(Debug, Clone, PartialEq)]
pub struct G2Point([u8; 96]);

fn abs_scalar(a: i128) -> Scalar {
    let mut le = [0u8; 32];
    le[..16].copy_from_slice(&a.unsigned_abs().to_le_bytes());
    Scalar::from_le_bytes(le)
}

impl G1Point {
    /// [|a|]G1, negated for a < 0 (`src/curves.rs:28-45`)
    pub fn from_i128(a: i128) -> Self {
        let p = Self::from_scalar(abs_scalar(a));
        if a < 0 { Self::infinity().sub(&p) } else { p }
    }

    /// [a]G1 (`src/curves.rs:50-61`)
    pub fn from_scalar(a: Scalar) -> Self {
        let mut out = [0u8; 48];
        unsafe { kzg_g1_from_scalar(a.to_le_bytes().as_ptr(), out.as_mut_ptr()) };
        G1Point(out)
    }

    pub fn sub(&self, b: &Self) -> Self {
        let mut out = [0u8; 48];
        unsafe { kzg_g1_sub(self.0.as_ptr(), b.0.as_ptr(), out.as_mut_ptr()) };
        G1Point(out)
    }

    pub fn add(&self, b: &Self) -> Self {
        let mut out = [0u8; 48];
        unsafe { kzg_g1_add(self.0.as_ptr(), b.0.as_ptr(), out.as_mut_ptr()) };
        G1Point(out)
    }

    pub fn mult(&self, a: &Scalar) -> Self {
        let mut out = [0u8; 48];
        unsafe { kzg_g1_mul(self.0.as_ptr(), a.to_le_bytes().as_ptr(), out.as_mut_ptr()) };
        G1Point(out)
    }

    fn infinity() -> Self {
        let mut c = [0u8; 48];
        c[0] = 0xc0;
        G1Point(c)
    }

    /// the 48 compressed bytes (what `Serialize` writes)
    pub fn as_compressed(&self) -> &[u8; 48] {
        &self.0
    }

    /// `Deserialize`'s checks (`src/curves.rs:119-144`): length, then blst_p1_uncompress (flags, x < q, on the curve;
    /// no subgroup check)
    pub fn from_compressed(v: &[u8]) -> Result<Self, anyhow::Error> {
        let c: [u8; 48] = v.try_into().map_err(|_| anyhow::anyhow!("Invalid length, expected 48, got {}", v.len()))?;
        if unsafe { kzg_g1_check(c.as_ptr()) } != KZG_OK {
            return Err(anyhow::anyhow!("Got error while uncompressing: BLST_BAD_ENCODING"));
        }
        Ok(G1Point(c))
    }

    /// bytes the engine produced itself (already canonical)
    pub(crate) fn from_engine(c: [u8; 48]) -> Self {
        G1Point(c)
    }
}

impl Serialize for G1Point {
    fn serialize<S: serde::Serializer>(&self, serializer: S) -> Result<S::Ok, S::Error> {
        serializer.serialize_bytes(&self.0)
    }
}

impl<'de> Deserialize<'de> for G1Point {
    fn deserialize<D: serde::Deserializer<'de>>(deserializer: D) -> Result<Self, D::Error> {
        let bytes = deserializer.deserialize_seq(ByteSeq)?;
        G1Point::from_compressed(&bytes).map_err(de::Error::custom)
    }
}

impl G2Point {
    /// [|a|]G2, negated for a < 0 (`src/curves.rs:203-220`)
    pub fn from_i128(a: i128) -> Self {
        let p = Self::from_scalar(abs_scalar(a));
        if a >= 0 {
            return p;
        }
        let mut out = [0u8; 96];
        unsafe { kzg_g2_neg(p.0.as_ptr(), out.as_mut_ptr()) };
        G2Point(out)
    }

    /// [a]G2 (`src/curves.rs:225-236`)
    pub fn from_scalar(a: Scalar) -> Self {
        let mut out = [0u8; 96];
        unsafe { kzg_g2_from_scalar(a.to_le_bytes().as_ptr(), out.as_mut_ptr()) };
        G2Point(out)
    }

    pub fn sub(&self, b: &Self) -> Self {
        let mut out = [0u8; 96];
        unsafe { kzg_g2_sub(self.0.as_ptr(), b.0.as_ptr(), out.as_mut_ptr()) };
        G2Point(out)
    }

    pub fn as_compressed(&self) -> &[u8; 96] {
        &self.0
    }

    /// `Deserialize`'s checks (`src/curves.rs:279-297`)
    pub fn from_compressed(v: &[u8]) -> Result<Self, anyhow::Error> {
        let c: [u8; 96] = v.try_into().map_err(|_| anyhow::anyhow!("Invalid length, expected 96, got {}", v.len()))?;
        if unsafe { kzg_g2_check(c.as_ptr()) } != KZG_OK {
            return Err(anyhow::anyhow!("Got error while uncompressing: BLST_BAD_ENCODING"));
        }
        Ok(G2Point(c))
    }

    pub(crate) fn from_engine(c: [u8; 96]) -> Self {
        G2Point(c)
    }
}

impl Serialize for G2Point {
    fn serialize<S: serde::Serializer>(&self, serializer: S) -> Result<S::Ok, S::Error> {
        serializer.serialize_bytes(&self.0)
    }
}

impl<'de> Deserialize<'de> for G2Point {
    fn deserialize<D: serde::Deserializer<'de>>(deserializer: D) -> Result<Self, D::Error> {
        let bytes = deserializer.deserialize_seq(ByteSeq)?;
        G2Point::from_compressed(&bytes).map_err(de::Error::custom)
    }
}

/// e(p1, p2), kept as its arguments: the reference only ever compares two pairing values
/// (`src/polynomial.rs:293`), and equality of two pairings is one product of Miller loops and one final
/// exponentiation in the engine (`kzg_pairing_equal`) instead of two full pairings.
#[derive(Debug, Clone)]
pub struct PairingValue {
    g1: G1Point,
    g2: G2Point,
}

impl PartialEq for PairingValue {
    fn eq(&self, other: &Self) -> bool {
        let mut equal = 0i32;
        let st = unsafe {
            kzg_pairing_equal(self.g1.0.as_ptr(), self.g2.0.as_ptr(), other.g1.0.as_ptr(), other.g2.0.as_ptr(), &mut equal)
        };
        st == KZG_OK && equal == 1
    }
}

/// `bilinear_map(p1, p2)` (`src/curves.rs:355-371`)
pub fn bilinear_map(p1: &G1Point, p2: &G2Point) -> PairingValue {
    PairingValue { g1: p1.clone(), g2: p2.clone() }
}
