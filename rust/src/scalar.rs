//! `Scalar` — an element of Fr kept as its canonical 32-byte little-endian integer (< r), the form the reference
//! serialises (`/root/reference/src/scalar.rs:226-233`).  Arithmetic goes to the engine's host helpers
//! (`kzg_fr_*`): same results as blst's `blst_fr_*`, since the encoding of a residue is unique.
use crate::ffi::*;
use serde::{de, Deserialize, Serialize};
use std::fmt::Display;

/// r = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001 (`src/scalar.rs:10`)
#[derive(Debug, Clone, PartialEq, Default)]
pub struct Scalar([u8; 32]);

impl From<i128> for Scalar {
    /// positive -> the value; zero or negative -> r - |a| with 0 -> 0 (`src/scalar.rs:27-48`); `i128::MIN` is fine
    fn from(a: i128) -> Self {
        let mut le = [0u8; 32];
        le[..16].copy_from_slice(&a.unsigned_abs().to_le_bytes());
        let magnitude = Scalar(le);
        if a > 0 { magnitude } else { magnitude.neg() }
    }
}

impl Scalar {
    /// 32 little-endian bytes; values >= r are reduced mod r (`src/scalar.rs:54-61`)
    pub fn from_le_bytes(b: [u8; 32]) -> Self {
        let mut out = [0u8; 32];
        unsafe { kzg_fr_reduce(b.as_ptr(), out.as_mut_ptr()) };
        Scalar(out)
    }

    /// 32 big-endian bytes (`src/scalar.rs:66-73`)
    pub fn from_be_bytes(mut b: [u8; 32]) -> Self {
        b.reverse();
        Self::from_le_bytes(b)
    }

    pub fn from_i128(a: i128) -> Self {
        Self::from(a)
    }

    pub fn to_le_bytes(&self) -> [u8; 32] {
        self.0
    }

    pub fn to_be_bytes(&self) -> [u8; 32] {
        let mut b = self.0;
        b.reverse();
        b
    }

    pub fn mul(&self, other: &Self) -> Self {
        let mut out = [0u8; 32];
        unsafe { kzg_fr_mul(self.0.as_ptr(), other.0.as_ptr(), out.as_mut_ptr()) };
        Scalar(out)
    }

    /// x^n, `pow(0) = 1` (`src/scalar.rs:122-187`)
    pub fn pow(&self, n: usize) -> Self {
        let mut out = [0u8; 32];
        unsafe { kzg_fr_pow(self.0.as_ptr(), n as u64, out.as_mut_ptr()) };
        Scalar(out)
    }

    pub fn add(&self, other: &Self) -> Self {
        let mut out = [0u8; 32];
        unsafe { kzg_fr_add(self.0.as_ptr(), other.0.as_ptr(), out.as_mut_ptr()) };
        Scalar(out)
    }

    pub fn sub(&self, other: &Self) -> Self {
        let mut out = [0u8; 32];
        unsafe { kzg_fr_sub(self.0.as_ptr(), other.0.as_ptr(), out.as_mut_ptr()) };
        Scalar(out)
    }

    pub fn neg(&self) -> Self {
        let mut out = [0u8; 32];
        unsafe { kzg_fr_neg(self.0.as_ptr(), out.as_mut_ptr()) };
        Scalar(out)
    }

    pub fn is_zero(&self) -> bool {
        self.0.iter().all(|b| *b == 0)
    }
}

impl Serialize for Scalar {
    /// the 32 LE bytes as a byte string: serde_json writes `[b0,b1,...,b31]` (`src/scalar.rs:226-233`)
    fn serialize<S: serde::Serializer>(&self, serializer: S) -> Result<S::Ok, S::Error> {
        serializer.serialize_bytes(&self.0)
    }
}

/// Collects a serde sequence of u8 — the shape all three canonical encodings arrive in from JSON.
pub(crate) struct ByteSeq;

impl<'de> de::Visitor<'de> for ByteSeq {
    type Value = Vec<u8>;

    fn expecting(&self, f: &mut std::fmt::Formatter) -> std::fmt::Result {
        f.write_str("Sequence of u8")
    }

    fn visit_seq<A: de::SeqAccess<'de>>(self, mut seq: A) -> Result<Vec<u8>, A::Error> {
        let mut bytes = Vec::with_capacity(seq.size_hint().unwrap_or(96));
        while let Some(b) = seq.next_element::<u8>()? {
            bytes.push(b);
        }
        Ok(bytes)
    }

    fn visit_bytes<E: de::Error>(self, v: &[u8]) -> Result<Vec<u8>, E> {
        Ok(v.to_vec())
    }
}

impl<'de> Deserialize<'de> for Scalar {
    fn deserialize<D: serde::Deserializer<'de>>(deserializer: D) -> Result<Self, D::Error> {
        let bytes = deserializer.deserialize_seq(ByteSeq)?;
        let le: [u8; 32] = bytes.as_slice().try_into().map_err(|_| {
            de::Error::custom(format!("Invalid byte array, expected length 32, got {}", bytes.len()))
        })?;
        Ok(Scalar::from_le_bytes(le))
    }
}

impl Display for Scalar {
    /// base-10 digits of the canonical value, "" for zero like the reference's digit loop (`src/scalar.rs:277-345`)
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        let mut q = self.0;
        let mut digits: Vec<u8> = vec![];
        while q.iter().any(|b| *b != 0) {
            let mut rem: u16 = 0;
            for byte in q.iter_mut().rev() {
                let cur = (rem << 8) | *byte as u16;
                *byte = (cur / 10) as u8;
                rem = cur % 10;
            }
            digits.push(b'0' + rem as u8);
        }
        digits.reverse();
        f.write_str(std::str::from_utf8(&digits).unwrap_or(""))
    }
}
