"""The Rust drop-in crate (rust/) cannot be compiled in this image (no cargo / rustc), so it is kept mechanically
checkable: the raw bindings are generated from include/kzg_b200.h and compared with it symbol by symbol, and the public
surface of the reference crate (names and parameter lists of src/scalar.rs, src/curves.rs,
src/trusted_setup.rs, src/polynomial.rs) must be present in the module of the same name."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUST = os.path.join(ROOT, "rust")


def test_ffi_rs_matches_the_c_header():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_rust_ffi.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "in step" in r.stdout


def test_ffi_rs_covers_every_exported_symbol(native):
    """every symbol the Python binding resolves in the .so (all of include/kzg_b200.h) is declared in ffi.rs"""
    from kzg_poly_commit_exploration_b200 import _native
    src = open(os.path.join(RUST, "src", "ffi.rs")).read()
    declared = set(re.findall(r"pub fn (kzg_\w+)\(", src))
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)


def _norm(s):
    return re.sub(r"\s+", " ", s).strip()


# module -> signatures (normalised text) the reference exports; file:line into /root/reference/src
SURFACE = {
    "scalar.rs": [
        "pub struct Scalar(",                                             # scalar.rs:7-8
        "impl From<i128> for Scalar",                                     # :27
        "pub fn from_le_bytes(b: [u8; 32]) -> Self",                      # :54
        "pub fn from_be_bytes(",                                          # :66
        "pub fn from_i128(a: i128) -> Self",                              # :78
        "pub fn to_le_bytes(&self) -> [u8; 32]",                          # :83
        "pub fn to_be_bytes(&self) -> [u8; 32]",                          # :96
        "pub fn mul(&self, other: &Self) -> Self",                        # :111
        "pub fn pow(&self, n: usize) -> Self",                            # :122
        "pub fn add(&self, other: &Self) -> Self",                        # :192
        "pub fn sub(&self, other: &Self) -> Self",                        # :203
        "pub fn neg(&self) -> Self",                                      # :212
        "pub fn is_zero(&self) -> bool",                                  # :221
        "impl Serialize for Scalar", "impl<'de> Deserialize<'de> for Scalar", "impl Display for Scalar",
        "#[derive(Debug, Clone, PartialEq, Default)]",
        "Invalid byte array, expected length 32, got {}",                 # :262-265
    ],
    "curves.rs": [
        "pub struct G1Point(", "pub struct G2Point(",                     # curves.rs:11, 186
        "pub fn from_i128(a: i128) -> Self",                              # :28, :203
        "pub fn from_scalar(a: Scalar) -> Self",                          # :50, :225
        "pub fn sub(&self, b: &Self) -> Self",                            # :66, :241
        "pub fn add(&self, b: &Self) -> Self",                            # :79
        "pub fn mult(&self, a: &Scalar) -> Self",                         # :90
        "impl Serialize for G1Point", "impl<'de> Deserialize<'de> for G1Point",
        "impl Serialize for G2Point", "impl<'de> Deserialize<'de> for G2Point",
        "pub fn bilinear_map(p1: &G1Point, p2: &G2Point) ->",             # :355
        "Invalid length, expected 48, got {}", "Invalid length, expected 96, got {}", "Got error while uncompressing",
    ],
    "trusted_setup.rs": [
        "pub struct SetupArtifactsGenerator",                             # trusted_setup.rs:9-14
        "pub fn new(secret: [u8; 32]) -> Self",                           # :20
        "pub struct SetupArtifact", "pub g1: G1Point", "pub g2: G2Point",  # :31-35
        "impl Iterator for SetupArtifactsGenerator", "type Item = SetupArtifact",   # :37-38
        "#[derive(Debug, Serialize, Deserialize)]",
    ],
    "polynomial.rs": [
        "pub struct Polynomial", "coefficients: Vec<Scalar>",             # polynomial.rs:9-12
        "impl TryFrom<Vec<i128>> for Polynomial", "impl TryFrom<Vec<Scalar>> for Polynomial", "impl From<Scalar> for Polynomial",
        "type Error = anyhow::Error",
        "pub fn degree(&self) -> u32",                                    # :95
        "pub fn evaluate(&self, x: Scalar) -> Result<Evaluation, anyhow::Error>",       # :112
        "pub fn sub(&self, other: &Self) -> Result<Self, anyhow::Error>",                # :128
        "pub fn divide_by_root(&self, root: &Scalar) -> Result<Self, anyhow::Error>",    # :150
        "pub fn commit(&self, setup_artifacts: &[SetupArtifact]) -> Result<G1Point, anyhow::Error>",   # :200
        "impl std::fmt::Display for Polynomial",
        "pub struct Evaluation", "pub point: Scalar", "pub result: Scalar",              # :249-253
        "pub fn generate_proof(&self, polynomial: &Polynomial, setup_artifacts: &[SetupArtifact]) -> Result<G1Point, anyhow::Error>",
        "pub fn verify_proof(&self, proof: &G1Point, commitment: &G1Point, setup_artifacts: &[SetupArtifact]) -> Result<bool, anyhow::Error>",
        "#[derive(Debug, Deserialize, Serialize)]", "#[derive(Debug, Clone, Deserialize, Serialize)]",
        "Setup does not allow for commitment generation of the polynomial. The polynomial degree is too high.",
        "Too many coefficients for polynomial, only 2**32 - 1 coefficients is supported. Got {}",
    ],
}


def test_rust_crate_keeps_the_reference_surface():
    for fname, needles in SURFACE.items():
        src = _norm(open(os.path.join(RUST, "src", fname)).read())
        for needle in needles:
            assert _norm(needle) in src, f"{fname}: missing `{needle}`"
    lib = open(os.path.join(RUST, "src", "lib.rs")).read()
    for mod in ("curves", "polynomial", "scalar", "trusted_setup"):          # src/lib.rs:1-4
        assert f"pub mod {mod};" in lib
    manifest = open(os.path.join(RUST, "Cargo.toml")).read()
    assert 'name = "kzg_poly_commit_exploration"' in manifest and "blst" not in manifest.split("[dependencies]")[1]


def test_rust_error_strings_are_the_library_strings(native):
    """statuses 1..3 carry the reference's anyhow messages, and the Rust shim forwards them verbatim (no suffix)"""
    L = native.lib()
    assert L.kzg_status_string(1).decode() in open(os.path.join(RUST, "src", "polynomial.rs")).read()
    eng = open(os.path.join(RUST, "src", "engine.rs")).read()
    assert 'anyhow::anyhow!("{msg}")' in eng and "kzg status" not in eng


def test_build_rs_lists_every_cuda_source():
    from kzg_poly_commit_exploration_b200 import build
    rs = open(os.path.join(RUST, "build.rs")).read()
    for src in build.SOURCES:
        assert f'"{src}"' in rs, src
