"""Signed-digit recoding of folded scalars (fr_fold_half + ceil(255/c) windows), host build of the engine's own
headers against Python bigints: the claim that W = ceil(255/c) windows always absorb the top carry (SURVEY A.2 with
the scalar folded to min(k, r - k)) is what lets c = 15 / 17 drop one window."""
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle")]
import kzg_oracle as O

SRC = os.path.join(ROOT, "tests", "cpp", "test_recode.cpp")
BIN = os.path.join(ROOT, "tests", "cpp", "test_recode.bin")
CSRC = os.path.join(ROOT, "kzg_poly_commit_exploration_b200", "csrc")


@pytest.fixture(scope="module")
def recode_bin():
    deps = [SRC, os.path.join(CSRC, "field.cuh"), os.path.join(CSRC, "recode.cuh")]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        subprocess.run(["g++", "-std=c++17", "-O1", "-I", CSRC, "-x", "c++", SRC, "-o", BIN], check=True)
    return BIN


def test_folded_signed_digits_reconstruct_the_scalar(recode_bin):
    R = O.R
    rnd = random.Random(255)
    scalars = [0, 1, 2, (R - 1) // 2, (R + 1) // 2, R - 1, R - 2, R, R + 5, 2**256 - 1, 2**255, 2**254, 2**254 - 1,
               2**253, (1 << 238) - 1, 0x8000 << 240]
    scalars += [rnd.getrandbits(256) for _ in range(200)]
    scalars += [(rnd.getrandbits(256) % R) | ((1 << 16) - 1) for _ in range(20)]          # long carry runs
    scalars += [((1 << c) - 1) << s for c in (13, 15, 16, 17) for s in (0, 100, 230)]
    windows = [str(c) for c in range(2, 25)]
    inp = "\n".join(f"{k:064x}" for k in scalars) + "\n"
    out = subprocess.run([recode_bin, *windows], input=inp, capture_output=True, text=True, check=True).stdout.split("\n")
    lines = [l for l in out if l]
    assert len(lines) == len(scalars) * len(windows)
    it = iter(lines)
    for k in scalars:
        kc = k % R if k < 3 * R else None            # inputs are 256-bit: two conditional subtractions reduce them
        assert kc is not None
        for c in range(2, 25):
            f = next(it).split()
            cc, flip, W = int(f[0]), int(f[1]), int(f[2])
            digs = [int(x) for x in f[3:3 + W]]
            carry = int(f[3 + W])
            assert cc == c and W == -(-255 // c)
            assert flip == (1 if kc > R - kc else 0)
            assert carry == 0, (hex(k), c)
            assert all(abs(d) <= 1 << (c - 1) for d in digs)
            assert sum(d << (c * j) for j, d in enumerate(digs)) % R == kc, (hex(k), c)
