#!/usr/bin/env python3
"""Writes tests/golden/bench_blob_setup.json: golden values for the SHARDED configs bench.py runs at every GPU count.

  * BASELINE config C4 (4096 blobs x 4096 coefficients U(0xC4, .), points U(0xC4F, b), bench-secret setup): for each
    eighth of the batch (512 consecutive blobs — every 1/2/4/8-way split by blob is a union of eighths) the group sum of
    the commitments, the group sum of the proofs and the sum of the evaluations mod r, by the known-secret identities
    commit = [p(s)]G1, proof = [(p(s) - p(z)) / (s - z)]G1 with the bigint oracle.
  * BASELINE config C5 (2^22 setup points): [s^i]G1 at every eighth boundary (i = k 2^19 - 1, k 2^19, k 2^19 + 1) and at
    the ends, so that every rank of an index-range split can check items it generated itself.
Derived here with oracle/kzg_oracle.py (the reference pins no G1 bytes); takes about a minute."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)
import kzg_oracle as O
from kzg_poly_commit_exploration_b200 import synthetic

R = O.R
s = O.fr_from_be_bytes(O.BENCH_SECRET_BE)
N = BLOBS = 4096
EIGHTH = BLOBS // 8


def horner(co, x):
    acc = 0
    for c in reversed(co):
        acc = (acc * x + c) % R
    return acc


out = {"workload": "4096 blobs x 4096 coefficients U(0xC4, b*4096+i), z_b = U(0xC4F, b), secret 00..1f", "eighths": []}
zs_raw = synthetic.scalar_stream(0xC4F, BLOBS).tobytes()
for e in range(8):
    raw = synthetic.scalar_stream(0xC4, EIGHTH * N, start=e * EIGHTH * N).tobytes()
    ps = qs = ys = 0
    for b in range(EIGHTH):
        row = [int.from_bytes(raw[(b * N + i) * 32: (b * N + i) * 32 + 32], "little") for i in range(N)]
        z = int.from_bytes(zs_raw[(e * EIGHTH + b) * 32: (e * EIGHTH + b) * 32 + 32], "little")
        p_s, y = horner(row, s), horner(row, z)
        ps = (ps + p_s) % R
        ys = (ys + y) % R
        qs = (qs + (p_s - y) * pow((s - z) % R, -1, R)) % R
    out["eighths"].append({"first_blob": e * EIGHTH, "blobs": EIGHTH,
                           "sum_commit": O.g1_compress(O.g1_from_scalar(ps)).hex(),
                           "sum_proof": O.g1_compress(O.g1_from_scalar(qs)).hex(),
                           "sum_y_le": O.fr_to_le_bytes(ys).hex()})
    print("eighth", e, "done", flush=True)

n5 = 1 << 22
idx = sorted({i for k in range(9) for i in (k * (n5 // 8) - 1, k * (n5 // 8), k * (n5 // 8) + 1) if 0 <= i < n5} | {2, n5 - 2})
out["setup_2^22"] = {"secret_be": O.BENCH_SECRET_BE.hex(), "items": {str(i): O.g1_compress(O.g1_from_scalar(pow(s, i, R))).hex() for i in idx}}
with open(os.path.join(ROOT, "tests", "golden", "bench_blob_setup.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote tests/golden/bench_blob_setup.json")
