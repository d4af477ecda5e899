#!/usr/bin/env python3
"""Writes tests/golden/kat_bench_inputs.json: known-answer vectors DERIVED (not shipped by the
reference — it pins no G1 bytes) from the reference's deterministic bench inputs
(benches/polynomial_commitment.rs:10-23, benches/evaluation_proof.rs:10-27) with the bigint oracle.
The same values were derived independently by the survey's scratch script."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kzg_oracle as O

s = O.fr_from_be_bytes(O.BENCH_SECRET_BE)
out = {"secret_be": O.BENCH_SECRET_BE.hex(), "setup_g1": [], "degrees": {}}
st = O.setup_generate(O.BENCH_SECRET_BE, 8)
out["setup_g1"] = [O.g1_compress(a.g1).hex() for a in st]
out["setup1_g2"] = O.g2_compress(st[1].g2).hex()
for d in (1, 100, 500, 1000, 1023, 2500):
    co, z = O.bench_polynomial(d), O.bench_input_point(d)
    out["degrees"][str(d)] = {
        "commit": O.g1_compress(O.commit_known_secret(co, s)).hex(),
        "y_le": O.fr_to_le_bytes(O.horner(co, z)).hex(),
        "proof": O.g1_compress(O.proof_known_secret(co, s, z)).hex(),
    }
with open(os.path.join(ROOT, "tests", "golden", "kat_bench_inputs.json"), "w") as f:
    json.dump(out, f, indent=1)
# the commitment of bench.py's workload (BASELINE config C3): n = 2^20 scalars U(0xC3, i) over the bench-secret setup,
# by the known-secret identity [p(s)]G1; bench.py asserts every timed result against it at every GPU count
sys.path.insert(0, ROOT)
from kzg_poly_commit_exploration_b200 import synthetic
raw = synthetic.scalar_stream(0xC3, 1 << 20).tobytes()
ints = [int.from_bytes(raw[i: i + 32], "little") for i in range(0, len(raw), 32)]
with open(os.path.join(ROOT, "tests", "golden", "bench_commitment.json"), "w") as f:
    json.dump({"workload": "n=2^20, scalars U(0xC3,i), secret 00..1f", "commit": O.g1_compress(O.commit_known_secret(ints, s)).hex()}, f, indent=1)
print("wrote golden vectors")
