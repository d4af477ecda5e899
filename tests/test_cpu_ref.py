"""The C restatement of the reference loops (oracle/cpu_ref.c) against the Python oracle and the
golden vectors — it is bench.py's cpu_baseline, so it has to be right before it is timed."""
import json
import os
import random

import cpu_ref
import kzg_oracle as O
import util

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kat_bench_inputs.json")


def test_setup_and_golden_vectors():
    with open(GOLDEN) as f:
        kat = json.load(f)
    st = cpu_ref.RefSetup(O.BENCH_SECRET_BE, 1024)
    assert [st.export(i).hex() for i in range(8)] == kat["setup_g1"]
    mt = cpu_ref.RefSetup(O.BENCH_SECRET_BE, 64, threads=3)
    assert all(mt.export(i) == st.export(i) for i in range(64))
    for deg in ("1", "100", "1000", "1023"):
        v = kat["degrees"][deg]
        d = int(deg)
        co = util.from_ints(O.bench_polynomial(d))
        z = util.le(O.bench_input_point(d))
        rc, c = st.commit(co)
        y = cpu_ref.evaluate(co, z)
        rc2, pr = st.open(co, z, y)
        assert (rc, rc2) == (0, 0)
        assert (c.hex(), y.hex(), pr.hex()) == (v["commit"], v["y_le"], v["proof"])
        assert st.commit(co, threads=4)[1] == c


def test_random_inputs_and_error_codes():
    rnd = random.Random(3)
    secret = rnd.randbytes(32)      # >= r about half of the time: exercises the reduction
    st = cpu_ref.RefSetup(secret, 40)
    so = O.setup_generate(secret, 40)
    for n in (1, 2, 17, 40):
        ints = [rnd.randrange(1 << 256) for _ in range(n)]
        co = util.from_ints(ints)
        p = O.Polynomial.try_from_scalars(ints)
        z = rnd.randrange(1 << 256)
        rc, c = st.commit(co)
        assert rc == 0 and c == O.g1_compress(p.commit(so))
        y = cpu_ref.evaluate(co, util.le(z))
        ev = p.evaluate(z)
        assert y == O.fr_to_le_bytes(ev.result)
        rc, pr = st.open(co, util.le(z), y)
        if n == 1:
            assert rc == 0 and pr == O.g1_compress(None)
        else:
            assert rc == 0 and pr == O.g1_compress(ev.generate_proof(p, so))
    assert st.commit(util.from_ints([1] * 41))[0] == 1                       # degree too high
    assert st.open(util.from_ints([1, 2, 3]), util.le(5), util.le(1))[0] == 3  # constant terms do not add up
    assert st.open(util.from_ints([7]), util.le(5), util.le(1))[0] == 2        # constant polynomial
    assert st.commit(util.from_ints([0, 0, 0]))[1] == O.g1_compress(None)
