"""The straight-line PTX of the field kernels is verified on the CPU: tools/gen_field_ptx.py runs the
exact instruction lists it emits through a PTX-subset interpreter against Python bigints, and the
generated header in the tree must be what the generator produces."""
import os

import gen_consts
import gen_field_ptx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "kzg_poly_commit_exploration_b200", "csrc")


def test_instruction_lists_against_bigints():
    assert gen_field_ptx.check(iters=120, seed=99)


def test_generated_header_is_current():
    with open(os.path.join(CSRC, "field_ptx_gen.cuh")) as f:
        assert f.read() == gen_field_ptx.emit_cpp()


def test_multiply_instruction_budget():
    # 2N^2 + N multiply instructions per Montgomery product (300 wide-MAC for Fq, 136 for Fr)
    for name, (n, p) in gen_field_ptx.FIELDS.items():
        prog = gen_field_ptx.gen_mont_mul(n, p)
        muls = sum(1 for op, _, _ in prog.ins if op.startswith(("mul", "mad")))
        # every lo/hi PAIR fuses into one IMAD.WIDE: (muls - n) / 2 wide ops + n single multiplies for m
        assert muls == 4 * n * n + n, (name, muls)
        assert (muls - n) // 2 + n == {12: 300, 8: 136}[n]
