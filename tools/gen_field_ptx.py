#!/usr/bin/env python3
"""Generator + CPU verifier for the sm_100a Montgomery field kernels (Fq 12x32, Fr 8x32).

The hot arithmetic of the engine is written as straight-line PTX: carry chains of
`mad.lo.cc / madc.hi.cc` pairs, which ptxas fuses into IMAD.WIDE.U32(.X) with the
carry in a predicate.  Each field routine is ONE inline-asm statement, so the
condition-code register never has to survive between C++ statements.

This script
  * builds each routine as a list of abstract PTX instructions,
  * RUNS that exact instruction list in a tiny PTX-subset interpreter against
    Python bigints (`--check`, also called from tests/test_field_ptx.py), so the
    sequence is validated on a CPU-only box before it ever reaches a GPU,
  * emits kzg_poly_commit_exploration_b200/csrc/field_ptx_gen.cuh.

Multiplication layout (derived for this engine; the well-known "even/odd
column" arrangement that keeps 64-bit partial products limb-aligned):
  E[k] sits at limb position k, O[k] at position k+1.  Row i adds a_even*b_i into
  E and a_odd*b_i into O, then m = E[0]*(-p^-1), then p_even*m into E and
  p_odd*m into O.  E[0] is now 0; dividing by 2^32 is a ROLE SWAP (O becomes the
  new E) plus a lazy two-limb shift of the old E that is folded into the next
  row's first carry chain.  2N^2+N multiply instructions per product.
"""
from __future__ import annotations

import argparse
import os
import random
import sys

Q = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
M32 = 0xFFFFFFFF


def limbs(x: int, n: int):
    return [(x >> (32 * i)) & M32 for i in range(n)]


class Prog:
    """Straight-line PTX-subset program.  Operands: str = register, int = immediate."""

    def __init__(self, name, n_out, n_in_groups):
        self.name = name
        self.ins = []
        self.regs = {}  # family -> count

    def reg(self, fam, count):
        self.regs[fam] = count

    def e(self, op, dst, *src):
        self.ins.append((op, dst, src))

    # ---- interpreter -------------------------------------------------
    def run(self, env):
        cf = 0
        preds = {}

        def val(x):
            return x if isinstance(x, int) else env[x]

        for op, dst, src in self.ins:
            s = [val(x) for x in src if not (isinstance(x, str) and x.startswith("pq"))]
            if op == "mov":
                env[dst] = s[0]
            elif op == "mul.lo":
                env[dst] = (s[0] * s[1]) & M32
            elif op == "mul.hi":
                env[dst] = (s[0] * s[1]) >> 32
            elif op in ("add.cc", "addc.cc", "addc", "add"):
                t = s[0] + s[1] + (cf if op.startswith("addc") else 0)
                env[dst] = t & M32
                if op.endswith(".cc"):
                    cf = t >> 32
            elif op in ("sub.cc", "subc.cc", "subc", "sub"):
                t = s[0] - s[1] - (cf if op.startswith("subc") else 0)
                env[dst] = t & M32
                if op.endswith(".cc"):
                    cf = 1 if t < 0 else 0
            elif op in ("mad.lo.cc", "madc.lo.cc", "madc.lo", "mad.lo"):
                t = ((s[0] * s[1]) & M32) + s[2] + (cf if op.startswith("madc") else 0)
                env[dst] = t & M32
                if op.endswith(".cc"):
                    cf = t >> 32
            elif op in ("mad.hi.cc", "madc.hi.cc", "madc.hi", "mad.hi"):
                t = ((s[0] * s[1]) >> 32) + s[2] + (cf if op.startswith("madc") else 0)
                env[dst] = t & M32
                if op.endswith(".cc"):
                    cf = t >> 32
            elif op == "and":
                env[dst] = s[0] & s[1]
            elif op == "setp.ne":
                preds[dst] = s[0] != s[1]
            elif op == "selp":
                env[dst] = s[0] if preds[src[2]] else s[1]
            else:
                raise ValueError(op)
        return env

    # ---- emission ----------------------------------------------------
    def ptx_lines(self):
        out = []
        for fam, cnt in self.regs.items():
            if fam == "pq":
                out.append(f".reg .pred pq<{cnt}>;")
            else:
                out.append(f".reg .u32 {fam}<{cnt}>;")

        def fmt(x):
            if isinstance(x, int):
                return f"0x{x:08x}"
            return x

        for op, dst, src in self.ins:
            if op == "setp.ne":
                out.append(f"setp.ne.u32 {dst}, {fmt(src[0])}, {fmt(src[1])};")
            elif op == "selp":
                out.append(f"selp.u32 {fmt(dst)}, {fmt(src[0])}, {fmt(src[1])}, {src[2]};")
            elif op == "and":
                out.append(f"and.b32 {fmt(dst)}, {fmt(src[0])}, {fmt(src[1])};")
            elif op == "mov":
                out.append(f"mov.u32 {fmt(dst)}, {fmt(src[0])};")
            else:
                out.append(f"{op}.u32 {fmt(dst)}, " + ", ".join(fmt(x) for x in src) + ";")
        return out


def arg(group, k):
    """Symbolic C operand: group 'o' outputs, 'a','b' inputs."""
    return f"${group}{k}"


# ----------------------------------------------------------------------
# Routines
# ----------------------------------------------------------------------
def cond_sub_p(P, n, p_l, src_fam, out_group):
    """out = src - p if src >= p else src  (src < 2p)."""
    P.reg("u", n)
    P.reg("pq", 1)
    P.reg("bm", 1)
    for k in range(n):
        P.e("sub.cc" if k == 0 else "subc.cc", f"u{k}", f"{src_fam}{k}", p_l[k])
    P.e("subc", "bm0", 0, 0)
    P.e("setp.ne", "pq0", "bm0", 0)  # borrow -> keep src
    for k in range(n):
        P.e("selp", arg(out_group, k), f"{src_fam}{k}", f"u{k}", "pq0")


def gen_mont_mul(n, p, square=False):
    p_l = limbs(p, n)
    inv = (-pow(p, -1, 1 << 32)) & M32
    P = Prog("mul", n, 2)
    P.reg("x", n)  # role array X
    P.reg("y", n)  # role array Y
    P.reg("m", 1)
    P.reg("r", n)
    A = [arg("a", k) for k in range(n)]
    B = [arg("a" if square else "b", k) for k in range(n)]
    E, O = "x", "y"
    for i in range(n):
        bi = B[i]
        if i == 0:
            for j in range(0, n, 2):
                P.e("mul.lo", f"{O}{j}", A[j + 1], bi)
                P.e("mul.hi", f"{O}{j+1}", A[j + 1], bi)
            for j in range(0, n, 2):
                P.e("mul.lo", f"{E}{j}", A[j], bi)
                P.e("mul.hi", f"{E}{j+1}", A[j], bi)
        else:
            # roles were swapped at the end of the previous row: E is the old O (already aligned),
            # O is the old E, still at its old position -> fold old E[1] into E[0], shift by two lazily.
            P.e("add.cc", f"{E}0", f"{E}0", f"{O}1")
            for j in range(0, n - 2, 2):
                P.e("madc.lo.cc", f"{O}{j}", A[j + 1], bi, f"{O}{j+2}")
                P.e("madc.hi.cc", f"{O}{j+1}", A[j + 1], bi, f"{O}{j+3}")
            P.e("madc.lo.cc", f"{O}{n-2}", A[n - 1], bi, 0)
            P.e("madc.hi", f"{O}{n-1}", A[n - 1], bi, 0)
            for j in range(0, n, 2):
                P.e("mad.lo.cc" if j == 0 else "madc.lo.cc", f"{E}{j}", A[j], bi, f"{E}{j}")
                P.e("madc.hi.cc", f"{E}{j+1}", A[j], bi, f"{E}{j+1}")
            P.e("addc", f"{O}{n-1}", f"{O}{n-1}", 0)
        P.e("mul.lo", "m0", f"{E}0", inv)
        for j in range(0, n, 2):
            P.e("mad.lo.cc" if j == 0 else "madc.lo.cc", f"{O}{j}", p_l[j + 1], "m0", f"{O}{j}")
            P.e("madc.hi.cc", f"{O}{j+1}", p_l[j + 1], "m0", f"{O}{j+1}")
        for j in range(0, n, 2):
            P.e("mad.lo.cc" if j == 0 else "madc.lo.cc", f"{E}{j}", p_l[j], "m0", f"{E}{j}")
            P.e("madc.hi.cc", f"{E}{j+1}", p_l[j], "m0", f"{E}{j+1}")
        P.e("addc", f"{O}{n-1}", f"{O}{n-1}", 0)
        E, O = O, E
    # after the swap: E is the aligned array (old O), O is the old E with O[0] == 0.
    for k in range(n):
        op = "add.cc" if k == 0 else ("addc.cc" if k < n - 1 else "addc")
        P.e(op, f"r{k}", f"{E}{k}", f"{O}{k+1}" if k + 1 < n else 0)
    cond_sub_p(P, n, p_l, "r", "o")
    return P


def gen_add_mod(n, p):
    p_l = limbs(p, n)
    P = Prog("add", n, 2)
    P.reg("t", n)
    for k in range(n):
        op = "add.cc" if k == 0 else ("addc.cc" if k < n - 1 else "addc")
        P.e(op, f"t{k}", arg("a", k), arg("b", k))
    cond_sub_p(P, n, p_l, "t", "o")
    return P


def gen_sub_mod(n, p):
    p_l = limbs(p, n)
    P = Prog("sub", n, 2)
    P.reg("t", n)
    P.reg("bm", 1)
    P.reg("pm", n)
    for k in range(n):
        P.e("sub.cc" if k == 0 else "subc.cc", f"t{k}", arg("a", k), arg("b", k))
    P.e("subc", "bm0", 0, 0)
    for k in range(n):
        P.e("and", f"pm{k}", "bm0", p_l[k])
    for k in range(n):
        op = "add.cc" if k == 0 else ("addc.cc" if k < n - 1 else "addc")
        P.e(op, arg("o", k), f"t{k}", f"pm{k}")
    return P


def gen_reduce_once(n, p):
    """out = a - p if a >= p else a (a < 2p)."""
    p_l = limbs(p, n)
    P = Prog("red", n, 1)
    P.reg("t", n)
    for k in range(n):
        P.e("mov", f"t{k}", arg("a", k))
    cond_sub_p(P, n, p_l, "t", "o")
    return P


FIELDS = {"fq": (12, Q), "fr": (8, R)}
ROUTINES = {
    "mul": (gen_mont_mul, 2),
    "sqr": (lambda n, p: gen_mont_mul(n, p, square=True), 1),
    "add": (gen_add_mod, 2),
    "sub": (gen_sub_mod, 2),
    "red": (gen_reduce_once, 1),
}


# ----------------------------------------------------------------------
# CPU verification of the instruction lists
# ----------------------------------------------------------------------
def run_routine(P, n, a, b=None):
    env = {}
    for k in range(n):
        env[arg("a", k)] = (a >> (32 * k)) & M32
        if b is not None:
            env[arg("b", k)] = (b >> (32 * k)) & M32
    P.run(env)
    return sum(env[arg("o", k)] << (32 * k) for k in range(n))


def check(iters=300, seed=1):
    rnd = random.Random(seed)
    for fname, (n, p) in FIELDS.items():
        Rm = 1 << (32 * n)
        Rinv = pow(Rm, -1, p)
        progs = {r: g(n, p) for r, (g, _) in ROUTINES.items()}
        edge = [0, 1, 2, p - 1, p - 2, (p - 1) // 2, (1 << (32 * n - 32)) % p, Rm % p, (Rm * Rm) % p]
        vals = edge + [rnd.randrange(p) for _ in range(iters)]
        for idx, a in enumerate(vals):
            b = vals[(idx * 7 + 3) % len(vals)] if idx < len(edge) * 2 else rnd.randrange(p)
            for bb in ([b] + (edge if idx < len(edge) else [])):
                assert run_routine(progs["mul"], n, a, bb) == a * bb * Rinv % p, (fname, "mul", a, bb)
                assert run_routine(progs["add"], n, a, bb) == (a + bb) % p, (fname, "add")
                assert run_routine(progs["sub"], n, a, bb) == (a - bb) % p, (fname, "sub")
            assert run_routine(progs["sqr"], n, a) == a * a * Rinv % p, (fname, "sqr")
            assert run_routine(progs["red"], n, a) == a, (fname, "red")
            assert run_routine(progs["red"], n, a + p) == a, (fname, "red2")
        # Montgomery multiplication must also accept an unreduced (< 2^(32n)) second operand
        # when the first is reduced (used for "bytes -> Montgomery" loads): result is reduced.
        for _ in range(50):
            a = rnd.randrange(p)
            b = rnd.randrange(Rm)
            if (a * b + (Rm - 1) * p) // Rm < 2 * p:  # always true since a < p
                assert run_routine(progs["mul"], n, a, b) == a * b * Rinv % p
    return True


# ----------------------------------------------------------------------
# C++ emission
# ----------------------------------------------------------------------
def emit_cpp():
    out = []
    out.append("// GENERATED by tools/gen_field_ptx.py — do not edit; regenerate with\n"
               "//   python tools/gen_field_ptx.py --emit\n"
               "// Straight-line sm_100a PTX for Montgomery arithmetic in 32-bit limbs.\n"
               "// The instruction lists are verified on the CPU by the same script (--check).\n"
               "#pragma once\n#include <cstdint>\n")
    out.append("#if defined(__CUDA_ARCH__)\n")
    for fname, (n, p) in FIELDS.items():
        for rname, (g, nin) in ROUTINES.items():
            P = g(n, p)
            # operand numbering: outputs 0..n-1, a n..2n-1, b 2n..3n-1
            def sub(tok):
                if isinstance(tok, str) and tok.startswith("$"):
                    grp, k = tok[1], int(tok[2:])
                    base = {"o": 0, "a": n, "b": 2 * n}[grp]
                    return f"%{base + k}"
                return tok

            P2 = Prog(P.name, 0, 0)
            P2.regs = P.regs
            P2.ins = [(op, sub(dst), tuple(sub(s) for s in src)) for op, dst, src in P.ins]
            lines = P2.ptx_lines()
            sig_in = "const uint32_t* a" + (", const uint32_t* b" if nin == 2 else "")
            out.append(f"// {len(P.ins)} PTX instructions\n")
            out.append(f"__device__ __forceinline__ void {fname}_{rname}_ptx(uint32_t* o, {sig_in}) {{\n")
            out.append("  asm(\"{\\n\\t\"\n")
            for ln in lines:
                out.append(f"      \"{ln}\\n\\t\"\n")
            out.append("      \"}\"\n")
            outs = ", ".join(f"\"=r\"(o[{k}])" for k in range(n))
            ins = ", ".join(f"\"r\"(a[{k}])" for k in range(n))
            if nin == 2:
                ins += ", " + ", ".join(f"\"r\"(b[{k}])" for k in range(n))
            out.append(f"      : {outs}\n      : {ins});\n}}\n\n")
    out.append("#endif  // __CUDA_ARCH__\n")
    return "".join(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--emit", action="store_true")
    ap.add_argument("--iters", type=int, default=300)
    args = ap.parse_args()
    if args.check or not args.emit:
        check(args.iters)
        print("field PTX instruction lists verified against bigints (fq, fr: mul sqr add sub red)")
    if args.emit:
        here = os.path.dirname(os.path.abspath(__file__))
        dst = os.path.join(here, "..", "kzg_poly_commit_exploration_b200", "csrc", "field_ptx_gen.cuh")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w") as f:
            f.write(emit_cpp())
        print("wrote", os.path.normpath(dst))


if __name__ == "__main__":
    sys.exit(main())
