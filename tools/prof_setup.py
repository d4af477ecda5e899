"""Setup generation timing: XYZZ walk (setup_tree_min=0) against the affine-tree path.  usage: prof_setup.py [logn]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import kzg_oracle as O
import kzg_poly_commit_exploration_b200 as pkg
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 22
n = 1 << logn
ctx = pkg.Context(0)
s = O.fr_from_be_bytes(O.BENCH_SECRET_BE)
for mode, val in (("xyzz walk", 0), ("affine tree", 1)):
    ctx.set_option("setup_tree_min", val)
    ts = []
    for rep in range(4):
        ctx.sync(); t0 = time.perf_counter()
        st = ctx.setup_generate(O.BENCH_SECRET_BE, n)
        ctx.sync(); ts.append(time.perf_counter() - t0)
        if rep == 3:
            pts = st.export_g1(n - 2, 2)
            ok = pts[48:] == O.g1_compress(O.g1_mul(O.G1, pow(s, n - 1, O.R)))
        st.close()
    print(f"n=2^{logn} {mode}: " + " ".join(f"{t*1e3:.2f}" for t in ts) + f" ms  ({n / min(ts[1:]) / 1e6:.1f} M points/s) ok={ok}", flush=True)
ctx.close()
