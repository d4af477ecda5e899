"""G2 setup timing: XYZZ walk (setup_tree_min=0) against the affine trees over Fq2.  usage: prof_setup_g2.py [logn]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import kzg_oracle as O
import kzg_poly_commit_exploration_b200 as pkg
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n = 1 << logn
ctx = pkg.Context(0)
s = O.fr_from_be_bytes(O.BENCH_SECRET_BE)
st = ctx.setup_generate(O.BENCH_SECRET_BE, n)
st.generate_g2(O.BENCH_SECRET_BE, 2)          # builds the table
for mode, val in (("xyzz walk", 0), ("affine trees", 1)):
    ctx.set_option("setup_tree_min", val)
    ts = []
    for rep in range(3):
        ctx.sync(); t0 = time.perf_counter()
        st.generate_g2(O.BENCH_SECRET_BE, n)
        ctx.sync(); ts.append(time.perf_counter() - t0)
    ok = st.export_g2(n - 1, 1) == O.g2_compress(O.g2_from_scalar(pow(s, n - 1, O.R)))
    print(f"G2 n=2^{logn} {mode}: " + " ".join(f"{t*1e3:.2f}" for t in ts) + f" ms  ({n / min(ts) / 1e6:.2f} M points/s) ok={ok}", flush=True)
st.close(); ctx.close()
