"""prof_msm with options: python tools/prof_msm_opts.py logn reps name=value ..."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import kzg_oracle as O
import util
import kzg_poly_commit_exploration_b200 as pkg
logn = int(sys.argv[1]); reps = int(sys.argv[2])
n = 1 << logn
ctx = pkg.Context(0)
for kv in sys.argv[3:]:
    k, v = kv.split("="); ctx.set_option(k, int(v))
st = ctx.setup_generate(O.BENCH_SECRET_BE, n)
st.prepare_msm(0)
scal = util.scalar_stream_bytes(0xC3, n)
exp = O.g1_compress(O.commit_known_secret(util.to_ints(scal), O.fr_from_be_bytes(O.BENCH_SECRET_BE)))
for r in range(reps):
    got = ctx.msm_g1(st, scal)
    print(f"rep {r}: acc {ctx.last_kernel_ms(0):.3f} ms dev {ctx.last_kernel_ms(1):.3f} ms ok={got == exp}")
