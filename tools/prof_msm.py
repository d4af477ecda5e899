"""Short driver for ncu: one setup (n = 2^logn), a few commit-MSMs, result checked with the known-secret
identity.  usage: python tools/prof_msm.py [logn] [reps] [window_bits]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import kzg_oracle as O
import util
import kzg_poly_commit_exploration_b200 as pkg

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n = 1 << logn
ctx = pkg.Context(0)
st = ctx.setup_generate(O.BENCH_SECRET_BE, n)
st.prepare_msm(c)
scal = util.scalar_stream_bytes(0xC3, n)
exp = O.g1_compress(O.commit_known_secret(util.to_ints(scal), O.fr_from_be_bytes(O.BENCH_SECRET_BE)))
for r in range(reps):
    t = time.time(); got = ctx.msm_g1(st, scal); dt = time.time() - t
    print(f"rep {r}: e2e {dt*1e3:.2f} ms acc {ctx.last_kernel_ms(0):.3f} ms dev {ctx.last_kernel_ms(1):.3f} ms ok={got == exp}")
    assert got == exp
st.close(); ctx.close()
