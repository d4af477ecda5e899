"""Tuning sweep for the accumulate kernel (not a benchmark): acc_minblocks x acc_split at n = 2^20."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import kzg_oracle as O
import util
import kzg_poly_commit_exploration_b200 as pkg

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << logn
ctx = pkg.Context(0)
st = ctx.setup_generate(O.BENCH_SECRET_BE, n)
st.prepare_msm(16)
scal = util.scalar_stream_bytes(0xC3, n)
exp = O.g1_compress(O.commit_known_secret(util.to_ints(scal), O.fr_from_be_bytes(O.BENCH_SECRET_BE)))
cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16]
for mb in cs:
    st.prepare_msm(mb)
    for sp in (-1, 1, 2, 3, 4):
        ctx.set_option("acc_split", sp)
        best = None
        for rep in range(8):
            got = ctx.msm_g1(st, scal)
            r = (ctx.last_kernel_ms(0), ctx.last_kernel_ms(1), got == exp)
            if rep and (best is None or r[0] < best[0]): best = r
        print(f"c={mb} acc_split={sp}: acc {best[0]:.3f} ms  total-dev {best[1]:.3f} ms  dominant kernel {ctx.last_kernel_ms(2):.3f} ms ok={best[2]}", flush=True)
