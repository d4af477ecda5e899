"""GPU bring-up / tuning script: probes + timing sweeps (not a benchmark)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import kzg_oracle as O
import util
import kzg_poly_commit_exploration_b200 as pkg

ctx = pkg.Context(0)
print("imad wide-MAC/s %.3e" % ctx.probe_imad_peak())
print("fq_mul/s %.3e" % ctx.probe_fq_mul())
S = O.fr_from_be_bytes(O.BENCH_SECRET_BE)
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [16, 18, 20]
splits = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [-1]
n = 1 << logn
t = time.time(); st = ctx.setup_generate(O.BENCH_SECRET_BE, n); ctx.sync(); print(f"n=2^{logn} setup_generate {time.time()-t:.3f}s")
scal = util.scalar_stream_bytes(0xC3, n)
exp = O.g1_compress(O.commit_known_secret(util.to_ints(scal), S))
for c in cs:
    t = time.time(); st.prepare_msm(c); ctx.sync(); print(f"  c={c} prepare_msm {time.time()-t:.3f}s")
    for sp in splits:
        ctx.set_option("acc_split", sp)
        best = None
        for rep in range(4):
            t = time.time(); got = ctx.msm_g1(st, scal); dt = time.time() - t
            r = (dt * 1e3, ctx.last_kernel_ms(0), ctx.last_kernel_ms(1), got == exp)
            if rep and (best is None or r[2] < best[2]): best = r
        print(f"  c={c} split={sp}: e2e {best[0]:.2f} ms  acc {best[1]:.3f} ms  total-dev {best[2]:.3f} ms ok={best[3]}")
st.close()
