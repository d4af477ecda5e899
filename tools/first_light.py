"""GPU bring-up script: probes + timing of each stage at a few sizes (not a benchmark)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import kzg_oracle as O
import util
import kzg_poly_commit_exploration_b200 as pkg

ctx = pkg.Context(0)
print("imad wide-MAC/s", ctx.probe_imad_peak())
print("fq_mul/s", ctx.probe_fq_mul())
S = O.fr_from_be_bytes(O.BENCH_SECRET_BE)
for logn in [int(a) for a in sys.argv[1:]] or [16, 20]:
    n = 1 << logn
    t = time.time(); st = ctx.setup_generate(O.BENCH_SECRET_BE, n); ctx.sync(); print(f"n=2^{logn} setup_generate {time.time()-t:.3f}s")
    for c in ([14, 16] if logn <= 16 else [16, 18, 20]):
        t = time.time(); st.prepare_msm(c); ctx.sync(); print(f"  c={c} prepare_msm {time.time()-t:.3f}s")
        scal = util.scalar_stream_bytes(0xC3, n)
        exp = O.g1_compress(O.commit_known_secret(util.to_ints(scal), S))
        for rep in range(3):
            t = time.time(); got = ctx.msm_g1(st, scal); dt = time.time() - t
            print(f"  c={c} msm e2e {dt*1e3:.2f} ms  acc {ctx.last_kernel_ms(0):.3f} ms total-dev {ctx.last_kernel_ms(1):.3f} ms ok={got==exp}")
    st.close()
