"""Experiment: two contexts (own streams + scratch) on one GPU sharing one setup, MSMs submitted alternately so that the
latency-bound tail of one overlaps the bulk kernels of the next.  usage: python tools/exp_two_ctx.py [logn] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import kzg_oracle as O
import kzg_poly_commit_exploration_b200 as pkg
from kzg_poly_commit_exploration_b200 import synthetic

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = 1 << logn
a, b = pkg.Context(0), pkg.Context(0)
st = a.setup_generate(synthetic.BENCH_SECRET_BE, n)
st.prepare_msm(0)
host = torch.from_numpy(synthetic.scalar_stream(0xC3, n)).pin_memory()
ref = a.msm_g1(st, host.data_ptr(), n)
assert b.msm_g1(st, host.data_ptr(), n) == ref

def one_ctx(k):
    prev = None
    for _ in range(k):
        t = a.msm_g1_submit(st, host.data_ptr(), n)
        if prev is not None: assert a.msm_g1_wait(prev) == ref
        prev = t
    assert a.msm_g1_wait(prev) == ref

def two_ctx(k):
    pend = []
    for i in range(k):
        c = (a, b)[i & 1]
        pend.append((c, c.msm_g1_submit(st, host.data_ptr(), n)))
        if len(pend) > 2:
            cc, t = pend.pop(0)
            assert cc.msm_g1_wait(t) == ref
    for cc, t in pend:
        assert cc.msm_g1_wait(t) == ref

for name, fn in (("one ctx, 2 in flight", one_ctx), ("two ctx alternating", two_ctx)):
    fn(4); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(steps); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"n=2^{logn} {name}: {dt / steps * 1e3:.3f} ms per MSM, {n * steps / dt / 1e6:.1f} M points/s", flush=True)
