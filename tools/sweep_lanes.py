"""Calls in flight (lanes) sweep: device-resident commit-MSMs submitted `lanes` deep, CUDA-event time per MSM.
usage: python tools/sweep_lanes.py logn lanes[,lanes..] [window_bits] [name=value ...]"""
import os, sys
from collections import deque
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import kzg_oracle as O
import util
import kzg_poly_commit_exploration_b200 as pkg
from kzg_poly_commit_exploration_b200 import synthetic

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
lanes_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3, 4]
rest = sys.argv[3:]
window = int(rest.pop(0)) if rest and "=" not in rest[0] else 0
n = 1 << logn
steps = 48 if logn >= 19 else 144
scal = util.scalar_stream_bytes(0xC3, n)
exp = O.g1_compress(O.commit_known_secret(util.to_ints(scal), O.fr_from_be_bytes(O.BENCH_SECRET_BE)))
for lanes in lanes_list:
    ctx = pkg.Context(0)
    ctx.set_option("lanes", lanes)
    depth = lanes
    for kv in rest:
        k, v = kv.split("="); ctx.set_option(k, int(v))
        if k == "msm_coalesce":
            depth = lanes * int(v)                   # tickets that can be out at once
    st = ctx.setup_generate(synthetic.BENCH_SECRET_BE, n)
    st.prepare_msm(window)
    host = torch.from_numpy(synthetic.scalar_stream(0xC3, n)).pin_memory()
    dev = host.cuda()
    torch.cuda.synchronize()
    ref = ctx.msm_g1_dev(st, dev.data_ptr(), n)
    ok = ref == exp                                  # known-secret identity
    ctx.msm_g1_dev(st, dev.data_ptr(), n); ctx.sync()
    single = ctx.last_kernel_ms(1)

    host_s = [0.0]

    def run(k, submit):
        import time
        q = deque()
        for _ in range(k):
            if len(q) == depth:
                assert ctx.msm_g1_wait(q.popleft()) == ref
            t0 = time.perf_counter()
            q.append(submit())
            host_s[0] += time.perf_counter() - t0
        while q:
            assert ctx.msm_g1_wait(q.popleft()) == ref

    res = []
    for name, submit in (("dev", lambda: ctx.msm_g1_submit_dev(st, dev.data_ptr(), n)),
                         ("e2e", lambda: ctx.msm_g1_submit(st, host.data_ptr(), n))):
        run(2 * depth + 2, submit)
        ctx.sync(); ctx.timer_start()
        host_s[0] = 0.0
        run(steps, submit)
        ms = ctx.timer_stop() / steps
        res.append(f"{name} {ms:.3f} ms ({n / ms / 1e3:.1f} M points/s, host {host_s[0] / steps * 1e3:.3f} ms in submit)")
    print(f"n=2^{logn} c={window or 'auto'} {' '.join(rest)} lanes={lanes}: one call {single:.3f} ms  " + "  ".join(res) + f"  ok={ok}", flush=True)
    del host, dev
    st.close(); ctx.close()
