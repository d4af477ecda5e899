"""Short driver for ncu: one 4096-point setup with the 15-bit fixed-base table, `reps` commit+open batches of `blobs`
blobs x 4096 coefficients on the device.  usage: python tools/prof_blob.py [blobs] [reps] [window_bits] [name=value ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import kzg_poly_commit_exploration_b200 as pkg
from kzg_poly_commit_exploration_b200 import synthetic
blobs = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
c = int(sys.argv[3]) if len(sys.argv) > 3 else 15
n = 4096
ctx = pkg.Context(0)
for kv in sys.argv[4:]:
    k, v = kv.split("="); ctx.set_option(k, int(v))
st = ctx.setup_generate(synthetic.BENCH_SECRET_BE, n)
st.prepare_fixed(n, c)
dev = torch.device("cuda:0")
coeffs = torch.from_numpy(synthetic.scalar_stream(0xC4, blobs * n)).to(dev)
zs = torch.from_numpy(synthetic.scalar_stream(0xC4F, blobs)).to(dev)
cm = torch.zeros(blobs * 48, dtype=torch.uint8, device=dev); pr = torch.zeros_like(cm)
ys = torch.zeros(blobs * 32, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
for r in range(reps):
    ctx.sync(); t0 = time.perf_counter()
    ctx.commit_open_batch_dev(st, coeffs.data_ptr(), blobs, n, zs.data_ptr(), cm.data_ptr(), ys.data_ptr(), pr.data_ptr())
    ctx.sync(); dt = time.perf_counter() - t0
    import hashlib
    h = hashlib.sha256(cm.cpu().numpy().tobytes() + pr.cpu().numpy().tobytes() + ys.cpu().numpy().tobytes()).hexdigest()[:16]
    print(f"rep {r}: {dt*1e3:.2f} ms  {blobs / dt:.0f} blobs/s  r1_add {ctx.last_kernel_ms(2):.3f} ms  outputs {h}", flush=True)
st.close(); ctx.close()
