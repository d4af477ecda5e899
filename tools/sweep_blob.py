"""Tuning sweep (not a benchmark): fixed-base window size for the blob batch (commit + proof), 512 blobs x 4096."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import kzg_poly_commit_exploration_b200 as pkg
from kzg_poly_commit_exploration_b200 import synthetic

n, blobs = 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 512
cs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [10, 12, 13, 14]
dev = torch.device("cuda:0")
ctx = pkg.Context(0)
setup = ctx.setup_generate(synthetic.BENCH_SECRET_BE, n)
coeffs = torch.from_numpy(synthetic.scalar_stream(0xC4, blobs * n)).to(dev)
zs = torch.from_numpy(synthetic.scalar_stream(0xC4F, blobs)).to(dev)
commits = torch.zeros(blobs * 48, dtype=torch.uint8, device=dev)
proofs = torch.zeros(blobs * 48, dtype=torch.uint8, device=dev)
ys = torch.zeros(blobs * 32, dtype=torch.uint8, device=dev)
ref = None
for c in cs:
    t0 = time.perf_counter(); setup.prepare_fixed(n, c); ctx.sync(); tb = time.perf_counter() - t0
    def step():
        ctx.commit_batch_dev(setup, coeffs.data_ptr(), blobs, n, commits.data_ptr())
        ctx.open_batch_dev(setup, coeffs.data_ptr(), blobs, n, zs.data_ptr(), ys.data_ptr(), proofs.data_ptr())
        ctx.sync()
    step()
    t0 = time.perf_counter(); step(); step(); ms = (time.perf_counter() - t0) * 500
    out = bytes(commits.cpu().numpy().tobytes()) + bytes(proofs.cpu().numpy().tobytes())
    if ref is None: ref = out
    W = (255 + c - 1) // c
    print(f"c={c} W={W} table {n*W*(1<<(c-1))*96/2**30:.1f} GiB built in {tb:.2f}s: {ms:.1f} ms per {blobs} blobs = {blobs/ms*1e3:.0f} blobs/s same={out == ref}", flush=True)
