"""Short driver for ncu: `groups` groups of `B` fused commit-MSMs (option msm_coalesce) at n = 2^logn, checked with the
known-secret identity.  usage: python tools/prof_msm_fused.py [logn] [B] [groups] [name=value ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import torch
import kzg_oracle as O
import util
import kzg_poly_commit_exploration_b200 as pkg

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 17
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
groups = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n = 1 << logn
ctx = pkg.Context(0)
ctx.set_option("lanes", 1)
ctx.set_option("msm_coalesce", B)
for kv in sys.argv[4:]:
    k, v = kv.split("="); ctx.set_option(k, int(v))
st = ctx.setup_generate(O.BENCH_SECRET_BE, n)
st.prepare_msm(0)
scal = util.scalar_stream_bytes(0xC3, n)
exp = O.g1_compress(O.commit_known_secret(util.to_ints(scal), O.fr_from_be_bytes(O.BENCH_SECRET_BE)))
dev = torch.frombuffer(bytearray(scal), dtype=torch.uint8).cuda()
torch.cuda.synchronize()
for g in range(groups):
    ts = [ctx.msm_g1_submit_dev(st, dev.data_ptr(), n) for _ in range(B)]
    for t in ts:
        assert ctx.msm_g1_wait(t) == exp
    print(f"group {g}: acc {ctx.last_kernel_ms(0):.3f} ms dev {ctx.last_kernel_ms(1):.3f} ms")
st.close(); ctx.close()
