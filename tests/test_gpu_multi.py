"""Two-GPU test of the peer-memory exchange (BASELINE config C3 sharding): one process per GPU, point-range shards,
kzg_msm_g1_sharded_dev on both ranks == the single-GPU commitment == the known-secret identity.  Skipped on boxes
with fewer than two GPUs (the single-GPU suite covers the kernel with world = 1 and a dead peer)."""
import os
import socket
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, coalesce, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    import kzg_poly_commit_exploration_b200 as pkg
    from kzg_poly_commit_exploration_b200 import parallel, synthetic

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)      # only carries the 64-byte IPC handles
    torch.cuda.set_device(rank)
    ctx = pkg.Context(rank)
    ctx.set_option("msm_coalesce", coalesce)      # > 1: consecutive collective calls run as one bucket problem, one exchange launch
    depth = 4 * coalesce
    parallel.connect_peers(ctx, world, rank)
    first, count = parallel.shard_range(n_total, world, rank)
    setup = ctx.setup_generate(synthetic.BENCH_SECRET_BE, count, first=first)
    scal = torch.from_numpy(synthetic.scalar_stream(0xC3, count, start=first)).to(torch.device(f"cuda:{rank}"))
    outs = [ctx.msm_g1_sharded_dev(setup, scal.data_ptr(), count) for _ in range(5)]   # every mailbox slot, back to back
    # `depth` collective calls in flight (four lanes x fused calls), 19 + 2 * depth calls: the mailboxes (2 x 16 deep) wrap
    from collections import deque
    q = deque()
    for _ in range(19 + 2 * depth):
        if len(q) == depth:
            outs.append(ctx.msm_g1_wait(q.popleft()))
        q.append(ctx.msm_g1_sharded_submit_dev(setup, scal.data_ptr(), count))
    while q:
        outs.append(ctx.msm_g1_wait(q.popleft()))
    # the same from pinned host buffers (copy on the lane's copy stream)
    host = torch.from_numpy(synthetic.scalar_stream(0xC3, count, start=first)).pin_memory()
    for _ in range(6):
        if len(q) == depth:
            outs.append(ctx.msm_g1_wait(q.popleft()))
        q.append(ctx.msm_g1_sharded_submit(setup, host.data_ptr(), count))
    while q:
        outs.append(ctx.msm_g1_wait(q.popleft()))
    ret.put((rank, [o.hex() for o in outs]))
    dist.barrier()
    setup.close(); ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("coalesce", [1, 4])
def test_two_gpu_peer_exchange_matches_single_gpu(coalesce):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kzg_oracle as O
    from kzg_poly_commit_exploration_b200 import synthetic

    world, n_total = 2, 70001
    mpc = mp.get_context("spawn")
    ret = mpc.Queue()
    port = _free_port()
    procs = [mpc.Process(target=_worker, args=(r, world, port, n_total, coalesce, ret)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(ret.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    full = synthetic.scalar_stream(0xC3, n_total)
    ints = [int.from_bytes(full[i].tobytes(), "little") for i in range(n_total)]
    expect = O.g1_compress(O.commit_known_secret(ints, O.fr_from_be_bytes(synthetic.BENCH_SECRET_BE))).hex()
    for r in range(world):
        assert len(got[r]) == 5 + 19 + 8 * coalesce + 6 and got[r] == [expect] * len(got[r]), r
