"""World-size-2 gloo test of the multi-GPU plumbing on the CPU: point-range sharding of the scalar
stream and bases, all-gather of one fixed-size partial blob per rank in rank order, fold on rank 0.
(The device-side partial/fold kernels are covered by tests/test_gpu_parity.py::test_sharded_partials_combine.)"""
import ctypes as C
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL = 301   # deliberately not divisible by the world size


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import kzg_oracle as O
    import kzg_poly_commit_exploration_b200 as pkg
    from kzg_poly_commit_exploration_b200 import parallel, synthetic

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = parallel.shard_range(N_TOTAL, world, rank)
    s = O.fr_from_be_bytes(synthetic.BENCH_SECRET_BE)
    scal = synthetic.scalar_stream(0xC3, count, start=first)
    # rank-local partial commitment over ITS bases [first, first+count): sum c_i [s^i] G (checker arithmetic)
    acc = 0
    for i in range(count):
        acc = (acc + int.from_bytes(scal[i].tobytes(), "little") * pow(s, first + i, O.R)) % O.R
    blob = torch.zeros(192, dtype=torch.uint8)
    blob[:48] = torch.frombuffer(bytearray(O.g1_compress(O.g1_mul(O.G1, acc))), dtype=torch.uint8)
    blob[48] = rank
    allb = parallel.gather_blobs(blob, world)
    if rank == 0:
        L = pkg.lib()
        out = (C.c_uint8 * 48).from_buffer_copy(O.g1_compress(None))
        for g in range(world):
            assert int(allb[192 * g + 48]) == g            # rank-major layout
            part = (C.c_uint8 * 48).from_buffer_copy(bytes(allb[192 * g: 192 * g + 48].tolist()))
            assert L.kzg_g1_add(out, part, out) == 0
        full = synthetic.scalar_stream(0xC3, N_TOTAL)
        co = [int.from_bytes(full[i].tobytes(), "little") for i in range(N_TOTAL)]
        ret.put(bytes(out) == O.g1_compress(O.commit_known_secret(co, s)))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_partition_the_workload():
    from kzg_poly_commit_exploration_b200 import parallel, synthetic

    for n, world in [(1 << 20, 8), (301, 2), (7, 4), (5, 8)]:
        covered = []
        for r in range(world):
            f, c = parallel.shard_range(n, world, r)
            covered += list(range(f, f + c))
        assert covered == list(range(n))
    whole = synthetic.scalar_stream(9, 100)
    parts = [synthetic.scalar_stream(9, c, start=f) for f, c in (parallel.shard_range(100, 3, r) for r in range(3))]
    assert (torch.from_numpy(whole) == torch.cat([torch.from_numpy(p) for p in parts])).all()


def test_two_rank_gather_and_fold_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    ok = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def _handle_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from kzg_poly_commit_exploration_b200 import parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    handle = bytes([rank + 1]) * 64                      # stands for the CUDA IPC handle of rank `rank`'s mailbox
    ret.put((rank, parallel.exchange_handles(handle, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_handle_exchange_gloo():
    """host side of parallel.connect_peers: every rank ends up with all mailbox handles, rank-major"""
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_handle_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(ret.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert got[0] == got[1] == bytes([1]) * 64 + bytes([2]) * 64
