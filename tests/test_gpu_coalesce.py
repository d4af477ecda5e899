"""Calls fused into one bucket problem (option msm_coalesce, csrc/msm.cu msm_bucket_multi, csrc/engine_one.cu
flush_open_group): B commits over the same setup that are in flight together run as ONE MSM with B times the buckets.
Parity bar: every ticket returns the bytes of the one-shot call (and, where the oracle finishes in seconds, of the
known-secret identity), whatever the grouping — full groups, partial groups sent off by a wait, a sync or a call of
another shape, host and device scalars mixed, adversarial scalars, the peer exchange."""
import pytest

import kzg_oracle as O
import util

pytestmark = pytest.mark.gpu

SECRET = O.BENCH_SECRET_BE
S = O.fr_from_be_bytes(SECRET)


def _ctx(lanes, coalesce):
    import kzg_poly_commit_exploration_b200 as pkg
    c = pkg.Context(0)
    c.set_option("lanes", lanes)
    c.set_option("msm_coalesce", coalesce)
    return c


@pytest.mark.parametrize("lanes,coalesce", [(1, 2), (1, 4), (2, 3), (2, 4), (4, 4)])
def test_fused_calls_equal_one_shot_calls(lanes, coalesce):
    from kzg_poly_commit_exploration_b200._native import KzgNativeError
    ctx = _ctx(lanes, coalesce)
    F = lanes * coalesce
    n_setup = 20000
    setup = ctx.setup_generate(SECRET, n_setup)
    # runs of equal lengths (they fuse), changes of length (the open group is sent off), empty sums, the full setup
    sizes = [3000] * 5 + [17] * 2 + [0] * 3 + [20000] * (coalesce + 1) + [1] + [3000] * 2 + [19999] * 3
    sets = [util.scalar_stream_bytes(0x60 + i, n) for i, n in enumerate(sizes)]
    expect = [ctx.msm_g1(setup, s, n) for s, n in zip(sets, sizes)]
    assert expect[0] == O.g1_compress(O.commit_known_secret(util.to_ints(sets[0]), S))
    assert expect[7] == O.g1_compress(None)
    pending = []
    for i in range(len(sets)):
        if len(pending) == F:
            with pytest.raises(KzgNativeError):
                ctx.msm_g1_submit(setup, sets[i], sizes[i])       # every slot of the queue is out
            t, j = pending.pop(0)
            assert ctx.msm_g1_wait(t) == expect[j], (j, sizes[j])
        pending.append((ctx.msm_g1_submit(setup, sets[i], sizes[i]), i))
    with pytest.raises(KzgNativeError):
        ctx.set_option("msm_coalesce", 1)                         # fixed once a ticket is out
    assert ctx.msm_g1(setup, sets[0], sizes[0]) == expect[0]      # the one-shot call still works beside pending tickets
    for t, j in reversed(pending):                                # newest first: the first wait sends the open group off
        assert ctx.msm_g1_wait(t) == expect[j], (j, sizes[j])
    with pytest.raises(KzgNativeError):
        ctx.msm_g1_wait(pending[0][0])                            # already collected
    setup.close()
    ctx.close()


def test_fused_device_scalars_sync_and_mixed_sources():
    import torch
    ctx = _ctx(2, 4)
    n = 1 << 15
    setup = ctx.setup_generate(SECRET, n)
    host = [util.scalar_stream_bytes(0x70 + i, n) for i in range(6)]
    expect = [ctx.msm_g1(setup, h) for h in host]
    devs = [torch.frombuffer(bytearray(h), dtype=torch.uint8).cuda() for h in host]
    torch.cuda.synchronize()
    for rep in range(3):
        # device, host, device, host in ONE group; then a partial group of two that kzg_ctx_sync sends off
        ts = [ctx.msm_g1_submit_dev(setup, devs[0].data_ptr(), n), ctx.msm_g1_submit(setup, host[1]),
              ctx.msm_g1_submit_dev(setup, devs[2].data_ptr(), n), ctx.msm_g1_submit(setup, host[3]),
              ctx.msm_g1_submit_dev(setup, devs[4].data_ptr(), n), ctx.msm_g1_submit(setup, host[5])]
        ctx.sync()
        for k in (5, 0, 3, 1, 4, 2):
            assert ctx.msm_g1_wait(ts[k]) == expect[k], (rep, k)
    setup.close()
    ctx.close()


def test_fused_adversarial_scalars():
    """all-equal scalars (every entry of a call in ONE bucket: the heavy-bucket path inside the fused problem), zeros,
    r - 1, values >= r, and one ordinary call riding in the same group"""
    ctx = _ctx(1, 4)
    n = 40000
    setup = ctx.setup_generate(SECRET, n)
    r = O.R
    same = util.le(0x1234567) * n
    top = util.le(r - 1) * n
    big = b"".join(util.le((r + 5 + i) % (1 << 256)) for i in range(n))       # >= r: reduced by the recoder
    ordinary = util.scalar_stream_bytes(0x77, n)
    sets = [same, top, big, ordinary]
    expect = [ctx.msm_g1(setup, s) for s in sets]
    sum_pow = (pow(S, n, r) - 1) * pow(S - 1, -1, r) % r                     # sum of s^i, i < n
    assert expect[0] == O.g1_compress(O.g1_from_scalar(0x1234567 * sum_pow % r))
    assert expect[1] == O.g1_compress(O.g1_from_scalar((r - 1) * sum_pow % r))
    ts = [ctx.msm_g1_submit(setup, s) for s in sets]
    for k in range(4):
        assert ctx.msm_g1_wait(ts[k]) == expect[k], k
    zeros = bytes(32 * n)
    ts = [ctx.msm_g1_submit(setup, zeros), ctx.msm_g1_submit(setup, ordinary), ctx.msm_g1_submit(setup, zeros)]
    assert ctx.msm_g1_wait(ts[1]) == expect[3]
    assert ctx.msm_g1_wait(ts[0]) == O.g1_compress(None)
    assert ctx.msm_g1_wait(ts[2]) == O.g1_compress(None)
    setup.close()
    ctx.close()


@pytest.mark.parametrize("window", [0, 12, 16])
def test_fused_large_share(window):
    """n = 2^17, the per-GPU share of the 8-GPU MSM (BASELINE config C3), four calls fused, two groups in flight"""
    import torch
    ctx = _ctx(2, 4)
    if window:
        ctx.set_option("msm_window_bits", window)
    n = 1 << 17
    setup = ctx.setup_generate(SECRET, n)
    host = [util.scalar_stream_bytes(0xC3, n, start=i * n) for i in range(8)]
    devs = [torch.frombuffer(bytearray(h), dtype=torch.uint8).cuda() for h in host]
    torch.cuda.synchronize()
    one = _ctx(1, 1)
    s1 = one.setup_generate(SECRET, n)
    expect = [one.msm_g1(s1, h) for h in host]
    assert expect[0] == O.g1_compress(O.commit_known_secret(util.to_ints(host[0]), S))
    ts = [ctx.msm_g1_submit_dev(setup, d.data_ptr(), n) for d in devs]
    for k in range(8):
        assert ctx.msm_g1_wait(ts[k]) == expect[k], k
    s1.close(); one.close()
    setup.close()
    ctx.close()


def test_fused_peer_exchange_single_rank():
    """collective calls fused: consecutive sequence numbers, one block of k_exchange_combine per call"""
    import torch
    ctx = _ctx(2, 4)
    n = 6000
    setup = ctx.setup_generate(SECRET, n)
    ctx.p2p_connect(ctx.p2p_init(1, 0))
    host = [util.scalar_stream_bytes(0x90 + i, n) for i in range(11)]
    expect = [ctx.msm_g1(setup, h) for h in host]
    devs = [torch.frombuffer(bytearray(h), dtype=torch.uint8).cuda() for h in host]
    torch.cuda.synchronize()
    ts = []
    for k in range(11):                        # 8 in flight at most: two full groups, then a partial one of three
        if len(ts) == 8:
            t, j = ts.pop(0)
            assert ctx.msm_g1_wait(t) == expect[j], j
        ts.append((ctx.msm_g1_sharded_submit_dev(setup, devs[k].data_ptr(), n), k))
    for t, j in ts:
        assert ctx.msm_g1_wait(t) == expect[j], j
    # a plain call between collective ones does not disturb their numbering
    a = ctx.msm_g1_sharded_submit(setup, host[0])
    b = ctx.msm_g1_submit(setup, host[1])
    c = ctx.msm_g1_sharded_submit(setup, host[2])
    assert ctx.msm_g1_wait(c) == expect[2]
    assert ctx.msm_g1_wait(a) == expect[0]
    assert ctx.msm_g1_wait(b) == expect[1]
    setup.close()
    ctx.close()


def test_fused_calls_on_a_group_context(native):
    """single-process multi-device context: members that share a device run unfused (they allocate in a phase of their
    own), members on distinct devices fuse — either way the bytes are the single-device ones"""
    import torch
    ndev = torch.cuda.device_count()
    devs = list(range(ndev)) if ndev >= 2 else [0, 0]
    g = native.Context(devs)
    g.set_option("msm_coalesce", 4)
    g.set_option("setup_layout", 1)
    n = 50000
    sg = g.setup_generate(SECRET, n)
    host = [util.scalar_stream_bytes(0xA0 + i, n) for i in range(10)]
    one = native.Context(0)
    s1 = one.setup_generate(SECRET, n)
    expect = [one.msm_g1(s1, h) for h in host]
    cap = 4 * (4 if ndev >= 2 else 1)
    pend = []
    for k in range(10):
        if len(pend) == min(cap, 8):
            t, j = pend.pop(0)
            assert g.msm_g1_wait(t) == expect[j], j
        pend.append((g.msm_g1_submit(sg, host[k]), k))
    g.sync()
    for t, j in pend:
        assert g.msm_g1_wait(t) == expect[j], j
    s1.close(); one.close()
    sg.close()
    g.close()


@pytest.mark.parametrize("window,coalesce,n", [(12, 1, 30000), (17, 2, 1 << 16), (19, 4, 1 << 17), (20, 2, 1 << 17), (20, 4, 1 << 15)])
def test_second_level_bucket_reduction(window, coalesce, n):
    """option msm_l2 (msm.cu finish_level2; automatic from 18-bit windows on): the bucket sums of level 1 are reduced as a
    second bucket problem (digit = hi * 2^s + lo).  Same bytes as the running-sum tail and as the known-secret identity,
    alone and fused, uniform and adversarial scalars (one bucket per window: the heavy path of both levels)."""
    import torch
    ref_ctx = _ctx(1, 1)
    s_ref = ref_ctx.setup_generate(SECRET, n)
    r = O.R
    sets = [util.scalar_stream_bytes(0xB0 + i, n) for i in range(3)]
    sets.append(util.le(0x0123456789ABCDEF0123456789ABCDEF % r) * n)             # all equal
    sets.append(b"".join(util.le((r - 1 - i) % r) for i in range(n)))            # just below r: folded to small values
    expect = [ref_ctx.msm_g1(s_ref, s) for s in sets]
    assert expect[0] == O.g1_compress(O.commit_known_secret(util.to_ints(sets[0]), S))
    ctx = _ctx(2, coalesce)
    ctx.set_option("msm_window_bits", window)
    ctx.set_option("msm_l2", 1)
    setup = ctx.setup_generate(SECRET, n)
    for k, s in enumerate(sets):
        assert ctx.msm_g1(setup, s) == expect[k], ("one-shot", k)
    devs = [torch.frombuffer(bytearray(s), dtype=torch.uint8).cuda() for s in sets]
    torch.cuda.synchronize()
    ts = [ctx.msm_g1_submit_dev(setup, d.data_ptr(), n) for d in devs[:2 * coalesce]]
    for k, t in enumerate(ts):
        assert ctx.msm_g1_wait(t) == expect[k], ("in flight", k)
    ts = [ctx.msm_g1_submit(setup, s) for s in (sets[3], sets[0], sets[4], sets[3])[:2 * coalesce]]
    want = [expect[3], expect[0], expect[4], expect[3]]
    for k, t in enumerate(ts):
        assert ctx.msm_g1_wait(t) == want[k], ("adversarial in flight", k)
    setup.close(); ctx.close()
    s_ref.close(); ref_ctx.close()


def test_batches_of_long_rows_run_fused():
    """kzg_commit_batch / kzg_open_batch with rows beyond the fixed-base table (long polynomials): rows go through the
    bucket MSM four at a time as one problem — same bytes as one commit / open per row (src/polynomial.rs:200-215, 260-269)"""
    ctx = _ctx(1, 1)
    n, rows = 9000, 7
    setup = ctx.setup_generate(SECRET, n)
    coeffs = util.scalar_stream_bytes(0xD1, rows * n)
    zs = util.scalar_stream_bytes(0xD2, rows)
    one = [ctx.commit(setup, coeffs[r * n * 32:(r + 1) * n * 32]) for r in range(rows)]
    assert one[0] == O.g1_compress(O.commit_known_secret(util.to_ints(coeffs[:n * 32]), S))
    assert ctx.commit_batch(setup, coeffs, rows, n) == b"".join(one)
    ys, proofs = ctx.open_batch(setup, coeffs, rows, n, zs)
    for r in range(rows):
        y, pr = ctx.open(setup, coeffs[r * n * 32:(r + 1) * n * 32], zs[r * 32:(r + 1) * 32])
        assert ys[r * 32:(r + 1) * 32] == y and proofs[r * 48:(r + 1) * 48] == pr, r
    setup.close(); ctx.close()


@pytest.mark.parametrize("sort", [1, 2])
def test_alternative_bucket_sorts(sort):
    """option msm_sort: 1 = cub radix sort (lone calls), 2 = two-level partition sort (shared-memory atomics) — kept for
    comparison with the default counting sort; same bytes, alone and fused, uniform and all-equal scalars"""
    import torch
    n = 1 << 16
    ref = _ctx(1, 1)
    s0 = ref.setup_generate(SECRET, n)
    sets = [util.scalar_stream_bytes(0xE0 + i, n) for i in range(3)] + [util.le(12345678901234567890) * n]
    expect = [ref.msm_g1(s0, s) for s in sets]
    ctx = _ctx(2, 2)
    ctx.set_option("msm_sort", sort)
    ctx.set_option("msm_window_bits", 14)
    setup = ctx.setup_generate(SECRET, n)
    for k, s in enumerate(sets):
        assert ctx.msm_g1(setup, s) == expect[k], k
    ts = [ctx.msm_g1_submit(setup, s) for s in sets]
    for k, t in enumerate(ts):
        assert ctx.msm_g1_wait(t) == expect[k], k
    setup.close(); ctx.close(); s0.close(); ref.close()
