"""Single-process multi-device contexts (kzg_ctx_create_multi, csrc/abi.cu): ONE host caller, the calls shard internally —
the shape of the reference, which is a single process (src/main.rs:170).  Parity bar: every result equals the
single-device result and the known-secret identities of the oracle, bit for bit.

The group layer is exercised on every box: with one GPU the group lists device 0 twice (two members share the device:
own streams, own mailboxes, the exchange kernel spins beside the other member's kernels); with more GPUs it spans them
all (this is the test that must NOT skip on a multi-GPU box)."""
import ctypes as C
import os
import subprocess
import sys

import pytest

import kzg_oracle as O
import util

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SECRET = O.BENCH_SECRET_BE
S = O.fr_from_be_bytes(SECRET)


def g1c(k):
    return O.g1_compress(O.g1_from_scalar(k % O.R))


@pytest.fixture(scope="module", params=["shared2", "shared3", "all"])
def group(request, native):
    import torch
    n = torch.cuda.device_count()
    devs = {"shared2": [0, 0], "shared3": [0, 0, 0], "all": list(range(n))}[request.param]
    if request.param == "all" and n < 2:
        pytest.skip("one GPU here: the shared-device groups cover the path")
    g = native.Context(devs)
    assert g.device_count == len(devs)
    yield g
    g.close()


def test_group_range_sharded_msm_commit_open(group, ctx):
    G = group.device_count
    n = 30011                                                  # ragged: the shards differ in length
    group.set_option("setup_layout", 1)
    sg = group.setup_generate(SECRET, n)
    s1 = ctx.setup_generate(SECRET, n)
    assert len(sg) == n
    # items across every shard boundary equal the single-device items and the oracle
    assert sg.export_g1(0, n) == s1.export_g1(0, n)
    base = n // G
    for i in (0, 1, base - 1, base, base + 1, n - 1):
        assert sg.export_g1(i, 1) == g1c(pow(S, i, O.R)), i
    scal = util.scalar_stream_bytes(0xC3, n)
    ints = util.to_ints(scal)
    expect = O.g1_compress(O.commit_known_secret(ints, S))
    assert ctx.msm_g1(s1, scal) == expect
    assert group.msm_g1(sg, scal) == expect
    assert group.commit(sg, scal) == expect
    # shorter polynomials: inside member 0's range (no exchange), across a boundary, empty, and one that trims to nothing
    for m in (0, 1, base // 2, base + 7, n - 1):
        assert group.commit(sg, scal[: m * 32]) == O.g1_compress(O.commit_known_secret(ints[:m], S)), m
    assert group.commit(sg, bytes(32 * 5)) == O.g1_compress(None)
    # trimmed length decides (src/polynomial.rs:201-205): n + 3 coefficients, the last three zero
    assert group.commit(sg, scal + bytes(96)) == expect
    with pytest.raises(Exception, match="degree"):
        group.commit(sg, scal + util.le(1))
    # calls in flight: 4 deep, 13 calls (the 8-deep mailboxes wrap)
    from collections import deque
    q, outs = deque(), []
    for _ in range(13):
        if len(q) == 4:
            outs.append(group.msm_g1_wait(q.popleft()))
        q.append(group.msm_g1_submit(sg, scal))
    while q:
        outs.append(group.msm_g1_wait(q.popleft()))
    assert outs == [expect] * 13
    # open: evaluation, proof, divide_by_root's errors (src/polynomial.rs:159-166, 188-192)
    z = O.scalar_stream(0xC2F, 0)
    y, pr = group.open(sg, scal, util.le(z))
    assert int.from_bytes(y, "little") == O.horner(ints, z)
    assert pr == O.g1_compress(O.proof_known_secret(ints, S, z))
    assert (y, pr) == ctx.open(s1, scal, util.le(z))
    with pytest.raises(Exception, match="constant terms do not add up"):
        group.open(sg, scal, util.le(z), util.le((O.horner(ints, z) + 1) % O.R))
    assert group.evaluate(scal, util.le(z)) == y
    # device-resident scalars (on device 0): the other members read their slice over the peer mapping
    import torch
    import numpy as np
    dev = torch.from_numpy(np.frombuffer(scal, dtype=np.uint8).copy()).to("cuda:0")
    assert group.msm_g1_dev(sg, dev.data_ptr(), n) == expect
    t = group.msm_g1_submit_dev(sg, dev.data_ptr(), n)
    assert group.msm_g1_wait(t) == expect
    # G2 half: range-sharded like the G1 items
    sg.generate_g2(SECRET, base + 5)
    assert sg.g2_len() == base + 5
    s1.generate_g2(SECRET, base + 5)
    assert sg.export_g2(0, base + 5) == s1.export_g2(0, base + 5)
    assert sg.export_g2(1, 1) == O.g2_compress(O.g2_from_scalar(S))
    # the one-process-per-GPU entry points refuse a group, and setups do not cross context kinds
    with pytest.raises(Exception, match="shards internally"):
        group.msm_g1_sharded_dev(sg, dev.data_ptr(), n)
    with pytest.raises(Exception, match="another kind of context"):
        ctx.msm_g1(sg, scal)
    with pytest.raises(Exception, match="another kind of context"):
        group.msm_g1(s1, scal)
    with pytest.raises(Exception, match="replicated"):
        sg.prepare_fixed(64, 6)
    sg.close(); s1.close()
    group.set_option("setup_layout", 0)


def test_group_import_roundtrip(group, ctx):
    n = 3001
    group.set_option("setup_layout", 1)
    s1 = ctx.setup_generate(SECRET, n)
    g1 = s1.export_g1(0, n)
    sg = group.setup_import_g1(g1)
    assert sg.export_g1(0, n) == g1
    scal = util.scalar_stream_bytes(0xA1, n)
    assert group.commit(sg, scal) == ctx.commit(s1, scal)
    # a malformed point in the LAST shard (flipping a bit of x lands off-curve about half the time; find one that does)
    k = n - 2
    for flip in range(1, 64):
        bad = bytearray(g1)
        bad[48 * k + 47] ^= flip
        try:
            O.g1_uncompress(bytes(bad[48 * k: 48 * k + 48]))
        except O.KzgError:
            break
    try:
        with pytest.raises(Exception, match="uncompressing"):
            group.setup_import_g1(bytes(bad))
    finally:
        sg.close(); s1.close()
        group.set_option("setup_layout", 0)


def test_group_replicated_blob_batch(group, ctx):
    n, blobs = 512, 1031                                  # > tree_min_rows rows on some members, ragged split
    group.set_option("setup_layout", 0)
    sg = group.setup_generate(SECRET, n)                  # auto layout: replicated
    s1 = ctx.setup_generate(SECRET, n)
    sg.prepare_fixed(n, 10)
    s1.prepare_fixed(n, 10)
    assert sg.fixed_plan()["c"] == 10
    coeffs = util.scalar_stream_bytes(0xC4, blobs * n)
    zs = util.scalar_stream_bytes(0xC4F, blobs)
    ref = ctx.commit_open_batch(s1, coeffs, blobs, n, zs)
    assert group.commit_open_batch(sg, coeffs, blobs, n, zs) == ref
    assert group.commit_batch(sg, coeffs, blobs, n) == ref[0]
    assert group.open_batch(sg, coeffs, blobs, n, zs) == (ref[1], ref[2])
    for b in (0, blobs // group.device_count, blobs - 1):
        co = util.to_ints(coeffs[b * n * 32: (b + 1) * n * 32])
        z = int.from_bytes(zs[32 * b: 32 * b + 32], "little")
        assert ref[0][48 * b: 48 * b + 48] == O.g1_compress(O.commit_known_secret(co, S)), b
        assert ref[2][48 * b: 48 * b + 48] == O.g1_compress(O.proof_known_secret(co, S, z)), b
    # device-resident rows on device 0, outputs on device 0
    import numpy as np
    import torch
    cd = torch.from_numpy(np.frombuffer(coeffs, dtype=np.uint8).copy()).to("cuda:0")
    zd = torch.from_numpy(np.frombuffer(zs, dtype=np.uint8).copy()).to("cuda:0")
    cm = torch.zeros(blobs * 48, dtype=torch.uint8, device="cuda:0")
    ys = torch.zeros(blobs * 32, dtype=torch.uint8, device="cuda:0")
    pr = torch.zeros(blobs * 48, dtype=torch.uint8, device="cuda:0")
    group.commit_open_batch_dev(sg, cd.data_ptr(), blobs, n, zd.data_ptr(), cm.data_ptr(), ys.data_ptr(), pr.data_ptr())
    group.sync()
    assert (cm.cpu().numpy().tobytes(), ys.cpu().numpy().tobytes(), pr.cpu().numpy().tobytes()) == ref
    # a single commitment on a replicated setup: member 0 alone
    assert group.commit(sg, coeffs[: n * 32]) == ref[0][:48]
    sg.close(); s1.close()


def test_group_setup_c5_shape(group):
    """index-range sharded generation through the affine-tree path on every member (>= 2^18 items each)"""
    G = group.device_count
    n = G * (1 << 18) + 77
    group.set_option("setup_layout", 1)
    sg = group.setup_generate(SECRET, n, first=5)
    per = n // G
    for i in (0, 1, per - 1, per, per + 1, 2 * per - 1, n - 1):
        assert sg.export_g1(i, 1) == g1c(pow(S, i + 5, O.R)), i
    sg.close()
    group.set_option("setup_layout", 0)


def test_cpp_single_process_multi_gpu(native):
    """tests/cpp/test_multi.cpp: a C++ host commits n = 2^20 on all visible GPUs from one process, bit-equal to one GPU
    and to the golden commitment; with one GPU the group lists it twice."""
    import torch
    src = os.path.join(ROOT, "tests", "cpp", "test_multi.cpp")
    binp = os.path.join(ROOT, "tests", "cpp", "test_multi.bin")
    if not os.path.exists(binp) or os.path.getmtime(binp) < max(os.path.getmtime(src), os.path.getmtime(native.LIB_PATH)):
        subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), src, "-o", binp,
                        "-L", os.path.dirname(native.LIB_PATH), "-l:libkzg_b200.so",
                        "-Wl,-rpath," + os.path.dirname(native.LIB_PATH)], check=True)
    args = [binp, "20"] + (["0,0"] if torch.cuda.device_count() < 2 else [])
    r = subprocess.run(args, capture_output=True, text=True, timeout=900)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "multi-device C ABI test ok" in r.stdout
