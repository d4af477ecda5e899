"""The index arithmetic of the round-2 bucket reductions, restated over the integers mod a prime (the identities hold in
any abelian group, so the additive group of Z/p stands in for G1): what the kernels compute must equal sum_b (b+1) * B_b.

* k_reduce_l1k (csrc/msm_tail.cu): a lane walks 2^LK consecutive buckets with acc += B; run += acc, a warp of 32 lanes
  turns (S_l, W_l) into S = sum S_l and W = sum_l (2^LK * l * S_l + W_l) through ONE tree over Y_l = 2^LK * R_l + W_l,
  R_l = sum_{m >= l} S_m; levels 2 / 3 and the Horner step weigh chunk k by 32 * 2^LK * k.
* finish_level2 (csrc/msm.cu): digit d = b + 1 = hi * 2^s + lo; the points of bucket b go to the hi set (bucket hi - 1)
  and to the lo set (bucket lo - 1); result = 2^s * sum_h (h) * H_h + sum_l (l) * L_l, with the hi group of a bucket
  found as the contiguous range that starts at bucket b - lo."""
import random

P = (1 << 61) - 1


def reduce_reference(B):
    return sum((b + 1) * v for b, v in enumerate(B)) % P


def l1k(B, lk):
    """returns (S_out, W_out) per chunk of 32 * 2^lk buckets, as k_reduce_l1k does"""
    K, M = 1 << lk, len(B)
    chunks = (M + 32 * K - 1) // (32 * K)
    S_out, W_out = [], []
    for blk in range(chunks):
        S, Wl = [], []
        for lane in range(32):
            b0 = (blk * 32 + lane) * K
            acc = run = 0
            for i in range(K - 1, -1, -1):
                if b0 + i < M:
                    acc = (acc + B[b0 + i]) % P
                if i > 0:
                    run = (run + acc) % P
            S.append(acc); Wl.append(run)
        R = [sum(S[l:]) % P for l in range(32)]                 # suffix scan
        S_out.append(R[0])
        Y = [((R[l] << lk) if l >= 1 else 0) + Wl[l] for l in range(32)]
        W_out.append(sum(Y) % P)
    return S_out, W_out


def tail(B, lk):
    """levels 1..3 + Horner: S_total + sum W0 + 32 * 2^lk * (sum W1 + 32 * W2)"""
    S0, W0 = l1k(B, lk)
    S1, W1 = l1k(S0, 0)                                           # level 2: plain chunks of 32
    assert len(S1) <= 32
    S2, W2 = l1k(S1, 0)
    acc = W2[0]
    acc = (acc * 32 + sum(W1)) % P
    acc = (acc * (32 << lk) + sum(W0)) % P
    return (acc + S2[0]) % P


def test_running_sum_level_one_and_horner():
    rnd = random.Random(7)
    for M, lk in ((16384, 2), (32768, 2), (65536, 3), (65536, 1), (5000, 1), (1024, 0), (40000, 3), (131072, 3)):
        B = [rnd.randrange(P) if rnd.random() < 0.9 else 0 for _ in range(M)]
        assert tail(B, lk) == reduce_reference(B), (M, lk)


def level2(seg, pts, half, calls, c):
    """seg[0 .. calls * half] bounds of the dense points over the level-1 buckets; returns one result per call"""
    s_lo = (c - 1) // 2
    MS = 1 << ((c - 1) - s_lo)
    cnt = [0] * (2 * calls * MS)
    off_lo = {}
    for b in range(calls * half):                                 # k_l2_count
        call, d = divmod(b, half); d += 1
        hi, lo = d >> s_lo, d & ((1 << s_lo) - 1)
        if lo == 0:
            end = min(b + (1 << s_lo), (call + 1) * half)
            cnt[(call * 2) * MS + hi - 1] = seg[end] - seg[b]
            continue
        ln = seg[b + 1] - seg[b]
        if ln:
            off_lo[b] = cnt[(call * 2 + 1) * MS + lo - 1]
            cnt[(call * 2 + 1) * MS + lo - 1] += ln
    start2 = [0]
    for v in cnt:
        start2.append(start2[-1] + v)
    vals2 = [None] * start2[-1]
    for b in range(calls * half):                                 # k_l2_fill
        p0, ln = seg[b], seg[b + 1] - seg[b]
        if not ln:
            continue
        call, d = divmod(b, half); d += 1
        hi, lo = d >> s_lo, d & ((1 << s_lo) - 1)
        if hi:
            at = start2[(call * 2) * MS + hi - 1] + (p0 - seg[b - lo])
            for i in range(ln):
                assert vals2[at + i] is None
                vals2[at + i] = p0 + i
        if lo:
            at = start2[(call * 2 + 1) * MS + lo - 1] + off_lo[b]
            for i in range(ln):
                assert vals2[at + i] is None
                vals2[at + i] = p0 + i
    assert all(v is not None for v in vals2)
    res = []
    for call in range(calls):
        sets = []
        for which in range(2):
            g0 = (call * 2 + which) * MS
            sums = [sum(pts[v] for v in vals2[start2[g0 + k]:start2[g0 + k + 1]]) % P for k in range(MS)]
            sets.append(reduce_reference(sums))
        res.append(((sets[0] << s_lo) + sets[1]) % P)             # combine_level2
    return res


def test_second_level_split():
    rnd = random.Random(11)
    for c, calls in ((5, 1), (6, 3), (8, 2), (9, 4), (11, 2)):
        half = 1 << (c - 1)
        lens = [rnd.choice((0, 0, 1, 1, 2, 3)) for _ in range(calls * half)]
        lens[half - 1] = 2                                        # the top digit d = 2^(c-1): hi = MS, lo = 0
        lens[0] = 3                                               # d = 1: hi = 0, only a lo entry
        seg = [0]
        for ln in lens:
            seg.append(seg[-1] + ln)
        pts = [rnd.randrange(P) for _ in range(seg[-1])]
        want = []
        for call in range(calls):
            B = [sum(pts[seg[call * half + b]:seg[call * half + b + 1]]) % P for b in range(half)]
            want.append(reduce_reference(B))
        assert level2(seg, pts, half, calls, c) == want, (c, calls)
