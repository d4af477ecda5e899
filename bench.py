#!/usr/bin/env python3
"""bench.py — KZG commit MSM points/s at n = 2^20, 1..8 B200; EIP-4844 blob commits/s (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA engine (under torchrun for N > 1)
  python bench.py --gpus N --single-process                      # the same from ONE host process (kzg_ctx_create_multi)
  python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU loop (C port, all host threads)

A "step" is one commit-MSM of n = 2^20 synthetic scalars U(0xC3, i) (BASELINE config C3) against the
powers-of-tau setup of the reference's bench secret (00 01 .. 1f); every step's 48-byte commitment is read back and
compared with the golden value (known-secret identity, tests/golden/bench_commitment.json).
  value  : points/s, scalars resident in HBM, `lanes` x `coalesce` calls in flight (kzg_msm_g1_submit_dev / _wait; option
           msm_coalesce: consecutive calls run as ONE bucket problem), the one-call-at-a-time figure beside it
  e2e    : the same through the host-buffer C-ABI calls: pinned host scalars, 32 MiB H2D + 48 B D2H inside the timed region
  N > 1  : one process per GPU (torchrun); rank g owns bases / scalars [g n/N, (g+1) n/N) (point-range sharding); the
           192-byte partial sums are exchanged by ONE kernel over NVLink peer memory (k_exchange_combine), with an NCCL
           all-gather line beside it; strong scaling.  --single-process drives the same devices from one process.
The line also carries, each with roofline / golden checks of its own:
  blob_batch    : BASELINE config C4, 4096 blobs x 4096 coefficients, commit + proof, sharded by blob (the second half of
                  the metric): roofline, CPU baseline, golden checksums of every rank's shard
  other_configs : C5 (2^22 setup points, sharded by index range, golden items), C2 (degree 2^16), the reference's
                  Criterion degrees (1 / 100 / 500 / 1000 / 2500) beside the 1-thread CPU port
The folded-window / fixed-base tables are part of the setup object (built once per setup, reported in `config`), not of
the timed region.  Nothing here reads /root/reference; oracle/ is used only for the cpu_baseline legs."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N = 20
SEED = 0xC3
MAC_PER_FQ_MUL = 300          # 12x12 product + 12x12+12 reduction
MAC_PER_BUCKET_ADD = 3000     # XYZZ += affine: 8M + 2S = 10 Fq multiplications
CANONICAL_MAC_PER_POINT = 48000  # c = 16: 16 windows x 3000 (SURVEY 8d canonical figure)
CANONICAL_MAC_PER_BLOB = 8191 * CANONICAL_MAC_PER_POINT   # commit (4096 points) + proof (4095): 3.93e8
NOMINAL_LANES_PER_CLK_SM = 64    # 32-bit integer multiply lanes per clock and SM (SURVEY 8d's assumption)
# dram__bytes_read.sum + dram__bytes_write.sum per launch from `ncu --set full` captures (profiles/*_summary.txt)
TRAFFIC_NCU = {"k_seg_add<true>": 2.513057e9 + 0.738046720e9,          # n = 2^20, c = 17 (profiles/prof_seg_add_lean_r1_r02_summary.txt)
               "k_r1_add": 28.939577e9 + 8.170940e9}                     # 2048 rows x 4096, c = 15, row-minor mapping (profiles/prof_r1_add_r02n_summary.txt)
GOLDEN_C3_COMMITMENT = "a9a29f4f21d2c6481dc286c676d58f57b1796df6e04854503c6d5348c202851fd361a5ec4624bee2e1a27092e80f7241"
DTYPE = "u32-limb modular integer (Fq 381-bit, Fr 255-bit)"
CRITERION_DEGREES = (1, 100, 500, 1000, 2500)     # benches/polynomial_commitment.rs:26-32


# ---- clocks ---------------------------------------------------------------------------------------------------------
def _clock_sampler(stop, samples, device_index):
    """SM clock / throttle reasons DURING the timed region.  In-process NVML (nvidia_ml_py) when it loads: a sample
    costs microseconds, where spawning nvidia-smi costs ~0.4 s and yields 2-4 samples per run; nvidia-smi otherwise."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(device_index)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:                                   # NVML enumerates physical devices
            try:
                h = nv.nvmlDeviceGetHandleByIndex(int(vis.split(",")[device_index]))
            except Exception:
                pass
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not stop.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = get_reasons(h)
            try:
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
            except Exception:
                pw = 0.0
            samples.append([str(sm), str(mx), str(pw), hex(r)] +
                           ["Active" if r & bits[k] else "Not Active" for k in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")])
            stop.wait(0.01)
        return
    except Exception:
        pass
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(device_index)],
                               capture_output=True, text=True, timeout=5)
            if r.returncode == 0 and r.stdout.strip():
                samples.append([x.strip() for x in r.stdout.strip().split(",")])
        except Exception:
            pass
        stop.wait(0.1)


def _clocks_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = sorted(float(s[0]) for s in samples if s[0].replace(".", "").isdigit())
    reasons = set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for s in samples:
        for name, v in zip(names, s[4:8]):
            if v.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(samples[0][1]) if samples[0][1].replace(".", "").isdigit() else None,
            "reasons": sorted(reasons), "samples": len(samples)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def int_peaks(ctx, clocks=None):
    """the two denominators of every integer roofline in the line: the multiplier pipe's issue rate measured in this run
    (kzg_probe_imad_peak: independent IMAD.WIDE.U32) and the nominal 64 lanes x SMs x max SM clock"""
    peaks, _ = measured_peaks()
    mhz = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    sms = 148
    try:
        import torch
        sms = torch.cuda.get_device_properties(ctx.device).multi_processor_count
    except Exception:
        pass
    return {"measured": ctx.probe_imad_peak(), "nominal": NOMINAL_LANES_PER_CLK_SM * sms * mhz * 1e6,
            "fq_mul_per_s": ctx.probe_fq_mul(), "sms": sms, "sm_max_mhz": mhz}


# ---- the reference's CPU path (oracle/cpu_ref.c): cpu_baseline legs and --impl reference -------------------------------
def _cpu_ref():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_ref
    return cpu_ref


def cpu_baseline(points: int, threads: int):
    """The reference's commit loop (C port, oracle/cpu_ref.c) on the host cores: bounded sample."""
    from kzg_poly_commit_exploration_b200 import synthetic
    st = _cpu_ref().RefSetup(synthetic.BENCH_SECRET_BE, points, threads)
    coeffs = synthetic.scalar_stream(SEED, points).tobytes()
    t0 = time.perf_counter()
    rc, _ = st.commit(coeffs, threads)
    dt = time.perf_counter() - t0
    st.close()
    assert rc == 0
    return points / dt, dt


def cpu_baseline_blobs(threads: int, n: int = 4096):
    """commit + evaluate + proof of whole blobs with the reference's loops (src/polynomial.rs:112-123, 200-215, 260-269),
    one blob per host thread (the reference itself is single-threaded): blobs/s over a bounded sample"""
    from concurrent.futures import ThreadPoolExecutor
    from kzg_poly_commit_exploration_b200 import synthetic
    ref = _cpu_ref()
    st = ref.RefSetup(synthetic.BENCH_SECRET_BE, n, threads)
    blobs = 2 * threads
    co = synthetic.scalar_stream(0xC4, blobs * n).tobytes()
    zs = synthetic.scalar_stream(0xC4F, blobs).tobytes()

    def one(b):
        row, z = co[b * n * 32: (b + 1) * n * 32], zs[32 * b: 32 * b + 32]
        rc, c = st.commit(row, 1)
        y = ref.evaluate(row, z)
        rc2, p = st.open(row, z, y, 1)
        assert rc == 0 and rc2 == 0
        return c, y, p

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        outs = list(ex.map(one, range(blobs)))
    dt = time.perf_counter() - t0
    st.close()
    return blobs / dt, dt, blobs, outs


def run_reference(args):
    """--impl reference: the reference's own CPU path for this metric.  The Rust crate + blst cannot be
    built in this image (no cargo; blst un-vendored), so this is the C port of its loop
    (src/polynomial.rs:207-212) on all host threads, each step a bounded prefix of the 2^20 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 8192 * threads
    from kzg_poly_commit_exploration_b200 import synthetic
    st = _cpu_ref().RefSetup(synthetic.BENCH_SECRET_BE, sample, threads)
    coeffs = synthetic.scalar_stream(SEED, sample).tobytes()
    for _ in range(args.warmup):
        st.commit(coeffs, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc, _ = st.commit(coeffs, threads)
        assert rc == 0
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": "kzg_commit_msm_points_per_s", "value": value, "unit": "points/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE,
        "data": "synthetic", "config": {"workload": f"commit MSM n=2^{LOG_N} (C3), reference naive loop on a {sample}-point prefix per step"},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port",
                         "sample": f"{sample}-point prefix of the 2^{LOG_N} workload per step, loop split over {threads} threads "
                                   "(parallelised restatement; the reference itself is single-threaded)"},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---- host-side group sums of compressed points (kzg_g1_add: host code of the library, not the kernels under test) --------
def g1_sum(pkg, points48: bytes) -> bytes:
    L = pkg.lib()
    acc = (C.c_uint8 * 48)()
    acc[0] = 0xC0                                                   # infinity
    o = (C.c_uint8 * 48)()
    for i in range(0, len(points48), 48):
        assert L.kzg_g1_add(acc, (C.c_uint8 * 48).from_buffer_copy(points48[i: i + 48]), o) == 0
        acc, o = o, acc
    return bytes(acc)


def fr_sum_le(values32: bytes) -> bytes:
    R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    t = sum(int.from_bytes(values32[i: i + 32], "little") for i in range(0, len(values32), 32)) % R
    return t.to_bytes(32, "little")


def load_golden():
    p = os.path.join(ROOT, "tests", "golden", "bench_blob_setup.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        return json.load(f)


class Ranks:
    """torch.distributed plumbing of the one-process-per-GPU mode (NCCL for device tensors, a gloo group for host
    barriers that do not spin on the GPU); everything degenerates to no-ops for a single process."""

    def __init__(self, single_process: bool):
        import torch
        self.torch = torch
        self.world = 1 if single_process else int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = 0 if single_process else int(os.environ.get("RANK", "0"))
        self.local = 0 if single_process else int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.host_group = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{self.local}"))
            self.dist = dist
            self.host_group = dist.new_group(backend="gloo")
        torch.cuda.set_device(self.local)
        self.dev = torch.device(f"cuda:{self.local}")

    def barrier(self):
        if self.dist:
            self.dist.barrier()

    def host_barrier(self):
        if self.dist:
            self.dist.barrier(group=self.host_group)

    def max_(self, *vals):
        if not self.dist:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def all_true(self, flag: bool) -> bool:
        if not self.dist:
            return bool(flag)
        t = self.torch.tensor([1.0 if flag else 0.0], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item() > 0.5)

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


# ---- the headline: one MSM of n = 2^20 -----------------------------------------------------------------------------------
def msm_headline(args, R, pkg, synthetic, ctx, devices):
    """devices = list of device ids the context drives: [local] in the one-process-per-GPU mode, all of them for a
    single-process group context (then R.world == 1 and the engine shards internally)."""
    torch = R.torch
    world, rank, dev = R.world, R.rank, R.dev
    group = len(devices) > 1
    n_gpus = len(devices) if group else world
    n_total = 1 << LOG_N
    per = n_total // world                 # scalars this PROCESS holds
    first = rank * per
    if group:
        ctx.set_option("setup_layout", 1)
    setup = ctx.setup_generate(synthetic.BENCH_SECRET_BE, per, first=first)
    t0 = time.perf_counter()
    setup.prepare_msm(args.window)
    ctx.sync()
    table_s = time.perf_counter() - t0
    plan = setup.msm_plan(per)             # of one device's share
    per_dev = n_total // n_gpus

    host = torch.from_numpy(synthetic.scalar_stream(SEED, per, start=first)).pin_memory()
    scal_dev = host.to(dev)
    partial = torch.zeros(192, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(192 * world, dtype=torch.uint8, device=dev)
    ext_stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    torch.cuda.synchronize()

    use_p2p = world > 1
    if use_p2p:
        from kzg_poly_commit_exploration_b200 import parallel
        parallel.connect_peers(ctx, world, rank)     # torch.distributed carries the 64-byte IPC handles, once; ends with a barrier

    def step_dev():
        if world == 1:
            return ctx.msm_g1_dev(setup, scal_dev.data_ptr(), per)
        return ctx.msm_g1_sharded_dev(setup, scal_dev.data_ptr(), per)

    def step_nccl():
        # baseline exchange: partial MSM, NCCL all-gather of the 192-byte partials and the fold enqueued in order on the
        # engine's stream (torch sees it as an ExternalStream); no host synchronisation until the 48-byte result
        ctx.msm_g1_partial_dev(setup, scal_dev.data_ptr(), per, partial.data_ptr())
        with torch.cuda.stream(ext_stream):
            R.dist.all_gather_into_tensor(gathered, partial)
        return ctx.g1_combine_dev(gathered.data_ptr(), world)

    def step_e2e():
        if world == 1:
            return ctx.msm_g1(setup, host.data_ptr(), per)
        t = ctx.msm_g1_sharded_submit(setup, host.data_ptr(), per)
        return ctx.msm_g1_wait(t)

    r_ref = [None]
    depth = args.lanes * args.coalesce       # tickets out at once: `lanes` groups of `coalesce` fused calls

    def calls_in_flight(k, submit):
        from collections import deque
        q, out = deque(), None
        for _ in range(k):
            if len(q) == depth:
                out = ctx.msm_g1_wait(q.popleft())
                assert out == r_ref[0], "a pipelined step returned a different commitment"
            q.append(submit())
        while q:
            out = ctx.msm_g1_wait(q.popleft())
            assert out == r_ref[0], "a pipelined step returned a different commitment"
        return out

    def steps_dev_pipelined(k):
        if world == 1:
            return calls_in_flight(k, lambda: ctx.msm_g1_submit_dev(setup, scal_dev.data_ptr(), per))
        return calls_in_flight(k, lambda: ctx.msm_g1_sharded_submit_dev(setup, scal_dev.data_ptr(), per))

    def steps_e2e_pipelined(k):
        if world == 1:
            return calls_in_flight(k, lambda: ctx.msm_g1_submit(setup, host.data_ptr(), per))
        return calls_in_flight(k, lambda: ctx.msm_g1_sharded_submit(setup, host.data_ptr(), per))

    def sync_all():
        R.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def timed(fn_k, steps):
        """K steps between barriers; device time = CUDA events on the engine stream(s) (every lane / member joins before
        the stop event), max over ranks; the host clock of the same region beside it"""
        sync_all()
        ctx.timer_start()
        t0 = time.perf_counter()
        out = fn_k(steps)
        ms = ctx.timer_stop()
        sync_all()
        wall = (time.perf_counter() - t0) * 1e3
        ms, wall = R.max_(ms, wall)
        return ms, wall, out

    def loop(fn):
        def run(k):
            out = None
            for _ in range(k):
                out = fn()
            return out
        return run

    for _ in range(args.warmup):
        r_dev = step_dev()
    r_ref[0] = r_dev
    steps_dev_pipelined(max(2 * depth, args.warmup))

    stop, samples = threading.Event(), []
    th = threading.Thread(target=_clock_sampler, args=(stop, samples, R.local), daemon=True)
    if rank == 0:
        th.start()
    launches0 = ctx.launches
    ms_single, _, r_dev = timed(loop(step_dev), args.steps)          # one call at a time
    launches = ctx.launches - launches0
    ms_total, wall_total, r_pipe = timed(steps_dev_pipelined, args.steps)     # `lanes` calls in flight: the headline
    assert r_pipe == r_dev, "pipelined and one-at-a-time results differ"
    acc_ms, dom_ms, whole_ms = [], [], []
    for _ in range(min(args.steps, 5)):   # per-launch durations: CUDA events on the stream (lane) the call ran on
        step_dev()
        ctx.sync()
        acc_ms.append(ctx.last_kernel_ms(0))   # accumulation phase (halving rounds + bucket accumulate)
        dom_ms.append(ctx.last_kernel_ms(2))   # its dominant kernel
        whole_ms.append(ctx.last_kernel_ms(1))
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    ms_e2e_sync, _, r_e2e = timed(loop(step_e2e), args.steps)
    steps_e2e_pipelined(2 * depth)
    ms_e2e, _, r_pipe = timed(steps_e2e_pipelined, args.steps)
    assert r_pipe == r_e2e, "pipelined and synchronous host-buffer results differ"
    nccl = None
    if world > 1:                        # the NCCL all-gather exchange, one line per N for the record
        for _ in range(3):
            r_n = step_nccl()
        ms_nccl, _, r_n = timed(loop(step_nccl), args.steps)
        assert r_n == r_dev, "NCCL exchange and peer exchange disagree"
        nccl = {"value": n_total / (ms_nccl / args.steps * 1e-3), "unit": "points/s", "ms_per_step": ms_nccl / args.steps,
                "what": "partial MSM + ncclAllGather of the 192-byte partials + fold kernel, one call at a time "
                        "(compare with one_call_at_a_time)"}
    stop.set()
    if rank == 0:
        th.join(timeout=2)
    assert r_dev == r_e2e, "device-resident and host-buffer results differ"
    golden_ok = r_dev.hex() == GOLDEN_C3_COMMITMENT
    if LOG_N == 20:
        assert golden_ok, "the commitment differs from the known-secret golden value"
    clocks = _clocks_summary(samples)
    pk = int_peaks(ctx, clocks)
    setup.close()
    del host, scal_dev, partial, gathered
    torch.cuda.synchronize()
    if rank != 0:
        return None, pk

    med = lambda v: sorted(v)[len(v) // 2]
    ms_step = ms_total / args.steps
    single_step = ms_single / args.steps
    value = n_total / (ms_step * 1e-3)
    peaks, peaks_src = measured_peaks()
    acc, dom, whole = med(acc_ms), med(dom_ms), med(whole_ms)
    adds = plan["W"] * per_dev                      # bucket additions of the whole phase, per device
    rounds = plan["rounds"]
    tree = rounds > 0
    dom_adds = adds // 2 if tree else adds          # round 1 pairs up all entries: half of the additions
    dom_name = "k_seg_add_lean<true, 4> (round-1 pair additions of the affine halving, msm_tree.cu)" if tree else "k_bucket_acc"
    phase_achieved = adds * MAC_PER_BUCKET_ADD / (acc * 1e-3)
    dom_impl = dom_adds * (1500 if tree else 3000) / (dom * 1e-3)
    hbm_bytes = dom_adds * (2 * 96 + 2 * 4 + 48 + 96) if tree else adds * 100  # per pair: two table points, two indices, prefix in, point out
    canonical = per_dev * CANONICAL_MAC_PER_POINT / (ms_step * 1e-3)
    mode = ("one host process drives all devices (kzg_ctx_create_multi): worker thread per device, peer-memory exchange"
            if group else "one process per GPU (torchrun)" if world > 1 else "one process, one GPU")
    line = {
        "metric": "kzg_commit_msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "wall_ms_per_step": wall_total / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE, "data": "synthetic",
        "config": {"workload": f"single G1 commit-MSM n=2^{LOG_N} (BASELINE config C3), uniform scalars U(0xC3,i), bench-secret setup",
                   "mode": mode, "window_bits": plan["c"], "windows": plan["W"], "halving_rounds": rounds,
                   "halving_rounds_note": "of a call on its own; fused calls count their entries together (one round more per doubling)",
                   "sharding": f"point range, {per_dev} points/GPU",
                   "exchange": ("none" if n_gpus == 1 else "peer stores over NVLink + fold + compress in one kernel (k_exchange_combine)"),
                   "l2": f"inputs larger than L2 (table {plan['W'] * per_dev * plan['table_pitch'] / 2**20:.0f} MiB + scalars {per_dev * 32 / 2**20:.0f} MiB per GPU)",
                   "precomputed_table_s": round(table_s, 3), "calls_in_flight": depth, "lanes": args.lanes,
                   "fused_calls": args.coalesce,
                   "step": f"one complete MSM (48-byte result read back and compared with the golden commitment); steps are submitted {depth} deep "
                           f"(kzg_msm_g1_submit_dev / _wait): {args.lanes} lanes, each running {args.coalesce} consecutive calls as ONE bucket problem "
                           "(option msm_coalesce: bucket = call x 2^(c-1) + digit, same table), so the bulk kernels of one group overlap the "
                           "latency-bound phases of the others"},
        "one_call_at_a_time": {"value": n_total / (single_step * 1e-3), "unit": "points/s", "ms_per_step": single_step},
        "e2e": {"value": n_total / (ms_e2e / args.steps * 1e-3), "unit": "points/s", "h2d_bytes_per_step": n_total * 32,
                "d2h_bytes_per_step": 48 * world, "ms_per_step": ms_e2e / args.steps,
                "call": (f"kzg_msm_g1_submit / kzg_msm_g1_wait, {depth} calls in flight ({args.lanes} lanes x {args.coalesce} fused), pinned host scalars"
                         + ("" if world == 1 else " (kzg_msm_g1_sharded_submit: each rank copies its shard)")),
                "one_call_at_a_time": {"value": n_total / (ms_e2e_sync / args.steps * 1e-3), "unit": "points/s",
                                       "ms_per_step": ms_e2e_sync / args.steps}},
        "gpu_launches": int(launches),
        "commitment_c48": r_dev.hex(), "commitment_matches_golden": golden_ok,
        "clocks": clocks,
        "roofline": {
            "bound": "int32-multiply",
            "bound_note": "integer-multiply pipe (IMAD.WIDE), neither hbm nor tensor: the MSM is integer-compute-bound "
                          "(north_star); the HBM side of the dominant kernel is reported under dominant_kernel.hbm",
            "kernel": (f"bucket accumulation = k_seg_den + batch inversion + k_seg_add ({rounds} affine halving rounds, "
                       "msm_tree.cu) + k_bucket_acc" if tree else "k_bucket_acc"),
            "achieved": phase_achieved / 1e12, "peak": pk["measured"] / 1e12, "unit": "T wide-MAC/s",
            "frac": phase_achieved / pk["measured"],
            "peak_nominal": pk["nominal"] / 1e12, "frac_nominal": phase_achieved / pk["nominal"],
            "work": f"algorithmic: {adds} bucket additions x 3000 wide-MAC (XYZZ += affine = 10 Fq products, SURVEY 8d) per device",
            "peak_source": "kzg_probe_imad_peak in this run: issue rate of independent IMAD.WIDE.U32 (the round-1 probe let ptxas "
                           f"fold three of four products and under-reported it); nominal = {NOMINAL_LANES_PER_CLK_SM} lanes x "
                           f"{pk['sms']} SMs x {pk['sm_max_mhz']:.0f} MHz",
            "kernel_ms": acc, "call_ms": whole, "kernel_share_of_step": acc / single_step,
            "share_note": "shares are of the one-call-at-a-time step (kernel durations are measured without overlap, on the "
                          "lane the probed call ran on)",
            "dominant_kernel": {
                "name": dom_name, "ms": dom, "share_of_step": dom / single_step,
                "as_implemented": ("5 Fq products = 1500 wide-MAC per pair addition (denominators and the shared "
                                   "inversion run in k_seg_den / k_inv_*)" if tree else "10 Fq products = 3000 wide-MAC per addition"),
                "achieved": dom_impl / 1e12, "frac": dom_impl / pk["measured"], "frac_nominal": dom_impl / pk["nominal"],
                "hbm": {"achieved": hbm_bytes / (dom * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "frac": hbm_bytes / (dom * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0), "peak_source": peaks_src,
                        "algorithmic_bytes": hbm_bytes}},
            "fq_mul_probe": {"per_s": pk["fq_mul_per_s"], "wide_mac_per_s": pk["fq_mul_per_s"] * MAC_PER_FQ_MUL,
                             "frac_of_measured_peak": pk["fq_mul_per_s"] * MAC_PER_FQ_MUL / pk["measured"],
                             "what": "the shipped carry-chained Montgomery product on all SMs: the practical ceiling of every kernel above"},
            "int_pipe_busy_ncu": {"k_seg_add_lean<true, 4>": 0.757, "k_seg_add_lean<false, 4>": 0.821,
                                  "metric": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed (the pipe that "
                                            "executes IMAD / IMAD.WIDE), ncu --set full captures summarised in "
                                            "profiles/prof_seg_add_lean_r1_r02_summary.txt and prof_seg_add_lean_dense_r02_summary.txt"},
            "canonical_frac": canonical / pk["measured"], "canonical_frac_nominal": canonical / pk["nominal"],
            "canonical_note": "48 000 wide-MAC per point (SURVEY 8d) x points per device / pipelined step",
            "traffic": TRAFFIC_NCU["k_seg_add<true>"] if tree and per_dev == (1 << 20) and plan["c"] == 17 else None,
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one k_seg_add_lean<true, 4> launch at n = 2^20, c = 17 per device "
                            "(the dominant kernel; 1.20 x its algorithmic bytes: the pad of the 128-byte table pitch), "
                            "profiles/prof_seg_add_lean_r1_r02_summary.txt"},
    }
    if nccl:
        line["exchange_nccl"] = nccl
    return line, pk


# ---- BASELINE configs C5 / C2 / the reference's Criterion degrees ------------------------------------------------------------
def other_configs(args, R, pkg, synthetic, ctx, devices):
    torch = R.torch
    world, rank = R.world, R.rank
    group = len(devices) > 1
    n_gpus = len(devices) if group else world
    out = {}
    n5 = 1 << 22
    per = n5 // world
    first = rank * per
    if group:
        ctx.set_option("setup_layout", 1)
    st = ctx.setup_generate(synthetic.BENCH_SECRET_BE, 1 << 12)   # builds the generator table (once per ctx)
    st.close()
    dts = []
    for rep in range(3):                      # the first pass also pays for the scratch allocations (~1.5 GB)
        if rep:
            st.close()
        R.barrier()
        ctx.sync()
        t0 = time.perf_counter()
        st = ctx.setup_generate(synthetic.BENCH_SECRET_BE, per, first=first)
        ctx.sync()
        dts.append(R.max_(time.perf_counter() - t0)[0])
    dt = sorted(dts)[1]                       # median of 3
    # every rank checks the golden items that fall into the range it generated itself (tests/golden/bench_blob_setup.json)
    golden = load_golden()
    ok, checked = True, 0
    if golden:
        for k, v in golden["setup_2^22"]["items"].items():
            i = int(k)
            if first <= i < first + per:
                ok = ok and st.export_g1(i - first, 1).hex() == v
                checked += 1
    ok = R.all_true(ok)
    n_g2 = min(per // (len(devices) if group else 1), 1 << 18) * (len(devices) if group else 1)
    st.generate_g2(synthetic.BENCH_SECRET_BE, n_g2)         # builds the G2 generator table (once per ctx) and the scratch
    ctx.sync()
    g2_dts = []
    for _ in range(3):
        t0 = time.perf_counter()
        st.generate_g2(synthetic.BENCH_SECRET_BE, n_g2)
        ctx.sync()
        g2_dts.append(time.perf_counter() - t0)
    dt_g2 = sorted(g2_dts)[1]
    st.close()
    per_dev = n5 // n_gpus
    out["setup_2^22_g1"] = {"metric": "setup_points_per_s", "value": n5 / dt, "unit": "G1 points/s", "ms": dt * 1e3,
                            "sharding": f"index range, {per_dev} points/GPU", "wide_mac_per_point": 48000,
                            "matches_golden": bool(ok) if golden else None, "golden_items_checked_on_rank0": checked,
                            "path": "affine pairwise trees over the generator table (fixed_base_points_tree), 90 Fq products per point"
                                    if per_dev >= (1 << 18) else "XYZZ table walk + batched to-affine",
                            "g2_points_per_s_per_process": n_g2 / dt_g2, "g2_points_timed": n_g2}
    if world == 1 and not group:
        n2 = (1 << 16) + 1
        st = ctx.setup_generate(synthetic.BENCH_SECRET_BE, n2)
        st.prepare_msm(0)
        co = torch.from_numpy(synthetic.scalar_stream(0xC2, n2)).pin_memory()
        z = synthetic.scalar_stream(0xC2F, 1).tobytes()
        ctx.commit(st, co.data_ptr(), n2); ctx.open(st, co.data_ptr(), z, None, n2)
        t0 = time.perf_counter(); ctx.commit(st, co.data_ptr(), n2); t_commit = time.perf_counter() - t0
        t0 = time.perf_counter(); ctx.open(st, co.data_ptr(), z, None, n2); t_open = time.perf_counter() - t0
        t0 = time.perf_counter(); ctx.evaluate(co.data_ptr(), z, n2); t_eval = time.perf_counter() - t0
        st.close()
        out["degree_2^16"] = {"commit_ms": t_commit * 1e3, "evaluate_ms": t_eval * 1e3, "open_ms": t_open * 1e3,
                              "note": "host buffers through the C ABI (H2D + D2H inside), one polynomial, 1 GPU"}
        out["criterion_degrees"] = criterion_sweep(pkg, synthetic, ctx)
    return out


def criterion_sweep(pkg, synthetic, ctx):
    """The reference's own bench matrix (benches/polynomial_commitment.rs:10-47, benches/evaluation_proof.rs): c_i = 5^i + 10,
    z = 5^d + 20, degrees 1 / 100 / 500 / 1000 / 2500 — kzg_commit / kzg_open through the C ABI (host buffers, one call at
    a time, median of 7) beside the 1-thread C port of the reference loop (the reference is single-threaded)."""
    R_ = synthetic.R
    kat_path = os.path.join(ROOT, "tests", "golden", "kat_bench_inputs.json")
    kat = json.load(open(kat_path))["degrees"] if os.path.exists(kat_path) else {}
    ref = _cpu_ref()
    dmax = max(CRITERION_DEGREES)
    st = ctx.setup_generate(synthetic.BENCH_SECRET_BE, dmax + 1)
    cst = ref.RefSetup(synthetic.BENCH_SECRET_BE, dmax + 1, os.cpu_count() or 1)
    rows = {}
    for d in CRITERION_DEGREES:
        co = b"".join(((pow(5, i, R_) + 10) % R_).to_bytes(32, "little") for i in range(d + 1))
        z = ((pow(5, d, R_) + 20) % R_).to_bytes(32, "little")
        c = ctx.commit(st, co)
        y, p = ctx.open(st, co, z)
        tc, to = [], []
        for _ in range(7):
            t0 = time.perf_counter(); ctx.commit(st, co); tc.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); ctx.open(st, co, z); to.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); rc, c_cpu = cst.commit(co, 1); t_cpu_c = time.perf_counter() - t0
        t0 = time.perf_counter(); rc2, p_cpu = cst.open(co, z, y, 1); t_cpu_o = time.perf_counter() - t0
        k = kat.get(str(d))
        rows[str(d)] = {"gpu_commit_ms": sorted(tc)[3] * 1e3, "gpu_open_ms": sorted(to)[3] * 1e3,
                        "cpu_port_commit_ms": t_cpu_c * 1e3, "cpu_port_open_ms": t_cpu_o * 1e3,
                        "equal_to_cpu_port": rc == 0 and rc2 == 0 and c == c_cpu and p == p_cpu,
                        "matches_kat": (c.hex() == k["commit"] and p.hex() == k["proof"] and y.hex() == k["y_le"]) if k else None}
    st.close(); cst.close()
    return {"rows": rows, "cpu": "oracle/cpu_ref.c, 1 thread (kind: port)",
            "note": "latency of ONE small polynomial per call: the GPU path is launch- and latency-bound here (the reference's "
                    "Criterion sizes), throughput figures are the headline and blob_batch"}


# ---- BASELINE config C4: the blob batch -------------------------------------------------------------------------------------
def blob_batch(args, R, pkg, synthetic, ctx, devices, pk):
    """4096 blobs x 4096 coefficients, commit + evaluation proof per blob, sharded by BLOB (no data-path collective);
    blob (commit+proof)/s, its roofline (canonical 3.93e8 wide-MAC per blob; the dominant kernel k_r1_add timed alone with
    CUDA events), the CPU port beside it, and golden checksums of what every rank computed."""
    torch = R.torch
    world, rank, dev = R.world, R.rank, R.dev
    group = len(devices) > 1
    n_gpus = len(devices) if group else world
    blobs_total, n, steps = args.blobs, 4096, 2
    first, count = rank * (blobs_total // world), blobs_total // world
    if group:
        ctx.set_option("setup_layout", 2)
    setup = ctx.setup_generate(synthetic.BENCH_SECRET_BE, n)
    t0 = time.perf_counter()
    setup.prepare_fixed(n, args.fixed_window)      # 0 = the widest window whose table fits 66 % of the free memory
    ctx.sync()
    table_s = time.perf_counter() - t0
    fplan = setup.fixed_plan()
    host = torch.from_numpy(synthetic.scalar_stream(0xC4, count * n, start=first * n)).pin_memory()
    zs_host = torch.from_numpy(synthetic.scalar_stream(0xC4F, count, start=first)).pin_memory()
    coeffs = host.to(dev)
    zs = zs_host.to(dev)
    commits = torch.zeros(count * 48, dtype=torch.uint8, device=dev)
    proofs = torch.zeros(count * 48, dtype=torch.uint8, device=dev)
    ys = torch.zeros(count * 32, dtype=torch.uint8, device=dev)
    out_c = torch.zeros(count * 48, dtype=torch.uint8).pin_memory()
    out_p = torch.zeros(count * 48, dtype=torch.uint8).pin_memory()
    out_y = torch.zeros(count * 32, dtype=torch.uint8).pin_memory()
    torch.cuda.synchronize()

    def step_dev():
        # kzg_commit_open_batch_dev: the commitments on one lane of the context, the proofs on another
        ctx.commit_open_batch_dev(setup, coeffs.data_ptr(), count, n, zs.data_ptr(), commits.data_ptr(), ys.data_ptr(), proofs.data_ptr())
        ctx.sync()

    def step_e2e():
        # kzg_commit_open_batch: one host->device pass over the coefficients, chunk k+1 staged beside chunk k's kernels
        ctx.commit_open_batch(setup, host.data_ptr(), count, n, zs_host.data_ptr(), commits_out=out_c.data_ptr(),
                              y_out=out_y.data_ptr(), proofs_out=out_p.data_ptr())

    def timed(fn, k, device_timer=True):
        R.barrier()
        torch.cuda.synchronize(); ctx.sync()
        ctx.timer_start()
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        ms_dev = ctx.timer_stop()
        ctx.sync(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms_dev, wall = R.max_(ms_dev, wall)
        return ms_dev, wall

    step_dev()
    ms_dev, wall_dev = timed(step_dev, steps)
    ms_dev /= steps; wall_dev /= steps
    step_e2e()
    _, ms_e2e = timed(step_e2e, steps)             # host call: the host clock around it (copies + sync inside)
    ms_e2e /= steps
    c_dev, p_dev, y_dev = (t.cpu().numpy().tobytes() for t in (commits, proofs, ys))
    same = c_dev == out_c.numpy().tobytes() and p_dev == out_p.numpy().tobytes() and y_dev == out_y.numpy().tobytes()
    # golden checksums: the group sum of this rank's commitments / proofs and the sum of its evaluations against the
    # known-secret values of the eighths of the batch it owns (tests/golden/bench_blob_setup.json)
    golden = load_golden()
    gold_ok = None
    if golden and blobs_total == 4096 and world in (1, 2, 4, 8):
        mine = [e for e in golden["eighths"] if first <= e["first_blob"] < first + count]
        want_c = g1_sum(pkg, b"".join(bytes.fromhex(e["sum_commit"]) for e in mine))
        want_p = g1_sum(pkg, b"".join(bytes.fromhex(e["sum_proof"]) for e in mine))
        want_y = fr_sum_le(b"".join(bytes.fromhex(e["sum_y_le"]) for e in mine))
        gold_ok = g1_sum(pkg, c_dev) == want_c and g1_sum(pkg, p_dev) == want_p and fr_sum_le(y_dev) == want_y
        gold_ok = R.all_true(gold_ok)
    # the dominant kernel on its own: one 2048-row commit batch on one lane, nothing beside it
    dom = None
    if not group:
        rows = min(count, 2048)
        ctx.commit_batch_dev(setup, coeffs.data_ptr(), rows, n, commits.data_ptr()); ctx.sync()
        k_ms, tree_ms, call_ms = [], [], []
        for _ in range(3):
            ctx.commit_batch_dev(setup, coeffs.data_ptr(), rows, n, commits.data_ptr()); ctx.sync()
            k_ms.append(ctx.last_kernel_ms(2)); tree_ms.append(ctx.last_kernel_ms(0)); call_ms.append(ctx.last_kernel_ms(1))
        med = lambda v: sorted(v)[len(v) // 2]
        W = fplan["W"]
        pairs = rows * n * (W // 2)                         # pair additions of round 1 (W lookups: W/2 pairs + a passenger if W is odd)
        adds_all = rows * n * (W - 1)                       # additions of the whole row sums
        dom = {"name": "k_r1_add (round-1 pair additions over the fixed-base table, batch_tree.cu)", "ms": med(k_ms),
               "rows": rows, "share_of_commit_batch": med(k_ms) / med(call_ms),
               "algorithmic": f"{pairs} pair additions x 3000 wide-MAC", "achieved": pairs * MAC_PER_BUCKET_ADD / (med(k_ms) * 1e-3) / 1e12,
               "as_implemented": "5 Fq products = 1500 wide-MAC per pair addition (denominators in k_r1_den, shared inversion in k_inv_*)",
               "achieved_as_implemented": pairs * 1500 / (med(k_ms) * 1e-3) / 1e12,
               "unit": "T wide-MAC/s",
               "frac": pairs * MAC_PER_BUCKET_ADD / (med(k_ms) * 1e-3) / pk["measured"],
               "frac_nominal": pairs * MAC_PER_BUCKET_ADD / (med(k_ms) * 1e-3) / pk["nominal"],
               "hbm": {"achieved": pairs * (2 * 96 + 48 + 96 + 4) / (med(k_ms) * 1e-3) / 1e9, "unit": "GB/s",
                       "algorithmic_bytes": pairs * (2 * 96 + 48 + 96 + 4), "peak": measured_peaks()[0].get("hbm_gbs"),
                       "note": "two random 96-byte table points + 48-byte prefix in, 96-byte point out, scalar bytes amortised"},
               "tree_phase": {"ms": med(tree_ms), "achieved": adds_all * MAC_PER_BUCKET_ADD / (med(tree_ms) * 1e-3) / 1e12,
                              "frac": adds_all * MAC_PER_BUCKET_ADD / (med(tree_ms) * 1e-3) / pk["measured"],
                              "frac_nominal": adds_all * MAC_PER_BUCKET_ADD / (med(tree_ms) * 1e-3) / pk["nominal"],
                              "what": f"all rounds of the {rows}-row commit batch: {adds_all} additions x 3000 wide-MAC"},
               "traffic": TRAFFIC_NCU["k_r1_add"] if (rows == 2048 and fplan["c"] == 15) else None,
               "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one k_r1_add launch (2048 rows, c = 15): "
                               "profiles/prof_r1_add_r02n_summary.txt; 1.6 x the algorithmic bytes (96-byte table pitch)"}
    # the 13-bit table beside the default (memory / throughput trade on record), 1 GPU only
    alt = None
    if world == 1 and not group and args.alt_window and args.alt_window != fplan["c"]:
        setup.prepare_fixed(n, args.alt_window); ctx.sync()
        step_dev()
        ms_alt, _ = timed(step_dev, 1)
        ap = setup.fixed_plan()
        alt = {"fixed_base_window_bits": ap["c"], "table_gib": round(ap["table_bytes"] / 2**30, 2),
               "value": blobs_total / (ms_alt * 1e-3), "unit": "blobs/s"}
    setup.close()
    del host, zs_host, coeffs, zs, commits, proofs, ys
    torch.cuda.synchronize()
    cpu = None
    if world == 1 and not group and rank == 0:
        threads = os.cpu_count() or 1
        v, dt, sample, outs = cpu_baseline_blobs(threads, n)
        # the sample is the head of the batch: the CPU port and the GPU must agree on it, byte for byte
        agree = all(outs[b][0] == c_dev[48 * b: 48 * b + 48] and outs[b][1] == y_dev[32 * b: 32 * b + 32]
                    and outs[b][2] == p_dev[48 * b: 48 * b + 48] for b in range(min(sample, count)))
        cpu = {"value": v, "unit": "blobs/s", "cores": threads, "kind": "port", "equal_to_gpu_on_sample": bool(agree),
               "sample": f"the first {sample} blobs of the batch, commit + evaluate + proof with the reference's loops "
                         f"(src/polynomial.rs:112-123, 200-215, 260-269), one blob per thread on {threads} threads, {dt:.2f} s"}
    if rank != 0:
        return None
    value = blobs_total / (ms_dev * 1e-3)
    per_dev_rate = value / n_gpus
    res = {"metric": "blob_commit_plus_proof_per_s", "value": value, "unit": "blobs/s", "ms_per_batch": ms_dev,
           "wall_ms_per_batch": wall_dev, "n_gpus": n_gpus,
           "e2e": {"value": blobs_total / (ms_e2e * 1e-3), "unit": "blobs/s", "ms_per_batch": ms_e2e,
                   "h2d_bytes_per_batch": blobs_total * n * 32 + blobs_total * 32, "call": "kzg_commit_open_batch",
                   "d2h_bytes_per_batch": blobs_total * 128},
           "config": {"workload": f"{blobs_total} blobs x {n} coefficients (BASELINE config C4), commit + proof, sharded by blob",
                      "fixed_base_window_bits": fplan["c"], "lookups_per_coefficient": fplan["W"],
                      "window_choice": "given" if args.fixed_window else "auto: widest table within 66 % of the free device memory",
                      "table_gib": round(fplan["table_bytes"] / 2**30, 2), "table_build_s": round(table_s, 2)},
           "dev_equals_e2e": bool(same), "checksums_match_golden": gold_ok,
           "roofline": {"bound": "int32-multiply", "unit": "T wide-MAC/s",
                        "achieved": per_dev_rate * CANONICAL_MAC_PER_BLOB / 1e12, "peak": pk["measured"] / 1e12,
                        "frac": per_dev_rate * CANONICAL_MAC_PER_BLOB / pk["measured"],
                        "peak_nominal": pk["nominal"] / 1e12, "frac_nominal": per_dev_rate * CANONICAL_MAC_PER_BLOB / pk["nominal"],
                        "work": "canonical 8191 points x 48 000 wide-MAC = 3.93e8 per blob (SURVEY 8d), per device",
                        "as_implemented": f"{fplan['W']} table lookups per coefficient, {fplan['W'] - 1} affine pair additions of ~6.4 Fq products",
                        "dominant_kernel": dom},
           "scaling": f"strong ({blobs_total} blobs in total, {blobs_total // n_gpus} per GPU, no data-path collective)" if n_gpus > 1 else "n/a"}
    if alt:
        res["smaller_table"] = alt
    if cpu:
        res["cpu_baseline"] = cpu
    return res


# ---- N > 1 under torchrun: the same devices driven from ONE process, for the record --------------------------------------
def single_process_probe(args, R, pkg, synthetic):
    """rank 0 creates a kzg_ctx_create_multi context over all N devices (the other ranks wait on a host barrier, their GPUs
    idle) and runs the headline workload through it: the reference is a single process (src/main.rs:170)."""
    torch = R.torch
    res = None
    R.host_barrier()
    if R.rank == 0:
        try:
            devs = list(range(R.world))
            g = pkg.Context(devs)
            g.set_option("lanes", args.lanes)
            g.set_option("msm_coalesce", args.coalesce)
            g.set_option("setup_layout", 1)
            depth = args.lanes * args.coalesce
            n = 1 << LOG_N
            st = g.setup_generate(synthetic.BENCH_SECRET_BE, n)
            st.prepare_msm(args.window)
            host = torch.from_numpy(synthetic.scalar_stream(SEED, n)).pin_memory()
            sc = host.to(R.dev)
            from collections import deque

            def run(k, submit):
                q, out = deque(), None
                for _ in range(k):
                    if len(q) == depth:
                        out = g.msm_g1_wait(q.popleft())
                    q.append(submit())
                while q:
                    out = g.msm_g1_wait(q.popleft())
                return out

            steps = max(2 * depth, args.steps)
            out = {}
            for name, submit in (("value", lambda: g.msm_g1_submit_dev(st, sc.data_ptr(), n)),
                                 ("e2e", lambda: g.msm_g1_submit(st, host.data_ptr(), n))):
                run(2 * depth, submit)
                g.sync(); g.timer_start()
                t0 = time.perf_counter()
                r = run(steps, submit)
                ms = g.timer_stop()
                wall = (time.perf_counter() - t0) * 1e3
                out[name] = {"value": n / (ms / steps * 1e-3), "unit": "points/s", "ms_per_step": ms / steps, "wall_ms_per_step": wall / steps,
                             "matches_golden": r.hex() == GOLDEN_C3_COMMITMENT}
            t0 = time.perf_counter()
            r1 = g.msm_g1_dev(st, sc.data_ptr(), n)
            out["one_call_at_a_time_ms"] = (time.perf_counter() - t0) * 1e3
            out["matches_golden"] = r1.hex() == GOLDEN_C3_COMMITMENT and out["value"]["matches_golden"] and out["e2e"]["matches_golden"]
            out["what"] = (f"kzg_ctx_create_multi over {len(devs)} devices from rank 0 alone: worker thread per device, scalars resident on "
                           "device 0 (value; the other devices read their slice over NVLink) or in pinned host memory (e2e)")
            st.close()
            del host, sc
            g.close()
            res = out
        except Exception as e:          # never let the side measurement take the headline down
            res = {"error": repr(e)}
    R.host_barrier()
    return res


def run_ours(args):
    import kzg_poly_commit_exploration_b200 as pkg
    from kzg_poly_commit_exploration_b200 import synthetic

    if args.single_process and int(os.environ.get("RANK", "0")) != 0:
        return                                   # one process drives every device
    R = Ranks(args.single_process)
    devices = list(range(args.gpus)) if (args.single_process and args.gpus > 1) else [R.local]
    ctx = pkg.Context(devices if len(devices) > 1 else devices[0])
    ctx.set_option("lanes", args.lanes)
    if args.coalesce <= 0:      # auto: the smaller the share, the more a call's kernels are single partial waves — fuse four of them
        args.coalesce = 2 if ((1 << LOG_N) // max(1, args.gpus)) >= (1 << 20) else 4
    ctx.set_option("msm_coalesce", args.coalesce)

    line, pk = msm_headline(args, R, pkg, synthetic, ctx, devices)
    extra = other_configs(args, R, pkg, synthetic, ctx, devices)
    blob = blob_batch(args, R, pkg, synthetic, ctx, devices, pk) if args.blobs > 0 else None
    sp = single_process_probe(args, R, pkg, synthetic) if (R.world > 1 and not args.no_single_process_probe) else None

    if R.rank == 0:
        line["other_configs"] = extra
        if blob is not None:
            line["blob_batch"] = blob
        if sp is not None:
            line["single_process"] = sp
        if R.world == 1 and len(devices) == 1:
            threads = os.cpu_count() or 1
            cpu_points = min(1 << LOG_N, 49152 * threads)   # ~10 s of host work
            cpu_val, cpu_dt = cpu_baseline(cpu_points, threads)
            line["cpu_baseline"] = {"value": cpu_val, "unit": "points/s", "cores": threads, "kind": "port",
                                    "sample": f"{cpu_points}-point prefix of the same scalars, reference naive loop "
                                              f"(src/polynomial.rs:207-212) split over {threads} threads, {cpu_dt:.2f} s"}
        print(json.dumps(line), flush=True)
    R.torch.cuda.synchronize()
    ctx.close()
    R.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--window", type=int, default=int(os.environ.get("KZG_WINDOW", "0")))
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("KZG_LANES", "4")),
                    help="MSM calls in flight per GPU for the headline throughput (1..4)")
    ap.add_argument("--coalesce", type=int, default=int(os.environ.get("KZG_COALESCE", "0")),
                    help="consecutive submitted MSM calls run as ONE bucket problem (option msm_coalesce, 1..4; 0 = by share size)")
    ap.add_argument("--blobs", type=int, default=int(os.environ.get("KZG_BLOBS", "4096")),
                    help="size of the secondary blob batch (0 = skip)")
    ap.add_argument("--fixed-window", type=int, default=int(os.environ.get("KZG_FIXED_WINDOW", "0")),
                    help="window of the fixed-base table of the blob batch (0 = from the free device memory)")
    ap.add_argument("--alt-window", type=int, default=int(os.environ.get("KZG_ALT_WINDOW", "13")),
                    help="a second, smaller table timed beside the default on 1 GPU (0 = skip)")
    ap.add_argument("--single-process", action="store_true", default=os.environ.get("KZG_SINGLE_PROCESS", "") == "1",
                    help="drive all --gpus devices from ONE process (kzg_ctx_create_multi) instead of one process per GPU")
    ap.add_argument("--no-single-process-probe", action="store_true",
                    help="under torchrun, skip the short single-process measurement rank 0 makes over all devices")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
