#!/usr/bin/env python3
"""bench.py — KZG commit MSM points/s at n = 2^20 (BASELINE.json metric; config C3), 1..8 B200.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA engine
  python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU loop (port)

A "step" is one commit-MSM of n = 2^20 synthetic random scalars U(0xC3, i) against
the powers-of-tau setup generated from the reference's bench secret (00 01 .. 1f).
  value  : points/s with the scalars already resident in HBM (device result read back, 48 B)
  e2e    : the same through the host-buffer C-ABI call kzg_msm_g1 (pinned host scalars; the
           32 MiB H2D copy and the 48 B D2H read are inside the timed region)
  N > 1  : one process per GPU; rank g owns bases/scalars [g n/N, (g+1) n/N) (point-range sharding),
           partial sums (192 B) are all-gathered with NCCL and folded on the device: strong scaling.
The folded-window table (2^(c j) P_i, built once per setup) is NOT in the timed region: it is part
of the setup object, like the setup itself.
"""
from __future__ import annotations

import argparse
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
CANONICAL_MAC_PER_POINT = 48000  # c = 16: 16 windows x 3000 (SURVEY §8d canonical figure)
# dram__bytes_read.sum + dram__bytes_write.sum of one k_seg_add<true> launch at n = 2^20, c = 16, table on the
# 128-byte pitch, from the ncu capture summarised in profiles/prof_seg_add_r01j_summary.txt (bytes per launch;
# 3.728e9 + 0.824e9 on the 96-byte pitch, profiles/prof_seg_add_r01_summary.txt)
TRAFFIC_NCU = 2.707788288e9 + 0.814427904e9


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
    return {"hbm_gbs": 6650.0}, "fallback"


def cpu_baseline(points: int, threads: int):
    """The reference's commit loop (C port, oracle/cpu_ref.c) on the host cores: bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_ref
    from kzg_poly_commit_exploration_b200 import synthetic

    st = cpu_ref.RefSetup(synthetic.BENCH_SECRET_BE, points, threads)
    coeffs = synthetic.scalar_stream(SEED, points).tobytes()
    t0 = time.perf_counter()
    rc, _ = st.commit(coeffs, threads)
    dt = time.perf_counter() - t0
    st.close()
    assert rc == 0
    return points / dt, dt


def run_reference(args):
    """--impl reference: the reference's own CPU path for this metric.  The Rust crate + blst cannot be
    built in this image (no cargo; blst un-vendored), so this is the C port of its loop
    (src/polynomial.rs:207-212) on all host threads, each step a bounded prefix of the 2^20 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 8192 * threads
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cpu_ref
    from kzg_poly_commit_exploration_b200 import synthetic

    st = cpu_ref.RefSetup(synthetic.BENCH_SECRET_BE, sample, threads)
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
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32-limb modular integer (Fq 381-bit, Fr 255-bit)",
        "data": "synthetic", "config": {"workload": f"commit MSM n=2^{LOG_N} (C3), reference naive loop on a {sample}-point prefix per step"},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port",
                         "sample": f"{sample}-point prefix of the 2^{LOG_N} workload per step, loop split over {threads} threads "
                                   "(parallelised restatement; the reference itself is single-threaded)"},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist

    import kzg_poly_commit_exploration_b200 as pkg
    from kzg_poly_commit_exploration_b200 import synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    ctx = pkg.Context(local)
    ctx.set_option("lanes", args.lanes)

    n_total = 1 << LOG_N
    per = n_total // world
    first = rank * per
    setup = ctx.setup_generate(synthetic.BENCH_SECRET_BE, per, first=first)
    t0 = time.perf_counter()
    setup.prepare_msm(args.window)
    ctx.sync()
    table_s = time.perf_counter() - t0
    plan = setup.msm_plan(per)

    host = torch.from_numpy(synthetic.scalar_stream(SEED, per, start=first)).pin_memory()
    scal_dev = host.to(dev)
    partial = torch.zeros(192, dtype=torch.uint8, device=dev)
    gathered = torch.zeros(192 * world, dtype=torch.uint8, device=dev)
    ext_stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    torch.cuda.synchronize()

    use_p2p = world > 1 and args.exchange == "p2p"
    if use_p2p:
        from kzg_poly_commit_exploration_b200 import parallel
        parallel.connect_peers(ctx, world, rank)     # NCCL carries the 64-byte IPC handles, once

    def step_dev():
        if world == 1:
            return ctx.msm_g1_dev(setup, scal_dev.data_ptr(), per)
        if use_p2p:
            # partial MSM, then ONE kernel that stores the 192-byte partial into every peer's mailbox over NVLink,
            # waits for the others, folds and compresses (msm_tail.cu k_exchange_combine): no NCCL on the data path
            return ctx.msm_g1_sharded_dev(setup, scal_dev.data_ptr(), per)
        # partial MSM, NCCL all-gather of the 192-byte partials and the fold are enqueued in order on the
        # engine's stream (torch sees it as an ExternalStream): no host synchronisation until the 48-byte result
        ctx.msm_g1_partial_dev(setup, scal_dev.data_ptr(), per, partial.data_ptr())
        with torch.cuda.stream(ext_stream):
            dist.all_gather_into_tensor(gathered, partial)
        return ctx.g1_combine_dev(gathered.data_ptr(), world)

    def step_e2e():
        if world == 1:
            return ctx.msm_g1(setup, host.data_ptr(), per)
        with torch.cuda.stream(ext_stream):
            scal_dev.copy_(host, non_blocking=True)
        return step_dev()

    def calls_in_flight(k, submit):
        # submit / wait with `lanes` calls in flight (ticket t runs on lane t mod lanes of the context): every step is a
        # complete MSM whose 48-byte result is read back; the copy and the bulk kernels of one step overlap the
        # latency-bound phases (inversion chains, bucket reduction) of the others
        from collections import deque
        q, out = deque(), None
        for _ in range(k):
            if len(q) == args.lanes:
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

    wall_ms = [0.0]     # host clock of the last timed region (max over ranks), reported beside the device time

    def timed_block(fn_k, steps):
        barrier()
        ctx.timer_start()
        t0 = time.perf_counter()
        out = fn_k(steps)
        ms_dev = ctx.timer_stop()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        # device time: CUDA events on the engine stream (every lane joins it before the stop event; the peer exchange
        # and the NCCL all-gather are enqueued on it too), max over ranks
        ms = ms_dev
        if world > 1:
            t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0].item())
            wall = float(t[1].item())
        wall_ms[0] = wall
        return ms, out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    def timed(fn, steps):
        barrier()
        ctx.timer_start()
        t0 = time.perf_counter()
        out = None
        for _ in range(steps):
            out = fn()
        ms_dev = ctx.timer_stop()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        ms = ms_dev                                # CUDA events on the engine stream, max over ranks
        if world > 1:
            t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0].item())
            wall = float(t[1].item())
        wall_ms[0] = wall
        return ms, out

    r_ref = [None]
    for _ in range(args.warmup):
        r_dev = step_dev()
    r_ref[0] = r_dev
    pipelined = world == 1 or use_p2p
    if pipelined:
        steps_dev_pipelined(max(2 * args.lanes, args.warmup))

    stop, samples = threading.Event(), []
    th = threading.Thread(target=_clock_sampler, args=(stop, samples, local), daemon=True)
    if rank == 0:
        th.start()
    launches0 = ctx.launches
    ms_single, r_dev = timed(step_dev, args.steps)          # one call at a time
    launches = ctx.launches - launches0
    ms_total = ms_single
    if pipelined:                                            # two calls in flight: the headline throughput
        ms_total, r_pipe = timed_block(steps_dev_pipelined, args.steps)
        assert r_pipe == r_dev, "pipelined and one-at-a-time results differ"
    wall_total = wall_ms[0]
    acc_ms, dom_ms = [], []
    for _ in range(min(args.steps, 5)):   # per-launch durations, CUDA events on the stream the kernels run on
        step_dev()
        ctx.sync()
        acc_ms.append(ctx.last_kernel_ms(0))   # accumulation phase (halving rounds + bucket accumulate)
        dom_ms.append(ctx.last_kernel_ms(2))   # its dominant kernel
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    ms_e2e, r_e2e = timed(step_e2e, args.steps)
    ms_e2e_sync = ms_e2e
    if pipelined:
        steps_e2e_pipelined(2 * args.lanes)
        ms_e2e, r_pipe = timed_block(steps_e2e_pipelined, args.steps)
        assert r_pipe == r_e2e, "pipelined and synchronous host-buffer results differ"
    stop.set()
    if rank == 0:
        th.join(timeout=2)
    assert r_dev == r_e2e, "device-resident and host-buffer results differ"

    imad_peak = ctx.probe_imad_peak()
    fq_peak = ctx.probe_fq_mul()
    setup.close()
    extra = other_configs(ctx, synthetic, torch, dist, dev, world, rank)
    blob = None
    if args.blobs > 0:
        blob = blob_batch(ctx, pkg, synthetic, torch, dist, dev, world, rank, args.blobs, 2, args.fixed_window)

    if rank == 0:
        ms_step = ms_total / args.steps
        value = n_total / (ms_step * 1e-3)
        e2e_value = n_total / (ms_e2e / args.steps * 1e-3)
        c = args.window or int(os.environ.get("KZG_WINDOW", "0")) or None
        peaks, peaks_src = measured_peaks()
        # dominant kernel: bucket accumulate. Algorithmic work per launch = W * per bucket additions.
        import math
        from ctypes import c_int
        acc = sorted(acc_ms)[len(acc_ms) // 2]
        dom = sorted(dom_ms)[len(dom_ms) // 2]
        windows = plan                                  # kzg_msm_plan: what the engine actually uses
        adds = windows["W"] * per                       # bucket additions of the whole phase
        rounds = windows["rounds"]                      # affine halving rounds (msm.cu auto_tree_rounds)
        tree = rounds > 0
        dom_adds = adds // 2 if tree else adds          # round 1 pairs up all entries: half of the additions
        dom_name = "k_seg_add<true> (round-1 pair additions of the affine halving, msm_tree.cu)" if tree else "k_bucket_acc"
        achieved = dom_adds * MAC_PER_BUCKET_ADD / (dom * 1e-3)
        phase_achieved = adds * MAC_PER_BUCKET_ADD / (acc * 1e-3)
        hbm_bytes = dom_adds * (2 * 96 + 2 * 4 + 48 + 96) if tree else adds * 100  # per pair: two table points, two indices, prefix in, point out
        threads = os.cpu_count() or 1
        cpu_points = min(1 << LOG_N, 49152 * threads)   # ~10 s of host work
        cpu_val, cpu_dt = cpu_baseline(cpu_points, threads) if world == 1 else (None, None)
        line = {
            "metric": "kzg_commit_msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "wall_ms_per_step": wall_total / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32-limb modular integer (Fq 381-bit, Fr 255-bit)",
            "data": "synthetic",
            "config": {"workload": f"single G1 commit-MSM n=2^{LOG_N} (BASELINE config C3), uniform scalars U(0xC3,i), "
                                   "bench-secret setup", "window_bits": windows["c"], "windows": windows["W"],
                       "sharding": f"point range, {per} points/GPU",
                       "exchange": ("none" if world == 1 else "peer stores over NVLink + fold + compress in one kernel "
                                    "(k_exchange_combine)" if use_p2p else "NCCL all-gather of 192-byte partials + fold kernel"),
                       "l2": "inputs larger than L2 (table "
                       f"{windows['W'] * per * windows['table_pitch'] / 2**20:.0f} MiB + scalars {per * 32 / 2**20:.0f} MiB per GPU)",
                       "precomputed_table_s": round(table_s, 3),
                       "calls_in_flight": args.lanes if pipelined else 1,
                       "step": (f"one complete MSM (48-byte result read back); steps are submitted {args.lanes} deep "
                                "(kzg_msm_g1_submit_dev / _wait), so the bulk kernels of one step overlap the latency-bound phases of the others"
                                if pipelined else "one complete MSM, one call at a time")},
            "one_call_at_a_time": {"value": n_total / (ms_single / args.steps * 1e-3), "unit": "points/s",
                                   "ms_per_step": ms_single / args.steps},
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": per * 32 * world,
                    "d2h_bytes_per_step": 48, "ms_per_step": ms_e2e / args.steps,
                    "call": (f"kzg_msm_g1_submit / kzg_msm_g1_wait, {args.lanes} calls in flight (copy of step i+1 beside the kernels "
                             "of the older steps)" if world == 1 else
                             f"kzg_msm_g1_sharded_submit / kzg_msm_g1_wait, {args.lanes} collective calls in flight, each rank copies "
                             "its shard of the scalars from pinned host memory" if pipelined else
                             "host -> device copy + partial MSM + all-gather + combine per step"),
                    "one_call_at_a_time": {"value": n_total / (ms_e2e_sync / args.steps * 1e-3), "unit": "points/s",
                                           "ms_per_step": ms_e2e_sync / args.steps,
                                           "call": "kzg_msm_g1" if world == 1 else "copy + kzg_msm_g1_sharded_dev"}},
            "gpu_launches": int(launches),
            "clocks": _clocks_summary(samples),
            "roofline": {
                "bound": "int32-multiply",
                "bound_note": "integer-multiply pipe (IMAD.WIDE), neither hbm nor tensor: the MSM is integer-compute-bound "
                              "(north_star); the HBM side of the dominant kernel is reported under dominant_kernel.hbm",
                "kernel": (f"bucket accumulation = k_seg_den + batch inversion + k_seg_add ({rounds} affine halving rounds, "
                           "msm_tree.cu) + k_bucket_acc" if tree else "k_bucket_acc"),
                "achieved": phase_achieved / 1e12, "peak": imad_peak / 1e12, "unit": "T wide-MAC/s",
                "frac": phase_achieved / imad_peak,
                "work": f"algorithmic: {adds} bucket additions x 3000 wide-MAC (XYZZ += affine = 10 Fq products, SURVEY 8d)",
                "peak_source": "kzg_probe_imad_peak (independent mad.wide.u32 chains) in this run",
                "kernel_ms": acc, "kernel_share_of_step": acc / (ms_single / args.steps),
                "share_note": "shares are of the one-call-at-a-time step (kernel durations are measured without overlap)",
                "dominant_kernel": {
                    "name": dom_name, "ms": dom, "share_of_step": dom / (ms_single / args.steps),
                    "as_implemented": ("5 Fq products = 1500 wide-MAC per pair addition (denominators and the shared "
                                       "inversion run in k_seg_den / k_inv_*)" if tree else "10 Fq products = 3000 wide-MAC per addition"),
                    "achieved": dom_adds * (1500 if tree else 3000) / (dom * 1e-3) / 1e12,
                    "frac": dom_adds * (1500 if tree else 3000) / (dom * 1e-3) / imad_peak,
                    "hbm": {"achieved": hbm_bytes / (dom * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                            "frac": hbm_bytes / (dom * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0), "peak_source": peaks_src,
                            "algorithmic_bytes": hbm_bytes}},
                "fq_mul_per_s_probe": fq_peak,
                "canonical_frac": (n_total / world) * CANONICAL_MAC_PER_POINT / (ms_step * 1e-3) / imad_peak,
                "traffic": TRAFFIC_NCU if tree and world == 1 else None,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one k_seg_add<true> launch (the dominant "
                                "kernel), profiles/prof_seg_add_r01j_summary.txt"},
        }
        line["other_configs"] = extra
        if blob is not None:
            line["blob_batch"] = blob
        if cpu_val is not None:
            line["cpu_baseline"] = {"value": cpu_val, "unit": "points/s", "cores": threads, "kind": "port",
                                    "sample": f"{cpu_points}-point prefix of the same scalars, reference naive loop "
                                              f"(src/polynomial.rs:207-212) split over {threads} threads, {cpu_dt:.2f} s"}
        print(json.dumps(line), flush=True)
    # tensors that were used on the engine's stream (through the ExternalStream) must go before the stream does:
    # torch's pinned-memory allocator records an event on every stream a block was used on when the block is freed
    del host, scal_dev, partial, gathered
    torch.cuda.synchronize()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def other_configs(ctx, synthetic, torch, dist, dev, world, rank):
    """BASELINE configs C5 (powers-of-tau generation of 2^22 G1 points, sharded by index range) and C2 (degree
    2^16 commit + evaluate + open on one GPU), timed once each beside the headline; parity for both is in tests/."""
    out = {}
    n5 = 1 << 22
    per = n5 // world
    st = ctx.setup_generate(synthetic.BENCH_SECRET_BE, 1 << 12)   # builds the generator table (once per ctx)
    st.close()
    dts = []
    for rep in range(3):                      # the first pass also pays for the scratch allocations (~1.5 GB)
        if rep:
            st.close()
        if world > 1:
            dist.barrier()
        ctx.sync()
        t0 = time.perf_counter()
        st = ctx.setup_generate(synthetic.BENCH_SECRET_BE, per, first=rank * per)
        ctx.sync()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        dts.append(dt)
    dt = sorted(dts)[1]                       # median of 3
    t0 = time.perf_counter()
    n_g2 = min(per, 1 << 18)
    st.generate_g2(synthetic.BENCH_SECRET_BE, n_g2)         # builds the G2 generator table (once per ctx) and the scratch
    ctx.sync()
    t0 = time.perf_counter()
    st.generate_g2(synthetic.BENCH_SECRET_BE, n_g2)
    ctx.sync()
    dt_g2 = time.perf_counter() - t0
    st.close()
    out["setup_2^22_g1"] = {"metric": "setup_points_per_s", "value": n5 / dt, "unit": "G1 points/s", "ms": dt * 1e3,
                            "sharding": f"index range, {per} points/GPU", "wide_mac_per_point": 48000,
                            "path": "affine pairwise trees over the generator table (fixed_base_points_tree), 90 Fq products per point"
                                    if per >= (1 << 18) else "XYZZ table walk + batched to-affine",
                            "g2_points_per_s_per_gpu": n_g2 / dt_g2, "g2_points_timed": n_g2}
    if world == 1:
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
    return out


def blob_batch(ctx, pkg, synthetic, torch, dist, dev, world, rank, blobs_total, steps, fixed_window):
    """Secondary metric (BASELINE config C4): 4096 blobs x 4096 coefficients, commit + evaluation proof per
    blob, sharded by BLOB across ranks (no data-path collective); blob (commit+proof)/s."""
    n = 4096
    first, count = rank * (blobs_total // world), blobs_total // world
    setup = ctx.setup_generate(synthetic.BENCH_SECRET_BE, n)
    t0 = time.perf_counter()
    setup.prepare_fixed(n, fixed_window)
    ctx.sync()
    table_s = time.perf_counter() - t0
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

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(); ctx.sync()
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        ctx.sync(); torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    step_dev()
    ms_dev = timed(step_dev, steps) / steps
    step_e2e()
    ms_e2e = timed(step_e2e, steps) / steps
    same = bytes(commits.cpu().numpy().tobytes()) == bytes(out_c.numpy().tobytes()) and \
        bytes(proofs.cpu().numpy().tobytes()) == bytes(out_p.numpy().tobytes())
    setup.close()
    W = (255 + fixed_window - 1) // fixed_window   # windows_for (csrc/engine.cuh)
    return {"metric": "blob_commit_plus_proof_per_s", "value": blobs_total / (ms_dev * 1e-3), "unit": "blobs/s",
            "ms_per_batch": ms_dev, "e2e": {"value": blobs_total / (ms_e2e * 1e-3), "unit": "blobs/s", "ms_per_batch": ms_e2e,
                                            "h2d_bytes_per_batch": blobs_total * n * 32 + blobs_total * 32,
                                            "call": "kzg_commit_open_batch",
                                            "d2h_bytes_per_batch": blobs_total * 128},
            "config": {"workload": f"{blobs_total} blobs x {n} coefficients (BASELINE config C4), commit + proof, sharded by blob",
                       "fixed_base_window_bits": fixed_window, "lookups_per_coefficient": W,
                       "table_gib": round(n * W * (1 << (fixed_window - 1)) * 96 / 2**30, 2), "table_build_s": round(table_s, 2)},
            "dev_equals_e2e": bool(same),
            "scaling": f"strong ({blobs_total} blobs in total, {count} per GPU, no data-path collective)" if world > 1 else "n/a"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--window", type=int, default=int(os.environ.get("KZG_WINDOW", "0")))
    ap.add_argument("--exchange", default=os.environ.get("KZG_EXCHANGE", "p2p"), choices=["p2p", "nccl"],
                    help="multi-GPU exchange of the partial sums: fused peer-memory kernel (default) or NCCL all-gather")
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("KZG_LANES", "4")),
                    help="MSM calls in flight per GPU for the headline throughput (1..4)")
    ap.add_argument("--blobs", type=int, default=int(os.environ.get("KZG_BLOBS", "4096")),
                    help="size of the secondary blob batch (0 = skip)")
    ap.add_argument("--fixed-window", type=int, default=int(os.environ.get("KZG_FIXED_WINDOW", "15")))
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
