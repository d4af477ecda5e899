"""Thin object layer over the C ABI: Context (one CUDA device) and Setup (device-resident bases)."""
from __future__ import annotations

import ctypes as C

from . import _native as N


def _ptr(x):
    """bytes / bytearray / ctypes array / int (device or pinned-host address) -> void*"""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, (bytes, bytearray)):
        return C.cast(C.c_char_p(bytes(x)), C.c_void_p)
    return C.cast(x, C.c_void_p)


class Context:
    """One CUDA device (`Context(0)`) or several devices driven from this process (`Context([0, 1, 2, 3])`,
    kzg_ctx_create_multi): the same calls then shard internally (see include/kzg_b200.h)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        if isinstance(device, (list, tuple)):
            ids = (C.c_int32 * len(device))(*device)
            st = N.lib().kzg_ctx_create_multi(C.byref(self._h), ids, len(device))
            if st != N.KZG_OK:
                raise N.KzgNativeError(st, f"kzg_ctx_create_multi(devices={list(device)}) failed: no usable CUDA devices or no peer "
                                           "access between them (no CPU fallback)")
            self.device = int(device[0])
            self.devices = [int(d) for d in device]
            return
        st = N.lib().kzg_ctx_create(C.byref(self._h), device)
        if st != N.KZG_OK:
            raise N.KzgNativeError(st, f"kzg_ctx_create(device={device}) failed: no usable CUDA device (no CPU fallback)")
        self.device = device
        self.devices = [device]

    @property
    def device_count(self) -> int:
        return int(N.lib().kzg_ctx_device_count(self._h))

    def close(self):
        if self._h:
            N.lib().kzg_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st: int):
        if st != N.KZG_OK:
            msg = N.lib().kzg_last_error(self._h).decode(errors="replace")
            raise N.KzgNativeError(st, msg or N.lib().kzg_status_string(st).decode())

    # ---- misc -------------------------------------------------------------------------
    def sync(self):
        self.check(N.lib().kzg_ctx_sync(self._h))

    @property
    def stream(self) -> int:
        """cudaStream_t of this context (an int), e.g. for torch.cuda.ExternalStream."""
        return int(N.lib().kzg_ctx_stream(self._h) or 0)

    @property
    def launches(self) -> int:
        return int(N.lib().kzg_ctx_launch_count(self._h))

    def last_kernel_ms(self, which: int = 0) -> float:
        return float(N.lib().kzg_ctx_last_kernel_ms(self._h, which))

    def timer_start(self):
        self.check(N.lib().kzg_ctx_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float()
        self.check(N.lib().kzg_ctx_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def set_option(self, name: str, value: int):
        self.check(N.lib().kzg_ctx_set_option(self._h, name.encode(), value))

    def probe_imad_peak(self) -> float:
        v = C.c_double()
        self.check(N.lib().kzg_probe_imad_peak(self._h, C.byref(v)))
        return v.value

    def probe_fq_mul(self) -> float:
        v = C.c_double()
        self.check(N.lib().kzg_probe_fq_mul(self._h, C.byref(v)))
        return v.value

    def selftest_fq_mul(self, a: bytes, b: bytes) -> bytes:
        n = len(a) // 48
        o = N.out(n * 48)
        self.check(N.lib().kzg_selftest_fq_mul(self._h, _ptr(a), _ptr(b), n, o))
        return bytes(o)

    # ---- setup ------------------------------------------------------------------------
    def setup_generate(self, secret_be: bytes, n: int, first: int = 0) -> "Setup":
        assert len(secret_be) == 32
        h = C.c_void_p()
        self.check(N.lib().kzg_setup_generate(self._h, _ptr(secret_be), first, n, C.byref(h)))
        return Setup(self, h)

    def setup_import_g1(self, g1_c48: bytes) -> "Setup":
        if len(g1_c48) % 48:
            raise N.KzgNativeError(5, f"Invalid length, expected a multiple of 48, got {len(g1_c48)}")
        h = C.c_void_p()
        self.check(N.lib().kzg_setup_import_g1(self._h, _ptr(g1_c48), len(g1_c48) // 48, C.byref(h)))
        return Setup(self, h)

    # ---- bulk operations (host buffers) -------------------------------------------------
    def commit(self, setup: "Setup", coeffs_le32, n: int | None = None) -> bytes:
        n = len(coeffs_le32) // 32 if n is None else n
        o = N.out(48)
        self.check(N.lib().kzg_commit(self._h, setup._h, _ptr(coeffs_le32), n, o))
        return bytes(o)

    def msm_g1(self, setup: "Setup", scalars_le32, n: int | None = None) -> bytes:
        n = len(scalars_le32) // 32 if n is None else n
        o = N.out(48)
        self.check(N.lib().kzg_msm_g1(self._h, setup._h, _ptr(scalars_le32), n, o))
        return bytes(o)

    def msm_g1_submit(self, setup: "Setup", scalars_le32, n: int | None = None) -> int:
        """Queue one MSM (at most `lanes` in flight, default 4); the host buffer must outlive the matching msm_g1_wait."""
        n = len(scalars_le32) // 32 if n is None else n
        t = C.c_int32(-1)
        self.check(N.lib().kzg_msm_g1_submit(self._h, setup._h, _ptr(scalars_le32), n, C.byref(t)))
        return t.value

    def msm_g1_submit_dev(self, setup: "Setup", scalars_dev: int, n: int) -> int:
        t = C.c_int32(-1)
        self.check(N.lib().kzg_msm_g1_submit_dev(self._h, setup._h, _ptr(scalars_dev), n, C.byref(t)))
        return t.value

    def msm_g1_sharded_submit_dev(self, setup: "Setup", scalars_dev: int, n: int) -> int:
        """collective (every rank, same order): partial MSM + peer exchange + fold, collected with msm_g1_wait"""
        t = C.c_int32(-1)
        self.check(N.lib().kzg_msm_g1_sharded_submit_dev(self._h, setup._h, _ptr(scalars_dev), n, C.byref(t)))
        return t.value

    def msm_g1_sharded_submit(self, setup: "Setup", scalars_le32, n: int | None = None) -> int:
        """collective, this rank's shard of the scalars in a host buffer (kept alive until the matching msm_g1_wait)"""
        n = len(scalars_le32) // 32 if n is None else n
        t = C.c_int32(-1)
        self.check(N.lib().kzg_msm_g1_sharded_submit(self._h, setup._h, _ptr(scalars_le32), n, C.byref(t)))
        return t.value

    def msm_g1_wait(self, ticket: int) -> bytes:
        o = N.out(48)
        self.check(N.lib().kzg_msm_g1_wait(self._h, ticket, o))
        return bytes(o)

    def evaluate(self, coeffs_le32, z_le: bytes, n: int | None = None) -> bytes:
        n = len(coeffs_le32) // 32 if n is None else n
        o = N.out(32)
        self.check(N.lib().kzg_evaluate(self._h, _ptr(coeffs_le32), n, _ptr(z_le), o))
        return bytes(o)

    def quotient(self, coeffs_le32, z_le: bytes, y_le: bytes | None = None, n: int | None = None) -> bytes:
        n = len(coeffs_le32) // 32 if n is None else n
        o = N.out(max(n, 1) * 32)
        ql = C.c_uint64()
        self.check(N.lib().kzg_quotient(self._h, _ptr(coeffs_le32), n, _ptr(z_le), _ptr(y_le), o, C.byref(ql)))
        return bytes(o)[: ql.value * 32]

    def open(self, setup: "Setup", coeffs_le32, z_le: bytes, y_expected_le: bytes | None = None, n: int | None = None):
        n = len(coeffs_le32) // 32 if n is None else n
        y, pr = N.out(32), N.out(48)
        self.check(N.lib().kzg_open(self._h, setup._h, _ptr(coeffs_le32), n, _ptr(z_le), _ptr(y_expected_le), y, pr))
        return bytes(y), bytes(pr)

    def commit_batch(self, setup: "Setup", coeffs_le32, blobs: int, n: int, out=None) -> bytes:
        o = N.out(blobs * 48) if out is None else out
        self.check(N.lib().kzg_commit_batch(self._h, setup._h, _ptr(coeffs_le32), blobs, n, _ptr(o)))
        return bytes(o) if out is None else out

    def open_batch(self, setup: "Setup", coeffs_le32, blobs: int, n: int, z_le32, y_out=None, proofs_out=None):
        y = N.out(blobs * 32) if y_out is None else y_out
        pr = N.out(blobs * 48) if proofs_out is None else proofs_out
        self.check(N.lib().kzg_open_batch(self._h, setup._h, _ptr(coeffs_le32), blobs, n, _ptr(z_le32), _ptr(y), _ptr(pr)))
        return (bytes(y), bytes(pr)) if y_out is None else (y, pr)

    def commit_open_batch(self, setup: "Setup", coeffs_le32, blobs: int, n: int, z_le32, commits_out=None, y_out=None, proofs_out=None):
        """commit + evaluation proof of every blob, one host->device pass over the coefficients"""
        own = commits_out is None
        cm = N.out(blobs * 48) if commits_out is None else commits_out
        y = N.out(blobs * 32) if y_out is None else y_out
        pr = N.out(blobs * 48) if proofs_out is None else proofs_out
        self.check(N.lib().kzg_commit_open_batch(self._h, setup._h, _ptr(coeffs_le32), blobs, n, _ptr(z_le32), _ptr(cm), _ptr(y), _ptr(pr)))
        return (bytes(cm), bytes(y), bytes(pr)) if own else (cm, y, pr)

    # ---- device-resident variants (addresses as ints, e.g. torch.Tensor.data_ptr()) ---------
    def msm_g1_dev(self, setup: "Setup", scalars_dev: int, n: int) -> bytes:
        o = N.out(48)
        self.check(N.lib().kzg_msm_g1_dev(self._h, setup._h, _ptr(scalars_dev), n, o))
        return bytes(o)

    def msm_g1_partial_dev(self, setup: "Setup", scalars_dev: int, n: int, partial_dev: int):
        self.check(N.lib().kzg_msm_g1_partial_dev(self._h, setup._h, _ptr(scalars_dev), n, _ptr(partial_dev)))

    def p2p_init(self, world: int, rank: int) -> bytes:
        """allocate this rank's peer mailbox; returns its CUDA IPC handle (see parallel.connect_peers)"""
        h = N.out(64)
        self.check(N.lib().kzg_p2p_init(self._h, world, rank, h))
        return bytes(h)

    def p2p_connect(self, handles: bytes):
        self.check(N.lib().kzg_p2p_connect(self._h, _ptr(handles)))

    def msm_g1_sharded_dev(self, setup: "Setup", scalars_dev: int, n: int) -> bytes:
        """collective: partial MSM of this rank's shard, exchange over peer memory, fold + compress in one kernel"""
        o = N.out(48)
        self.check(N.lib().kzg_msm_g1_sharded_dev(self._h, setup._h, _ptr(scalars_dev), n, o))
        return bytes(o)

    def g1_combine_dev(self, partials_dev: int, count: int) -> bytes:
        o = N.out(48)
        self.check(N.lib().kzg_g1_combine_dev(self._h, _ptr(partials_dev), count, o))
        return bytes(o)

    def commit_batch_dev(self, setup: "Setup", coeffs_dev: int, blobs: int, n: int, out_dev: int):
        self.check(N.lib().kzg_commit_batch_dev(self._h, setup._h, _ptr(coeffs_dev), blobs, n, _ptr(out_dev)))

    def open_batch_dev(self, setup: "Setup", coeffs_dev: int, blobs: int, n: int, z_dev: int, y_dev: int, proofs_dev: int):
        self.check(N.lib().kzg_open_batch_dev(self._h, setup._h, _ptr(coeffs_dev), blobs, n, _ptr(z_dev), _ptr(y_dev), _ptr(proofs_dev)))

    def commit_open_batch_dev(self, setup: "Setup", coeffs_dev: int, blobs: int, n: int, z_dev: int, commits_dev: int, y_dev: int, proofs_dev: int):
        self.check(N.lib().kzg_commit_open_batch_dev(self._h, setup._h, _ptr(coeffs_dev), blobs, n, _ptr(z_dev), _ptr(commits_dev),
                                                     _ptr(y_dev), _ptr(proofs_dev)))

    def evaluate_dev(self, coeffs_dev: int, n: int, z_le: bytes) -> bytes:
        o = N.out(32)
        self.check(N.lib().kzg_evaluate_dev(self._h, _ptr(coeffs_dev), n, _ptr(z_le), o))
        return bytes(o)


class Setup:
    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle

    def __len__(self):
        return int(N.lib().kzg_setup_len(self._h))

    def msm_plan(self, n: int) -> dict:
        """window geometry and halving rounds the bucket MSM uses for n scalars (after prepare_msm)"""
        p = (C.c_int32 * 5)()
        self.ctx.check(N.lib().kzg_msm_plan(self.ctx._h, self._h, n, p))
        return {"c": p[0], "W": p[1], "buckets": p[2], "rounds": p[3], "table_pitch": p[4]}

    def export_g1(self, first: int = 0, count: int | None = None) -> bytes:
        count = len(self) - first if count is None else count
        o = N.out(max(count, 1) * 48)
        self.ctx.check(N.lib().kzg_setup_export_g1(self.ctx._h, self._h, first, count, o))
        return bytes(o)[: count * 48]

    def generate_g2(self, secret_be: bytes, count: int | None = None):
        """[s^i]G2 for items 0..count of this setup (src/trusted_setup.rs:64-72), kept on the device."""
        count = len(self) if count is None else count
        self.ctx.check(N.lib().kzg_setup_generate_g2(self.ctx._h, self._h, _ptr(secret_be), count))

    def import_g2(self, g2_c96: bytes):
        if len(g2_c96) % 96:
            raise N.KzgNativeError(5, f"Invalid length, expected a multiple of 96, got {len(g2_c96)}")
        self.ctx.check(N.lib().kzg_setup_import_g2(self.ctx._h, self._h, _ptr(g2_c96), len(g2_c96) // 96))

    def g2_len(self) -> int:
        return int(N.lib().kzg_setup_g2_len(self._h))

    def export_g2(self, first: int = 0, count: int | None = None) -> bytes:
        count = self.g2_len() - first if count is None else count
        o = N.out(max(count, 1) * 96)
        self.ctx.check(N.lib().kzg_setup_export_g2(self.ctx._h, self._h, first, count, o))
        return bytes(o)[: count * 96]

    def prepare_msm(self, window_bits: int = 0):
        self.ctx.check(N.lib().kzg_setup_prepare_msm(self.ctx._h, self._h, window_bits))

    def prepare_fixed(self, n_points: int, window_bits: int = 0):
        """window_bits = 0: the widest window whose table fits the memory budget (kzg_setup_fixed_plan tells which)"""
        self.ctx.check(N.lib().kzg_setup_prepare_fixed(self.ctx._h, self._h, n_points, window_bits))

    def fixed_plan(self) -> dict:
        p = (C.c_int64 * 4)()
        self.ctx.check(N.lib().kzg_setup_fixed_plan(self._h, p))
        return {"c": int(p[0]), "W": int(p[1]), "points": int(p[2]), "table_bytes": int(p[3])}

    def close(self):
        if self._h:
            N.lib().kzg_setup_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            if self.ctx._h:
                self.close()
        except Exception:
            pass
