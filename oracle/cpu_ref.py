"""ctypes wrapper of oracle/_build/libkzg_cpu_ref.so (C restatement of the reference loops).
TEST / BASELINE INFRASTRUCTURE ONLY — see cpu_ref.c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_build", "libkzg_cpu_ref.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        L = C.CDLL(_PATH)
        L.ref_setup_generate.restype = C.c_void_p
        L.ref_setup_generate.argtypes = [C.c_char_p, C.c_uint64]
        L.ref_setup_generate_mt.restype = C.c_void_p
        L.ref_setup_generate_mt.argtypes = [C.c_char_p, C.c_uint64, C.c_int]
        L.ref_setup_free.argtypes = [C.c_void_p]
        L.ref_setup_export.argtypes = [C.c_void_p, C.c_uint64, C.c_char_p]
        L.ref_commit.restype = C.c_int
        L.ref_commit.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_int]
        L.ref_evaluate.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p, C.c_char_p]
        L.ref_open.restype = C.c_int
        L.ref_open.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        _lib = L
    return _lib


class RefSetup:
    def __init__(self, secret_be: bytes, n: int, threads: int = 1):
        self.n = n
        self._h = lib().ref_setup_generate_mt(secret_be, n, threads)

    def export(self, i: int) -> bytes:
        o = C.create_string_buffer(48)
        lib().ref_setup_export(self._h, i, o)
        return o.raw

    def commit(self, coeffs: bytes, threads: int = 1):
        o = C.create_string_buffer(48)
        rc = lib().ref_commit(self._h, coeffs, len(coeffs) // 32, o, threads)
        return rc, o.raw

    def open(self, coeffs: bytes, z: bytes, y: bytes, threads: int = 1):
        o = C.create_string_buffer(48)
        rc = lib().ref_open(self._h, coeffs, len(coeffs) // 32, z, y, o, threads)
        return rc, o.raw

    def close(self):
        if self._h:
            lib().ref_setup_free(self._h)
            self._h = None

    def __del__(self):
        self.close()


def evaluate(coeffs: bytes, z: bytes) -> bytes:
    o = C.create_string_buffer(32)
    lib().ref_evaluate(coeffs, len(coeffs) // 32, z, o)
    return o.raw
