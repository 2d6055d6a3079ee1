"""ctypes binding of the C ABI in ``include/b200_lora.h`` (``ai_toolkit_b200/lib/libb200lora.so``).

PyTorch is used only as plumbing here: it owns device memory and streams; every compute call goes
through the raw-pointer C entry points.  There is no CPU or library fallback: when the shared library
is missing, or the device is not a B200, the calls raise ``B200Error``.
"""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200lora.so")

B200_OK = 0
ACT_NONE = 0
ACT_GELU_TANH = 1
GEMM_AUTO = 0
GEMM_1CTA_N256 = 1
GEMM_2CTA_N256 = 2
GEMM_1CTA_N128 = 3
GEMM_1CTA_N64 = 4


class B200Error(RuntimeError):
    pass


class GemmDesc(Structure):
    """Mirror of ``b200_gemm_desc`` (include/b200_lora.h)."""

    _fields_ = [
        ("M", c_int32), ("N", c_int32), ("K0", c_int32), ("K1", c_int32),
        ("A0", c_void_p), ("lda0", c_int32),
        ("B0", c_void_p), ("ldb0", c_int32),
        ("A1", c_void_p), ("lda1", c_int32),
        ("B1", c_void_p), ("ldb1", c_int32),
        ("bias", c_void_p),
        ("res", c_void_p), ("ldres", c_int32),
        ("gate", c_void_p), ("ldgate", c_int32),
        ("rows_per_sample", c_int32),
        ("aux_in", c_void_p), ("ldaux_in", c_int32),
        ("aux_out", c_void_p), ("ldaux_out", c_int32),
        ("out", c_void_p), ("ldo", c_int32),
        ("act", c_int32),
        ("out_f32", c_int32),
        ("splits", c_int32),
        ("config", c_int32),
    ]


_lib = None
_lib_lock = threading.Lock()

# name -> (restype, argtypes); the single source of truth for tests/test_cabi_symbols.py
SIGNATURES = {
    "b200_version": (c_int, []),
    "b200_last_error": (c_char_p, []),
    "b200_ctx_create": (c_int, [POINTER(c_void_p), c_int]),
    "b200_ctx_destroy": (c_int, [c_void_p]),
    "b200_ctx_launch_count": (c_int64, [c_void_p]),
    "b200_gemm_bf16": (c_int, [c_void_p, POINTER(GemmDesc), c_void_p]),
}


def load_library():
    """Load the shared library (no GPU needed) and declare every signature."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                f"{LIB_PATH} is missing: build it with `make` (or __graft_entry__.build()). "
                "ai_toolkit_b200 has no CPU / PyTorch fallback for its hot path."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def last_error() -> str:
    lib = load_library()
    msg = lib.b200_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = ""):
    if rc != B200_OK:
        raise B200Error(f"{what} failed with code {rc}: {last_error()}")


class Context:
    """One ``b200_ctx`` per (process, device)."""

    _instances = {}

    def __init__(self, device: int):
        lib = load_library()
        h = c_void_p()
        check(lib.b200_ctx_create(byref(h), int(device)), "b200_ctx_create")
        self.handle = h
        self.device = int(device)
        self.lib = lib

    @classmethod
    def get(cls, device: int | None = None) -> "Context":
        import torch

        if device is None:
            device = torch.cuda.current_device()
        device = int(device)
        ctx = cls._instances.get(device)
        if ctx is None:
            with torch.cuda.device(device):
                ctx = cls(device)
            cls._instances[device] = ctx
        return ctx

    def launch_count(self) -> int:
        return int(self.lib.b200_ctx_launch_count(self.handle))


def _stream_ptr(stream=None):
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return c_void_p(s.cuda_stream)


def _ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def _ld(t):
    return 0 if t is None else int(t.stride(0))


def gemm_bf16(a0, b0, out, *, a1=None, b1=None, bias=None, res=None, gate=None, rows_per_sample=0, aux_in=None,
              aux_out=None, act=ACT_NONE, out_f32=False, splits=1, config=GEMM_AUTO, stream=None, ctx=None):
    """out[M,N] = a0[M,K0] @ b0[N,K0]^T (+ a1 @ b1^T) with the fused epilogue of ``b200_gemm_bf16``."""
    ctx = ctx or Context.get(a0.device.index)
    d = GemmDesc()
    d.M, d.K0 = int(a0.shape[0]), int(a0.shape[1])
    d.N = int(b0.shape[0])
    d.K1 = 0 if a1 is None else int(a1.shape[1])
    d.A0, d.lda0 = _ptr(a0), _ld(a0)
    d.B0, d.ldb0 = _ptr(b0), _ld(b0)
    d.A1, d.lda1 = _ptr(a1), _ld(a1)
    d.B1, d.ldb1 = _ptr(b1), _ld(b1)
    d.bias = _ptr(bias)
    d.res, d.ldres = _ptr(res), _ld(res)
    d.gate, d.ldgate = _ptr(gate), _ld(gate)
    d.rows_per_sample = int(rows_per_sample)
    d.aux_in, d.ldaux_in = _ptr(aux_in), _ld(aux_in)
    d.aux_out, d.ldaux_out = _ptr(aux_out), _ld(aux_out)
    d.out = _ptr(out)
    d.ldo = int(out.stride(-2))
    d.act = int(act)
    d.out_f32 = 1 if out_f32 else 0
    d.splits = int(splits)
    d.config = int(config)
    check(ctx.lib.b200_gemm_bf16(ctx.handle, byref(d), _stream_ptr(stream)), "b200_gemm_bf16")
    return out
