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
LIB_PATH = os.environ.get("B200_LIB", os.path.join(_HERE, "lib", "libb200lora.so"))  # B200_LIB: tuning variants only

B200_OK = 0
ACT_NONE = 0
ACT_GELU_TANH = 1
GEMM_AUTO = 0
GEMM_1CTA_N256 = 1
GEMM_2CTA_N256 = 2
GEMM_1CTA_N128 = 3
GEMM_1CTA_N64 = 4
GEMM_SKINNY_CLUSTER = 5
GEMM_1CTA_N160 = 6
GEMM_1CTA_N192 = 7


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
        ("trans_a", c_int32), ("trans_b", c_int32),
        ("alpha", c_float),
        ("row_alpha", c_void_p),
        ("bias", c_void_p),
        ("res", c_void_p), ("ldres", c_int32),
        ("gate", c_void_p), ("ldgate", c_int32),
        ("rows_per_sample", c_int32),
        ("aux_in", c_void_p), ("ldaux_in", c_int32),
        ("aux_out", c_void_p), ("ldaux_out", c_int32),
        ("out", c_void_p), ("ldo", c_int32),
        ("act", c_int32),
        ("act_ncols", c_int32),
        ("f32_mode", c_int32),
        ("f32_trans", c_int32),
        ("n_store", c_int32),
        ("splits", c_int32),
        ("config", c_int32),
    ]


class RepackEntry(Structure):
    """Mirror of the ``b200_repack_lora`` table entry (32 bytes)."""

    _fields_ = [("src_off", c_int64), ("dst_off", c_int64), ("rows", c_int32), ("cols", c_int32),
                ("dst_ld", c_int32), ("pad", c_int32)]


_lib = None
_lib_lock = threading.Lock()

# name -> (restype, argtypes); the single source of truth for tests/test_cabi_symbols.py
_V, _I, _F, _L = c_void_p, c_int, c_float, c_int64
SIGNATURES = {
    "b200_version": (c_int, []),
    "b200_last_error": (c_char_p, []),
    "b200_ctx_create": (c_int, [POINTER(c_void_p), c_int]),
    "b200_ctx_destroy": (c_int, [c_void_p]),
    "b200_ctx_launch_count": (c_int64, [c_void_p]),
    "b200_gemm_bf16": (c_int, [c_void_p, POINTER(GemmDesc), c_void_p]),
    "b200_lora_wgrad": (c_int, [_V, _V, _I, _V, _I, _V, _I, _V, _I, _V, _V, _I, _I, _I, _I, _I, _F, _I, _V]),
    "b200_ln_modulate_fwd": (c_int, [_V, _V, _I, _V, _V, _I, _I, _V, _I, _V, _V, _I, _I, _F, _V]),
    "b200_ln_modulate_bwd": (c_int, [_V, _V, _I, _V, _I, _V, _V, _V, _I, _I, _V, _I, _V, _I, _I, _I, _V]),
    "b200_col_reduce": (c_int, [_V, _V, _I, _V, _I, _V, _V, _V, _I, _V, _I, _V, _V, _I, _I, _I, _I, _V]),
    "b200_qk_norm_rope_fwd": (c_int, [_V, _V, _V, _V, _I, _V, _V, _V, _V, _V, _V, _V, _I, _I, _I, _I, _I, _I, _F, _V]),
    "b200_qk_norm_rope_bwd": (c_int, [_V, _V, _V, _V, _V, _V, _I, _V, _V, _V, _V, _V, _V, _V, _I, _I, _I, _I, _I, _I,
                                      _I, _F, _V]),
    "b200_silu": (c_int, [_V, _V, _V, _L, _V]),
    "b200_timestep_embed": (c_int, [_V, _V, _V, _I, _I, _F, _F, _F, _V]),
    "b200_add_bf16": (c_int, [_V, _V, _V, _V, _V, _L, _V]),
    "b200_attn_fwd": (c_int, [_V, _V, _V, _V, _V, _I, _V, _I, _V, _I, _I, _I, _I, _F, _V]),
    "b200_attn_bwd": (c_int, [_V, _V, _V, _V, _V, _I, _V, _I, _V, _I, _V, _I, _V, _V, _V, _V, _V, _V, _I, _I, _I, _I, _F, _V]),
    "b200_attn_fwd_x": (c_int, [_V, _V, _V, _V, _V, _I, _V, _I, _V, _I, _I, _I, _I, _I, _F, _V]),
    "b200_attn_bwd_x": (c_int, [_V, _V, _V, _V, _V, _I, _V, _I, _V, _I, _V, _I, _V, _V, _V, _V, _V, _V, _I, _I, _I, _I, _I, _F, _V]),
    "b200_attn_small_fwd": (c_int, [_V, _V, _I, _V, _I, _V, _I, _V, _I, _V, _I, _I, _I, _I, _I, _F, _V]),
    "b200_attn_small_bwd": (c_int, [_V, _V, _I, _V, _I, _V, _I, _V, _I, _V, _I, _V, _V, _V, _I, _V, _I, _V, _I, _I, _I, _I, _I, _I, _F, _V]),
    "b200_attn_fwd_xd": (c_int, [_V, _V, _V, _V, _V, _I, _V, _I, _V, _I, _I, _I, _I, _I, _F, _I, _V]),
    "b200_attn_bwd_xd": (c_int, [_V, _V, _V, _V, _V, _I, _V, _I, _V, _I, _V, _I, _V, _V, _V, _V, _V, _V, _I, _I, _I, _I, _I, _F, _I, _V]),
    "b200_rms_rope_fwd": (c_int, [_V, _V, _I, _V, _V, _V, _V, _V, _I, _I, _I, _I, _I, _F, _I, _V]),
    "b200_rms_rope_bwd": (c_int, [_V, _V, _V, _I, _V, _V, _V, _V, _V, _I, _I, _I, _I, _I, _I, _I, _V]),
    "b200_lora_gemv_fwd": (c_int, [_V, _V, _I, _V, _I, _V, _V, _V, _I, _F, _V, _I, _V, _I, _I, _I, _V]),
    "b200_lora_gemv_bwd": (c_int, [_V, _V, _I, _V, _I, _V, _V, _V, _I, _F, _V, _V, _V, _I, _I, _I, _V]),
    "b200_lora_gemv_fwd_rows": (c_int, [_V, _V, _I, _V, _I, _V, _V, _V, _I, _F, _V, _V, _I, _V, _I, _I, _I, _V]),
    "b200_lora_gemv_bwd_rows": (c_int, [_V, _V, _I, _V, _I, _V, _V, _V, _I, _F, _V, _V, _V, _V, _I, _I, _I, _V]),
    "b200_ddpm_add_noise": (c_int, [_V, _V, _V, _V, _V, _I, _V, _I, _L, _V]),
    "b200_train_loss": (c_int, [_V, _V, _V, _V, _V, _V, _V, _V, _V, _I, _V, _V, _V, _I, _I, _I, _I, _I, _F, _V]),
    "b200_nchw_rows": (c_int, [_V, _V, _V, _I, _I, _I, _I, _I, _V]),
    "b200_im2col": (c_int, [_V, _V, _V, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _V]),
    "b200_col2im": (c_int, [_V, _V, _V, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _V]),
    "b200_mask_rows": (c_int, [_V, _V, _I, _V, _I, _V, _I, _L, _I, _V]),
    "b200_ln_affine_fwd": (c_int, [_V, _V, _I, _V, _V, _V, _I, _V, _V, _I, _I, _F, _V]),
    "b200_ln_affine_bwd": (c_int, [_V, _V, _I, _V, _I, _V, _V, _V, _V, _I, _V, _I, _I, _I, _V]),
    "b200_groupnorm_fwd": (c_int, [_V, _V, _V, _V, _V, _V, _V, _I, _I, _I, _I, _F, _I, _V]),
    "b200_groupnorm_bwd": (c_int, [_V, _V, _V, _V, _V, _V, _V, _V, _I, _I, _I, _I, _I, _V]),
    "b200_geglu_fwd": (c_int, [_V, _V, _I, _V, _I, _L, _I, _V]),
    "b200_geglu_bwd": (c_int, [_V, _V, _I, _V, _I, _V, _I, _L, _I, _V]),
    "b200_heads_pad": (c_int, [_V, _V, _V, _I, _I, _I, _I, _I, _I, _V]),
    "b200_heads_pad3": (c_int, [_V, _V, _V, _I, _I, _V, _V, _I, _I, _V, _V, _I, _I, _I, _I, _I, _I, _I, _V]),
    "b200_flow_add_noise": (c_int, [_V, _V, _V, _V, _V, _I, _I, _I, _I, _I, _V]),
    "b200_flow_loss": (c_int, [_V, _V, _V, _V, _V, _V, _V, _I, _I, _I, _I, _I, _F, _V]),
    "b200_grad_sumsq": (c_int, [_V, _V, _L, _V, _V]),
    "b200_clip_adamw": (c_int, [_V, _V, _V, _V, _V, _V, _V, _V, _V, _L, _V, _V]),
    "b200_repack_lora": (c_int, [_V, _V, _V, _V, _I, _V]),
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


def gemm_bf16(a0, b0, out, *, a1=None, b1=None, trans_a=False, trans_b=False, alpha=1.0, row_alpha=None, bias=None,
              res=None, gate=None, rows_per_sample=0, aux_in=None, aux_out=None, act=ACT_NONE, act_ncols=0, f32_mode=0,
              f32_trans=False, n_store=0, splits=1, config=GEMM_AUTO, M=None, N=None, stream=None, ctx=None):
    """acc = op(a0) @ op(b0)^T (+ op(a1) @ op(b1)^T) with the fused epilogue of ``b200_gemm_bf16``.

    trans_a: a0 is stored [K, M];  trans_b: b0 is stored [K, N]  (see include/b200_lora.h).
    """
    ctx = ctx or Context.get(a0.device.index)
    d = GemmDesc()
    if trans_a:
        d.K0, d.M = int(a0.shape[0]), int(a0.shape[1])
    else:
        d.M, d.K0 = int(a0.shape[0]), int(a0.shape[1])
    d.N = int(b0.shape[1]) if trans_b else int(b0.shape[0])
    if M is not None:
        d.M = int(M)
    if N is not None:
        d.N = int(N)
    d.K1 = 0 if a1 is None else (int(a1.shape[0]) if trans_a else int(a1.shape[1]))
    d.A0, d.lda0 = _ptr(a0), _ld(a0)
    d.B0, d.ldb0 = _ptr(b0), _ld(b0)
    d.A1, d.lda1 = _ptr(a1), _ld(a1)
    d.B1, d.ldb1 = _ptr(b1), _ld(b1)
    d.trans_a, d.trans_b = int(bool(trans_a)), int(bool(trans_b))
    d.alpha = float(alpha)
    d.row_alpha = _ptr(row_alpha)
    d.bias = _ptr(bias)
    d.res, d.ldres = _ptr(res), _ld(res)
    d.gate, d.ldgate = _ptr(gate), _ld(gate)
    d.rows_per_sample = int(rows_per_sample)
    d.aux_in, d.ldaux_in = _ptr(aux_in), _ld(aux_in)
    d.aux_out, d.ldaux_out = _ptr(aux_out), _ld(aux_out)
    d.out = _ptr(out)
    d.ldo = int(out.stride(-2)) if out.dim() >= 2 else int(out.shape[-1])
    d.act = int(act)
    d.act_ncols = int(act_ncols)
    d.f32_mode = int(f32_mode)
    d.f32_trans = int(bool(f32_trans))
    d.n_store = int(n_store)
    d.splits = int(splits)
    d.config = int(config)
    check(ctx.lib.b200_gemm_bf16(ctx.handle, byref(d), _stream_ptr(stream)), "b200_gemm_bf16")
    return out


def call(name, *args, stream=None, ctx=None, device=None):
    """Generic entry: ``call("b200_silu", x_ptr, y_ptr, n)`` -> prepends ctx, appends the stream, checks status."""
    ctx = ctx or Context.get(device)
    fn = getattr(ctx.lib, name)
    check(fn(ctx.handle, *args, _stream_ptr(stream)), name)
