"""Joint (text || image) attention of the DiT blocks: softmax(Q K^T / sqrt(d)) V over head-major
[B, H, L, 128] operands, output written token-major straight into the operand buffers of the following
projection GEMMs (rows < split -> o0, the text stream; rows >= split -> o1, the image stream / the
single-stream concat buffer), plus the log-sum-exp rows needed by the backward.

Backends: the hand-written tcgen05 flash-attention kernels (`b200_attn_fwd` / `b200_attn_bwd`).
"""
from __future__ import annotations

import math
from ctypes import c_void_p

import torch

from . import cabi


import os

# Bring-up aid ONLY (never the default, never used by bench.py / tests of the product path): route the
# attention core through torch SDPA to bisect engine bugs from attention-kernel bugs.
_DEBUG_TORCH = os.environ.get("B200_ATTN_DEBUG_TORCH", "0") == "1"


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


def _dbg_gather(o0, o1, B, H, L, split):
    D = H * 128
    parts = []
    if split > 0:
        parts.append(o0[:, :D].reshape(B, split, D))
    parts.append(o1[:, :D].reshape(B, L - split, D))
    return torch.cat(parts, 1)


def _dbg_fwd(Q, K, V, o0, o1, split):
    B, H, L, Dh = Q.shape
    o = torch.nn.functional.scaled_dot_product_attention(Q, K, V).transpose(1, 2).reshape(B, L, H * Dh)
    if split > 0:
        o0[:, :H * Dh].copy_(o[:, :split].reshape(-1, H * Dh))
    o1[:, :H * Dh].copy_(o[:, split:].reshape(-1, H * Dh))
    return torch.zeros((B, H, L), device=Q.device, dtype=torch.float32)


def _dbg_bwd(Q, K, V, do0, do1, split):
    B, H, L, Dh = Q.shape
    do = _dbg_gather(do0, do1, B, H, L, split).reshape(B, L, H, Dh).transpose(1, 2)
    q, k, v = (t.detach().clone().requires_grad_(True) for t in (Q, K, V))
    with torch.enable_grad():
        o = torch.nn.functional.scaled_dot_product_attention(q, k, v)
    return torch.autograd.grad(o, (q, k, v), do)


def fwd(Q, K, V, o0, o1, split):
    """-> lse [B, H, L] fp32 (natural log).  o0 [B*split, ld0] / o1 [B*(L-split), ld1] bf16 views (first H*128 cols)."""
    B, H, L, Dh = Q.shape
    if _DEBUG_TORCH:
        return _dbg_fwd(Q, K, V, o0, o1, split)
    lse = torch.empty((B, H, L), device=Q.device, dtype=torch.float32)
    ld0 = int(o0.stride(0)) if o0 is not None else 0
    ld1 = int(o1.stride(0))
    cabi.call("b200_attn_fwd", _p(Q), _p(K), _p(V), _p(o0), ld0, _p(o1), ld1, _p(lse), int(B), int(H), int(L), int(split),
              float(1.0 / math.sqrt(Dh)), device=Q.device.index)
    return lse


def bwd(Q, K, V, o0, o1, do0, do1, lse, split):
    """-> dQ, dK, dV [B, H, L, 128] bf16."""
    B, H, L, Dh = Q.shape
    if _DEBUG_TORCH:
        return _dbg_bwd(Q, K, V, do0, do1, split)
    dQ = torch.empty_like(Q)
    dK = torch.empty_like(K)
    dV = torch.empty_like(V)
    delta = torch.empty((B, H, L), device=Q.device, dtype=torch.float32)
    dOh = torch.empty_like(Q)
    ld0 = int(o0.stride(0)) if o0 is not None else 0
    ld1 = int(o1.stride(0))
    ldd0 = int(do0.stride(0)) if do0 is not None else 0
    ldd1 = int(do1.stride(0))
    cabi.call("b200_attn_bwd", _p(Q), _p(K), _p(V), _p(o0), ld0, _p(o1), ld1, _p(do0), ldd0, _p(do1), ldd1, _p(lse),
              _p(delta), _p(dOh), _p(dQ), _p(dK), _p(dV), int(B), int(H), int(L), int(split), float(1.0 / math.sqrt(Dh)),
              device=Q.device.index)
    return dQ, dK, dV
