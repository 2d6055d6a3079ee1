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


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


def fwd(Q, K, V, o0, o1, split, scale=None, head_live=128):
    """-> lse [B, H, L] fp32 (natural log).  o0 [B*split, ld0] / o1 [B*(L-split), ld1] bf16 views (first H*128 cols)."""
    B, H, L, Dh = Q.shape
    Lk = K.shape[2]  # != L: cross attention (Wan2.1 text cross-attention)
    lse = torch.empty((B, H, L), device=Q.device, dtype=torch.float32)
    ld0 = int(o0.stride(0)) if o0 is not None else 0
    ld1 = int(o1.stride(0))
    cabi.call("b200_attn_fwd_xd", _p(Q), _p(K), _p(V), _p(o0), ld0, _p(o1), ld1, _p(lse), int(B), int(H), int(L), int(Lk),
              int(split), float(1.0 / math.sqrt(Dh) if scale is None else scale), int(head_live), device=Q.device.index)
    return lse


def bwd(Q, K, V, o0, o1, do0, do1, lse, split, scale=None, head_live=128):
    """-> dQ, dK, dV [B, H, L, 128] bf16."""
    B, H, L, Dh = Q.shape
    dQ = torch.empty_like(Q)
    dK = torch.empty_like(K)
    dV = torch.empty_like(V)
    delta = torch.empty((B, H, L), device=Q.device, dtype=torch.float32)
    dOh = torch.empty_like(Q)
    ld0 = int(o0.stride(0)) if o0 is not None else 0
    ld1 = int(o1.stride(0))
    ldd0 = int(do0.stride(0)) if do0 is not None else 0
    ldd1 = int(do1.stride(0))
    cabi.call("b200_attn_bwd_xd", _p(Q), _p(K), _p(V), _p(o0), ld0, _p(o1), ld1, _p(do0), ldd0, _p(do1), ldd1, _p(lse),
              _p(delta), _p(dOh), _p(dQ), _p(dK), _p(dV), int(B), int(H), int(L), int(K.shape[2]), int(split),
              float(1.0 / math.sqrt(Dh) if scale is None else scale), int(head_live), device=Q.device.index)
    return dQ, dK, dV


def small_fwd(q, k, v, B, H, L, Lk, head_dim, scale=None):
    """Head dims 160 / 192 / 256 (csrc/small_attn.cu): token-major q [B*L, >= H*d], k / v [B*Lk, >= H*d] views -> (o [B*L, H*d],
    lse [B, H, L])."""
    o = torch.empty((B * L, H * head_dim), device=q.device, dtype=torch.bfloat16)
    lse = torch.empty((B, H, L), device=q.device, dtype=torch.float32)
    cabi.call("b200_attn_small_fwd", _p(q), int(q.stride(0)), _p(k), int(k.stride(0)), _p(v), int(v.stride(0)), _p(o),
              int(o.stride(0)), _p(lse), int(B), int(H), int(L), int(Lk), int(head_dim),
              float(1.0 / math.sqrt(head_dim) if scale is None else scale), device=q.device.index)
    return o, lse


def small_bwd(q, k, v, o, dO, lse, dq, dk, dv, B, H, L, Lk, head_dim, scale=None):
    """Gradients into the (possibly column-sliced) token-major views dq [B*L, .], dk / dv [B*Lk, .]."""
    delta = torch.empty((B, H, L), device=q.device, dtype=torch.float32)
    cabi.call("b200_attn_small_bwd", _p(q), int(q.stride(0)), _p(k), int(k.stride(0)), _p(v), int(v.stride(0)), _p(o),
              int(o.stride(0)), _p(dO), int(dO.stride(0)), _p(lse), _p(delta), _p(dq), int(dq.stride(0)), _p(dk),
              int(dk.stride(0)), _p(dv), int(dv.stride(0)), int(B), int(H), int(L), int(Lk), int(head_dim),
              float(1.0 / math.sqrt(head_dim) if scale is None else scale), device=q.device.index)
    return dq, dk, dv
