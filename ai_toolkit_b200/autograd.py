"""`torch.autograd.Function` wrapper that makes `LoRAModule.forward` a drop-in inside any eager model
(the reference's seam: toolkit/network_mixins.py:274-348).  Saves only X and Zc; backward launches the
dgrad / wgrad kernels of `linear.linear_bwd` and accumulates dA, dB straight into the network's flat
gradient buffer (so it returns no gradient tensors for them)."""
from __future__ import annotations

import torch

from . import cabi, ops
from .linear import GEMV_MAX_ROWS, RANK_PAD, linear_bwd, linear_fwd, lora_coeff


class _LoraLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, a_weight, b_weight, lora):  # a/b only tell autograd that the output depends on them
        lin = lora.org_module[0]
        out = torch.empty((x2.shape[0], lora.out_dim), device=x2.device, dtype=torch.bfloat16)
        zc = linear_fwd(lin, x2, out, lora=lora)
        ctx.lora = lora
        ctx.save_for_backward(x2, zc)
        return out

    @staticmethod
    def backward(ctx, dy):
        x2, zc = ctx.saved_tensors
        lora = ctx.lora
        lin = lora.org_module[0]
        dy = dy.contiguous()
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        linear_bwd(lin, dy, x2, zc, dx, lora=lora)
        return dx, None, None, None


class _LoraGemvFn(torch.autograd.Function):
    """M <= 8 rows (conditioning vectors): fp32 master weights, weight-streaming kernels."""

    @staticmethod
    def forward(ctx, x2, a_weight, b_weight, lora):
        lin = lora.org_module[0]
        alpha, row_alpha, _ = lora_coeff(lora, x2.shape[0])
        if row_alpha is not None:
            raise NotImplementedError("per-sample multipliers on a <= 8 row input")
        y, z = ops.lora_gemv_fwd(x2, lin.weight, lin.bias, lora.down_weight_2d(), lora.up_weight_2d(), alpha)
        ctx.lora, ctx.alpha = lora, alpha
        ctx.save_for_backward(x2, z)
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, z = ctx.saved_tensors
        lora = ctx.lora
        lin = lora.org_module[0]
        dyf = dy.float().contiguous()
        ops.lora_gemv_bwd(dyf, x2, z, lora.down_weight_2d(), lora.up_weight_2d(), ctx.alpha,
                          lora.lora_down.weight.grad.view(lora.lora_dim, lora.in_dim),
                          lora.lora_up.weight.grad.view(lora.out_dim, lora.lora_dim))
        dx = None
        if ctx.needs_input_grad[0]:
            # dX = dY W + (c dY B) A  through the tensor-core path (rows padded by TMA zero fill)
            dx = torch.empty_like(x2)
            dyb = dy.contiguous()
            t = torch.empty((dyb.shape[0], RANK_PAD), device=dy.device, dtype=torch.bfloat16)
            cabi.gemm_bf16(dyb, lora.b_pack, t, trans_b=True, alpha=ctx.alpha, config=cabi.GEMM_1CTA_N64)
            cabi.gemm_bf16(dyb, lin.weight, dx, a1=t, b1=lora.a_pack, trans_b=True)
        return dx, None, None, None


def lora_linear(lora, x):
    """y = org_forward(x) + bf16(multiplier * scale * up(down(x)))  fused; x [..., in] bf16 on a B200."""
    lin = lora.org_module[0]
    if lora.is_conv:
        raise NotImplementedError("1x1 Conv2d adapters: layout change to channels-last rows is not wired yet")
    if x.device.type != "cuda":
        raise cabi.B200Error(f"{lora.lora_name}: active LoRA on device {x.device}; the B200 path has no CPU / eager fallback")
    if x.dtype != torch.bfloat16 or lin.weight.dtype != torch.bfloat16:
        raise cabi.B200Error(f"{lora.lora_name}: the fused LoRA-Linear computes in bf16 (got x {x.dtype}, W {lin.weight.dtype})")
    net = lora.network_ref()
    net.refresh_packs()
    net.ensure_grad_views()
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    fn = _LoraGemvFn if x2.shape[0] <= GEMV_MAX_ROWS else _LoraLinearFn
    y = fn.apply(x2, lora.lora_down.weight, lora.lora_up.weight, lora)
    return y.view(*shape[:-1], lora.out_dim)
