"""`torch.autograd.Function` wrapper that makes `LoRAModule.forward` a drop-in inside any eager model
(the reference's seam: toolkit/network_mixins.py:274-348).  Saves only X and Zc; backward launches the
dgrad / wgrad kernels of `linear.linear_bwd` and accumulates dA, dB straight into the network's flat
gradient buffer (so it returns no gradient tensors for them).

Linear, 1x1 and k x k Conv2d adapters (toolkit/lora_special.py:95-111) all run the SAME fused tcgen05 GEMM: a conv is
a Linear over im2col rows [B Ho Wo, C kh kw] (`ops.im2col`, plain NCHW -> channels-last rows for 1x1 / stride 1), with the
LoRA down-projection (Conv k x k -> r) as the rank-side GEMM over the same rows and the up-projection (1x1) as the
second contraction segment of the tile.  dropout / rank_dropout (network_mixins.py:211-226) multiply the rank-side
activations Zc (forward) and T (backward) by masks drawn from torch's generator exactly as the reference draws them."""
from __future__ import annotations

import torch

from . import cabi, ops
from .linear import GEMV_MAX_ROWS, RANK_PAD, linear_bwd, linear_fwd, lora_coeff


def _dropout_masks(lora, n_rows, batch, device):
    """-> (row_mask [rows, r] fp32 | None, col_mask [batch, r] fp32 | None), scale factors included.
    Same generator consumption as `_call_forward`: F.dropout on the [.., r] activation, then torch.rand((B, r))."""
    row_mask = col_mask = None
    r = lora.lora_dim
    if lora.dropout is not None and lora.dropout > 0 and lora.training:
        row_mask = torch.nn.functional.dropout(torch.ones((n_rows, r), device=device, dtype=torch.float32), p=lora.dropout)
    if lora.rank_dropout is not None and lora.rank_dropout > 0 and lora.training:
        keep = torch.rand((batch, r), device=device) > lora.rank_dropout
        col_mask = keep.to(torch.float32) * (1.0 / (1.0 - lora.rank_dropout))  # "treat as if the rank is changed" (:224)
    return row_mask, col_mask


class _LoraLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x2, a_weight, b_weight, lora, batch):  # a/b only tell autograd that the output depends on them
        lin = lora.org_module[0]
        out = torch.empty((x2.shape[0], lora.out_dim), device=x2.device, dtype=torch.bfloat16)
        masks = _dropout_masks(lora, x2.shape[0], batch, x2.device) if lora.has_dropout() else (None, None)
        rps = x2.shape[0] // max(1, batch)
        if masks[0] is None and masks[1] is None:
            zc = linear_fwd(lin, x2, out, lora=lora)
        else:
            hook = lambda z: ops.mask_rows(z, lora.lora_dim, masks[0], masks[1], rps)  # noqa: E731
            zc = linear_fwd(lin, x2, out, lora=lora, zc_hook=hook)
        ctx.lora, ctx.masks, ctx.rps = lora, masks, rps
        ctx.save_for_backward(x2, zc)
        return out

    @staticmethod
    def backward(ctx, dy):
        x2, zc = ctx.saved_tensors
        lora = ctx.lora
        lin = lora.org_module[0]
        dy = dy.contiguous()
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        hook = None
        if ctx.masks[0] is not None or ctx.masks[1] is not None:
            hook = lambda t: ops.mask_rows(t, lora.lora_dim, ctx.masks[0], ctx.masks[1], ctx.rps)  # noqa: E731
        linear_bwd(lin, dy, x2, zc, dx, lora=lora, t_hook=hook)
        return dx, None, None, None, None


class _LoraGemvFn(torch.autograd.Function):
    """M <= 8 rows (conditioning vectors): fp32 master weights, weight-streaming kernels."""

    @staticmethod
    def forward(ctx, x2, a_weight, b_weight, lora, batch):
        lin = lora.org_module[0]
        alpha, row_alpha, rps = lora_coeff(lora, x2.shape[0])
        row_c = None
        if row_alpha is not None:  # per-sample multipliers: one coefficient per row group
            row_c = row_alpha.repeat_interleave(rps).contiguous() if rps > 1 else row_alpha
        y, z = ops.lora_gemv_fwd(x2, lin.weight, lin.bias, lora.down_weight_2d(), lora.up_weight_2d(), alpha, row_c=row_c)
        ctx.lora, ctx.alpha, ctx.row_c = lora, alpha, row_c
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
                          lora.lora_up.weight.grad.view(lora.out_dim, lora.lora_dim), row_c=ctx.row_c)
        dx = None
        if ctx.needs_input_grad[0]:
            # dX = dY W + (c dY B) A  through the tensor-core path (rows padded by TMA zero fill)
            dx = torch.empty_like(x2)
            dyb = dy.contiguous()
            t = torch.empty((dyb.shape[0], RANK_PAD), device=dy.device, dtype=torch.bfloat16)
            cabi.gemm_bf16(dyb, lora.b_pack, t, trans_b=True, alpha=ctx.alpha, row_alpha=ctx.row_c,
                           rows_per_sample=1 if ctx.row_c is not None else 0, config=cabi.GEMM_1CTA_N64)
            cabi.gemm_bf16(dyb, lin.weight, dx, a1=t, b1=lora.a_pack, trans_b=True)
        return dx, None, None, None, None


class _ConvRowsFn(torch.autograd.Function):
    """NCHW activation <-> GEMM rows: forward im2col (a plain tiled transpose for 1x1 / stride 1 / no padding), backward
    col2im.  Frozen-path plumbing of the conv-as-GEMM formulation; hand-written kernels (csrc/batch_ops.cu)."""

    @staticmethod
    def forward(ctx, x, k, s, p):
        ctx.shape, ctx.geom = tuple(x.shape), (k, s, p)
        if k == (1, 1) and s == (1, 1) and p == (0, 0):
            return ops.nchw_to_rows(x)
        return ops.im2col(x, k, s, p)

    @staticmethod
    def backward(ctx, drows):
        k, s, p = ctx.geom
        B, C, H, W = ctx.shape
        drows = drows.contiguous()
        if k == (1, 1) and s == (1, 1) and p == (0, 0):
            return ops.rows_to_nchw(drows, B, C, H, W), None, None, None
        return ops.col2im(drows, ctx.shape, k, s, p), None, None, None


class _RowsToNchwFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, B, C, H, W):
        ctx.dims = (B, C, H, W)
        return ops.rows_to_nchw(rows, B, C, H, W)

    @staticmethod
    def backward(ctx, dy):
        return ops.nchw_to_rows(dy.contiguous()), None, None, None, None


def lora_linear(lora, x):
    """y = org_forward(x) + bf16(multiplier * scale * up(down(x)))  fused; x [..., in] bf16 (Linear) or [B, C, H, W] bf16
    (Conv2d) on a B200."""
    lin = lora.org_module[0]
    if x.device.type != "cuda":
        raise cabi.B200Error(f"{lora.lora_name}: active LoRA on device {x.device}; the B200 path has no CPU / eager fallback")
    if x.dtype != torch.bfloat16 or lin.weight.dtype != torch.bfloat16:
        raise cabi.B200Error(f"{lora.lora_name}: the fused LoRA-Linear computes in bf16 (got x {x.dtype}, W {lin.weight.dtype})")
    net = lora.network_ref()
    net.refresh_packs()
    net.ensure_grad_views()
    if lora.is_conv:
        if x.dim() != 4:
            raise cabi.B200Error(f"{lora.lora_name}: Conv2d adapter expects [B, C, H, W], got {tuple(x.shape)}")
        B, C, H, W = x.shape
        k, s, p = lora.kernel_size, lora.stride, lora.padding
        Ho, Wo = ops.conv_out_hw(H, W, k, s, p)
        rows = _ConvRowsFn.apply(x.contiguous(), k, s, p)
        y = _LoraLinearFn.apply(rows, lora.lora_down.weight, lora.lora_up.weight, lora, B)
        return _RowsToNchwFn.apply(y, B, lora.out_dim, Ho, Wo)
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    if not x2.is_contiguous():
        x2 = x2.contiguous()
    fn = _LoraGemvFn if x2.shape[0] <= GEMV_MAX_ROWS else _LoraLinearFn
    y = fn.apply(x2, lora.lora_down.weight, lora.lora_up.weight, lora, int(shape[0]) if x.dim() > 1 else 1)
    return y.view(*shape[:-1], lora.out_dim)
