"""The adapter-bearing part of the SD1.5 / SDXL UNet on the B200 kernels: `Transformer2DModel` (GroupNorm -> proj_in ->
N x BasicTransformerBlock -> proj_out -> + residual), the reference's DEFAULT LoRA target for UNets
(`LoRANetwork.UNET_TARGET_REPLACE_MODULE = ["Transformer2DModel"]`, toolkit/kohya_lora.py:750; every Linear and 1x1 conv
inside it is wrapped, kohya names `lora_unet_<path>_attn1_to_q` ..., `.alpha` saved, scale = alpha / rank).

The containers below carry diffusers' class / module / parameter names (so the network discovers the same modules under a
UNet whose other blocks are ordinary eager modules); `Transformer2DFunction` is ONE autograd node per Transformer2DModel
call: its forward is a hand-scheduled launch sequence (fused LoRA GEMMs with bias / residual epilogues, LayerNorm-affine,
zero-padded head re-layout + the head-dim-128 tcgen05 attention kernels for self- and cross-attention, GEGLU) that keeps the
activations the backward needs; its backward returns dX and accumulates dA / dB into the network's flat gradient buffer.
SD1.5's 1x1-conv `proj_in` / `proj_out` and SDXL's Linear ones (`use_linear_projection`) are the same GEMM over channels-last
rows.  Heads of 40 / 64 / 80 columns are served by zero padding to 128 (SD1.5's 160-wide heads of the deepest level are not:
`NotImplementedError`).

What is NOT here: the frozen ResNet / down / up-sampling body of the UNet (no adapters with the default target list).  It runs
as eager PyTorch around these blocks -- a UNet engine would need GroupNorm-SiLU-conv fusions on top
of the kernels in csrc/unet_ops.cu and csrc/batch_ops.cu.
"""
from __future__ import annotations

import math
import weakref

import torch
import torch.nn as nn

from . import attention, cabi, ops
from .linear import linear_bwd, linear_fwd, live_lora, shared_input_bwd, shared_input_fwd


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container of the B200 Transformer2D engine; call Transformer2DModel.forward")


class Attention(_Holder):
    def __init__(self, query_dim, cross_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])


class GEGLU(_Holder):
    def __init__(self, d_in, d_out):
        super().__init__()
        self.proj = nn.Linear(d_in, d_out * 2)


class FeedForward(_Holder):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])


class BasicTransformerBlock(_Holder):
    def __init__(self, dim, heads, dim_head, cross_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)


class Transformer2DModel(nn.Module):
    """Same class name as diffusers': the reference's default `target_lin_modules` selects it by name."""

    def __init__(self, heads, dim_head, in_channels, num_layers=1, cross_dim=768, use_linear_projection=False, groups=32,
                 device=None, dtype=torch.bfloat16):
        super().__init__()
        if (dim_head > 128 and dim_head not in (160, 192, 256)) or dim_head % 4:
            raise NotImplementedError(f"attention head dim {dim_head}: the tcgen05 attention kernels are built for 128 and serve "
                                      "smaller heads by zero padding; 160 / 192 / 256 (SD1.5's deepest levels) run on the CUDA-core "
                                      "kernel (csrc/small_attn.cu)")
        inner = heads * dim_head
        self.heads, self.dim_head, self.inner, self.in_channels = heads, dim_head, inner, in_channels
        self.use_linear_projection = use_linear_projection
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            with torch.device(device if device is not None else "cpu"):
                self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
                self.proj_in = nn.Linear(in_channels, inner) if use_linear_projection else nn.Conv2d(in_channels, inner, 1)
                self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_dim)
                                                         for _ in range(num_layers)])
                self.proj_out = nn.Linear(inner, in_channels) if use_linear_projection else nn.Conv2d(inner, in_channels, 1)
        finally:
            torch.set_default_dtype(prev)
        self.requires_grad_(False)

    def forward(self, hidden_states, encoder_hidden_states=None, **kwargs):
        """[B, C, H, W] bf16, context [B, Lc, cross_dim] bf16 -> [B, C, H, W] (diffusers returns a tuple / dataclass; the
        reference's UNet blocks index `[0]`, so a 1-tuple is returned when `return_dict=False` is passed)."""
        anchor = _anchor(self)
        # (grad mode is off inside autograd.Function.forward, so the decision to keep activations is taken here)
        save = torch.is_grad_enabled() and (hidden_states.requires_grad or anchor is not None)
        out = Transformer2DFunction.apply(hidden_states, encoder_hidden_states, self, anchor, save)
        return (out,) if kwargs.get("return_dict") is False else out


def adopt_unet_transformers(unet: nn.Module) -> int:
    """Replace every diffusers-style `Transformer2DModel` of a LOADED UNet (any module whose class has that name and the
    diffusers attribute layout: norm / proj_in / transformer_blocks[i].{norm1, attn1, norm2, attn2, norm3, ff.net} / proj_out) by
    this package's container over the SAME tensors (`load_state_dict(assign=True)`: no second copy).  The rest of the UNet stays
    the caller's eager model; kohya adapter names do not change (same class name, same attribute paths).  Returns the count."""
    count = 0
    for parent in list(unet.modules()):
        for name, child in list(parent.named_children()):
            if type(child).__name__ != "Transformer2DModel" or isinstance(child, Transformer2DModel):
                continue
            blk0 = child.transformer_blocks[0]
            heads = int(blk0.attn1.heads)
            inner = int(blk0.attn1.to_q.weight.shape[0])
            w = child.proj_in.weight
            with torch.device("meta"):
                new = Transformer2DModel(heads, inner // heads, int(child.norm.num_channels), len(child.transformer_blocks),
                                         cross_dim=int(blk0.attn2.to_k.weight.shape[1]),
                                         use_linear_projection=isinstance(child.proj_in, nn.Linear), groups=int(child.norm.num_groups),
                                         dtype=w.dtype)
            if abs(float(child.norm.eps) - 1e-6) > 1e-12:
                raise NotImplementedError(f"Transformer2DModel GroupNorm eps {child.norm.eps}: the engine's container assumes 1e-6")
            new.load_state_dict(child.state_dict(), assign=True)
            new.requires_grad_(False)
            setattr(parent, name, new)
            count += 1
    return count


def _anchor(model):
    """A leaf that requires grad (any live adapter weight), so that autograd calls the backward even when the input of the
    first Transformer2DModel of a UNet carries no gradient; None when no adapter is live."""
    for mod in model.modules():
        ref = getattr(mod, "_b200_lora", None)
        lora = ref() if ref is not None else None
        if lora is not None and lora.is_live():
            return lora.lora_down.weight
    return None


def _empty(shape, like, dtype=torch.bfloat16):
    return torch.empty(shape, device=like.device, dtype=dtype)


def _w2(lin):
    return lin.weight if lin.weight.dim() == 2 else lin.weight.view(lin.weight.shape[0], -1)


class Transformer2DEngine:
    """Launch schedules of one Transformer2DModel (stateless: the saved activations travel through the autograd ctx)."""

    @staticmethod
    def _prepare(model):
        for mod in model.modules():
            ref = getattr(mod, "_b200_lora", None)
            lora = ref() if ref is not None else None
            if lora is not None and lora.is_live():
                net = lora.network_ref()
                if any(m.has_dropout() or (m.module_dropout and m.training) for m in net.get_all_modules()):
                    raise NotImplementedError("dropout variants with the fused Transformer2D engine: use the per-module path")
                done = getattr(model, "_b200_groups_for", None)
                if done is None or done() is not net:
                    # adapters that share an input: to_q / to_k / to_v of the self attention and to_k / to_v of the cross
                    # attention -> one rank-side product and one fused GEMM per group (ranks that are multiples of 8)
                    groups = []
                    for blk in model.transformer_blocks:
                        groups.append([live_lora(l) for l in (blk.attn1.to_q, blk.attn1.to_k, blk.attn1.to_v)])
                        groups.append([live_lora(l) for l in (blk.attn2.to_k, blk.attn2.to_v)])
                    net.register_fused_groups(groups)
                    model._b200_groups_for = weakref.ref(net)
                net.refresh_packs()
                net.ensure_grad_views()
                return

    @staticmethod
    def _attn_fwd(a, x, ctx, B, L, Lc, res, self_attn):
        """x [B L, C] queries; ctx [B Lc, Dc] keys / values source (x itself for self-attention) -> (out = res + to_out(attn), saved)"""
        H, d = a.heads, a.dim_head
        inner = H * d
        if self_attn:
            qkv = _empty((B * L, 3 * inner), x)
            z_qkv = shared_input_fwd((a.to_q, a.to_k, a.to_v), x, qkv)
            q, k, v = qkv[:, :inner], qkv[:, inner:2 * inner], qkv[:, 2 * inner:]
            z_q = None
        else:
            q = _empty((B * L, inner), x)
            z_q = linear_fwd(a.to_q, x, q, lora=live_lora(a.to_q))
            kv = _empty((B * Lc, 2 * inner), x)
            z_qkv = shared_input_fwd((a.to_k, a.to_v), ctx, kv)
            k, v = kv[:, :inner], kv[:, inner:]
        scale = 1.0 / math.sqrt(d)
        if d > 128:  # SD1.5's 160-wide heads (deepest levels, <= 1024 tokens): CUDA-core kernel on the token-major projections
            o, lse = attention.small_fwd(q, k, v, B, H, L, Lc, d, scale=scale)
            Q = K = V = o_pad = None
            live = d
            small = (q, k, v)
        else:
            Q = _empty((B, H, L, 128), x)
            K = _empty((B, H, Lc, 128), x)
            V = _empty((B, H, Lc, 128), x)
            ops.heads_pad_multi([(q, Q), (k, K), (v, V)], B, d)
            o_pad = _empty((B * L, H * 128), x)  # the attention kernels write 128 columns per head, token-major
            live = 64 if d <= 64 else 128  # zero-padded half heads: the kernels skip the padding
            lse = attention.fwd(Q, K, V, None, o_pad, 0, scale=scale, head_live=live)
            if d == 128:
                o = o_pad
            else:  # drop the zero columns: [B L, H, 128] -> [B L, H d]
                o = o_pad.view(B * L, H, 128)[:, :, :d].reshape(B * L, inner)
            small = None
        out = _empty(res.shape, x)
        z_o = linear_fwd(a.to_out[0], o, out, lora=live_lora(a.to_out[0]), res=res)
        return out, dict(x=x, ctx=ctx, z_q=z_q, z_qkv=z_qkv, Q=Q, K=K, V=V, o_pad=o_pad, o=o, lse=lse, z_o=z_o, scale=scale, live=live, small=small)

    @staticmethod
    def _attn_bwd(a, s, dout, B, L, Lc, self_attn):
        """-> dx [B L, C] through the queries (and, for self-attention, keys / values)"""
        H, d = a.heads, a.dim_head
        inner = H * d
        do = _empty((B * L, inner), dout)
        linear_bwd(a.to_out[0], dout, s["o"], s["z_o"], do, lora=live_lora(a.to_out[0]))
        dx = _empty(s["x"].shape, dout)
        if self_attn:  # keys / values come from the same x: one dgrad GEMM over the concatenated [dq | dk | dv]
            dqkv = _empty((B * L, 3 * inner), dout)
            dq, dk, dv = dqkv[:, :inner], dqkv[:, inner:2 * inner], dqkv[:, 2 * inner:]
        else:  # cross-attention: the text embeddings are data -> adapters only on the key / value side
            dq = _empty((B * L, inner), dout)
            dkv = _empty((B * Lc, 2 * inner), dout)
            dk, dv = dkv[:, :inner], dkv[:, inner:]
        if s["small"] is not None:
            q, k, v = s["small"]
            attention.small_bwd(q, k, v, s["o"], do, s["lse"], dq, dk, dv, B, H, L, Lc, d, scale=s["scale"])
        else:
            if d == 128:
                do_pad = do
            else:
                do_pad = torch.zeros((B * L, H * 128), device=dout.device, dtype=torch.bfloat16)
                do_pad.view(B * L, H, 128)[:, :, :d].copy_(do.view(B * L, H, d))
            dQ, dK, dV = attention.bwd(s["Q"], s["K"], s["V"], None, s["o_pad"], None, do_pad, s["lse"], 0, scale=s["scale"],
                                       head_live=s["live"])
            ops.heads_pad_multi([(dq, dQ), (dk, dK), (dv, dV)], B, d, to_heads=False)
        if self_attn:
            shared_input_bwd((a.to_q, a.to_k, a.to_v), dqkv, s["x"], s["z_qkv"], dx)
        else:
            linear_bwd(a.to_q, dq, s["x"], s["z_q"], dx, lora=live_lora(a.to_q))
            shared_input_bwd((a.to_k, a.to_v), dkv, s["ctx"], s["z_qkv"], None)
        return dx

    @classmethod
    def forward(cls, model, x, context, save=True):
        if x.device.type != "cuda":
            raise cabi.B200Error("Transformer2DModel (B200 engine) needs a B200; there is no CPU / eager fallback")
        cls._prepare(model)
        B, C, Hh, Ww = x.shape
        L = Hh * Ww
        Lc = context.shape[1]
        x = x.contiguous()
        ctx2 = context.reshape(B * Lc, context.shape[2]).contiguous()
        gn, gmean, grstd = ops.groupnorm_fwd(x, model.norm.weight, model.norm.bias, model.norm.num_groups, model.norm.eps)
        rows = ops.nchw_to_rows(gn)                    # [B L, C] channels-last rows
        h = _empty((B * L, model.inner), x)
        z_in = linear_fwd(model.proj_in, rows, h, lora=live_lora(model.proj_in))  # 1x1 conv == Linear over the rows
        blocks = []
        for blk in model.transformer_blocks:
            n1, m1, r1 = ops.ln_affine_fwd(h, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps)
            h1, s1 = cls._attn_fwd(blk.attn1, n1, n1, B, L, L, h, True)
            n2, m2, r2 = ops.ln_affine_fwd(h1, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps)
            h2, s2 = cls._attn_fwd(blk.attn2, n2, ctx2, B, L, Lc, h1, False)
            n3, m3, r3 = ops.ln_affine_fwd(h2, blk.norm3.weight, blk.norm3.bias, blk.norm3.eps)
            ff1, ff2 = blk.ff.net[0].proj, blk.ff.net[2]
            proj = _empty((B * L, ff1.out_features), x)
            z_f1 = linear_fwd(ff1, n3, proj, lora=live_lora(ff1))
            act = ops.geglu_fwd(proj)
            h3 = _empty(h.shape, x)
            z_f2 = linear_fwd(ff2, act, h3, lora=live_lora(ff2), res=h2)
            if save:
                blocks.append(dict(h=h, n1=n1, m1=m1, r1=r1, s1=s1, h1=h1, n2=n2, m2=m2, r2=r2, s2=s2, h2=h2, n3=n3, m3=m3, r3=r3,
                                   proj=proj, z_f1=z_f1, act=act, z_f2=z_f2))
            h = h3
        out_rows = _empty((B * L, C), x)
        z_out = linear_fwd(model.proj_out, h, out_rows, lora=live_lora(model.proj_out))
        y = ops.rows_to_nchw(out_rows, B, C, Hh, Ww)
        out = ops.add_bf16(y, x)                       # + residual
        saved = dict(x=x, gmean=gmean, grstd=grstd, rows=rows, z_in=z_in, blocks=blocks, h_last=h, z_out=z_out, dims=(B, C, Hh, Ww, Lc)) \
            if save else None
        return out, saved

    @classmethod
    def backward(cls, model, sv, dout):
        B, C, Hh, Ww, Lc = sv["dims"]
        L = Hh * Ww
        dout = dout.contiguous()
        d_rows = ops.nchw_to_rows(dout)
        dh = _empty((B * L, model.inner), dout)
        linear_bwd(model.proj_out, d_rows, sv["h_last"], sv["z_out"], dh, lora=live_lora(model.proj_out))
        for blk, s in zip(reversed(list(model.transformer_blocks)), reversed(sv["blocks"])):
            ff1, ff2 = blk.ff.net[0].proj, blk.ff.net[2]
            # h3 = h2 + ff2(geglu(ff1(LN3(h2))))
            dact = _empty(s["act"].shape, dout)
            linear_bwd(ff2, dh, s["act"], s["z_f2"], dact, lora=live_lora(ff2))
            dproj = ops.geglu_bwd(dact, s["proj"])
            dn3 = _empty(dh.shape, dout)
            linear_bwd(ff1, dproj, s["n3"], s["z_f1"], dn3, lora=live_lora(ff1))
            dh2 = ops.ln_affine_bwd(dn3, s["h2"], s["m3"], s["r3"], blk.norm3.weight, dres=dh)
            # h2 = h1 + attn2(LN2(h1), ctx)
            dn2 = cls._attn_bwd(blk.attn2, s["s2"], dh2, B, L, Lc, False)
            dh1 = ops.ln_affine_bwd(dn2, s["h1"], s["m2"], s["r2"], blk.norm2.weight, dres=dh2)
            # h1 = h + attn1(LN1(h))
            dn1 = cls._attn_bwd(blk.attn1, s["s1"], dh1, B, L, L, True)
            dh = ops.ln_affine_bwd(dn1, s["h"], s["m1"], s["r1"], blk.norm1.weight, dres=dh1)
        drows = _empty(sv["rows"].shape, dout)
        linear_bwd(model.proj_in, dh, sv["rows"], sv["z_in"], drows, lora=live_lora(model.proj_in))
        dgn = ops.rows_to_nchw(drows, B, C, Hh, Ww)
        dx = ops.groupnorm_bwd(dgn, sv["x"], model.norm.weight, model.norm.bias, sv["gmean"], sv["grstd"], model.norm.num_groups)
        return ops.add_bf16(dx, dout)                  # residual path


class Transformer2DFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, context, model, anchor, save):
        out, saved = Transformer2DEngine.forward(model, x, context, save=save)
        ctx.model, ctx.saved_state = model, saved
        return out

    @staticmethod
    def backward(ctx, dout):
        dx = Transformer2DEngine.backward(ctx.model, ctx.saved_state, dout)
        ctx.saved_state = None
        return dx, None, None, None, None
