"""Frozen Wan2.1 video DiT (BASELINE.json configs[3]: Wan2.1-T2V-1.3B LoRA r = 16) with diffusers' module names, executed by
the B200 kernels.

The reference trains against `diffusers.WanTransformer3DModel` (third-party; built at toolkit/models/wan21/wan21.py:339-342,
called at :596-602 with the raw 0..1000 timestep and the UMT5 embeddings; LoRA target `['WanTransformer3DModel']`, :330,
`transformer_only` keeps the Linears under `blocks`: 10 per block = attn1.to_q/to_k/to_v/to_out.0, attn2.*, ffn.net.0.proj,
ffn.net.2).  The container below has the same class name, module tree and parameter names (checkpoints load with
`load_state_dict`; the network finds the same modules and produces the same saved keys, which
`convert_lora_weights_before_save` maps to the original `diffusion_model.blocks.N.self_attn.q...` names, wan_lora_convert.py).

`WanEngine` is the hand-scheduled forward / backward (same kernels as the FLUX engine):
  * every block Linear = one fused tcgen05 GEMM (base + LoRA up-projection segment, bias / GELU-tanh / gate / residual
    epilogues) + one rank-side GEMM; q/k/v of the self-attention and k/v of the cross-attention are fused per shared input
  * RMSNorm across heads + 3-D RoPE + head-major re-layout: `b200_rms_rope_fwd/bwd` (csrc/wan_ops.cu)
  * self-attention (L = frames x h/2 x w/2 tokens) and text cross-attention (Lk = 512): `b200_attn_fwd_x / bwd_x`
  * LayerNorm + modulation and LayerNorm-affine: `b200_ln_modulate_fwd/bwd` (the affine norm2 is the same kernel with
    scale = weight - 1, shift = bias)
  * patchify = the packed layout of `b200_flow_add_noise` ([B, 16, F*H, W] -> tokens (f, h/2, w/2) x features (c, ph, pw)) feeding
    the patch-embedding GEMM; `proj_out`'s rows are permuted ONCE at set-up from diffusers' (ph, pw, c) order to (c, ph, pw), so
    the packed loss kernel reads the prediction directly (frozen layer, no adapter: transformer_only).
Deviation from the eager model's rounding points (documented in DESIGN.md): diffusers evaluates the modulation chains in
fp32 from fp32 tables; here the six modulation vectors are rounded to bf16 once per step and the LayerNorm kernels round as
the FLUX blocks do.  Parity is asserted against the oracle at max(1e-3, 1.5 x the bf16 eager oracle's own distance to fp32).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn

from . import attention, cabi, ops
from .cabi import ACT_GELU_TANH, gemm_bf16
from .linear import group_bwd, group_fwd, linear_bwd, linear_fwd, live_lora


@dataclass
class WanConfig:
    patch_size: tuple = (1, 2, 2)
    num_attention_heads: int = 12
    attention_head_dim: int = 128
    in_channels: int = 16
    out_channels: int = 16
    text_dim: int = 4096
    freq_dim: int = 256
    ffn_dim: int = 8960
    num_layers: int = 30
    eps: float = 1e-6

    @property
    def inner_dim(self):
        return self.num_attention_heads * self.attention_head_dim


def wan_1_3b_config() -> WanConfig:
    """Wan2.1-T2V-1.3B (public model card)."""
    return WanConfig()


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container of the B200 Wan engine; call WanTransformer3DModel.forward")


class _NormW(_Holder):
    def __init__(self, dim, bias=False):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        if bias:
            self.bias = nn.Parameter(torch.zeros(dim))


class _Attention(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.to_q = nn.Linear(dim, dim)
        self.to_k = nn.Linear(dim, dim)
        self.to_v = nn.Linear(dim, dim)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
        self.norm_q = _NormW(dim)
        self.norm_k = _NormW(dim)


class _GELUProj(_Holder):
    def __init__(self, d_in, d_out):
        super().__init__()
        self.proj = nn.Linear(d_in, d_out)


class _FeedForward(_Holder):
    def __init__(self, dim, inner):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim)])


class WanTransformerBlock(_Holder):
    def __init__(self, dim, ffn_dim):
        super().__init__()
        self.attn1 = _Attention(dim)
        self.attn2 = _Attention(dim)
        self.norm2 = _NormW(dim, bias=True)
        self.ffn = _FeedForward(dim, ffn_dim)
        self.scale_shift_table = nn.Parameter(torch.zeros(1, 6, dim))


class _MLP2(_Holder):
    def __init__(self, d_in, dim):
        super().__init__()
        self.linear_1 = nn.Linear(d_in, dim)
        self.linear_2 = nn.Linear(dim, dim)


class _Condition(_Holder):
    def __init__(self, dim, freq_dim, text_dim):
        super().__init__()
        self.time_embedder = _MLP2(freq_dim, dim)
        self.time_proj = nn.Linear(dim, dim * 6)
        self.text_embedder = _MLP2(text_dim, dim)


class WanTransformer3DModel(nn.Module):
    """Same name as diffusers' class: the reference's `target_lora_modules` (wan21.py:330) selects it by name."""

    def __init__(self, cfg: WanConfig = None, device=None, dtype=torch.bfloat16):
        super().__init__()
        cfg = cfg or wan_1_3b_config()
        assert cfg.attention_head_dim == 128 and tuple(cfg.patch_size) == (1, 2, 2)
        self.cfg = cfg
        dim = cfg.inner_dim
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            with torch.device(device if device is not None else "cpu"):
                self.patch_embedding = nn.Conv3d(cfg.in_channels, dim, kernel_size=cfg.patch_size, stride=cfg.patch_size)
                self.condition_embedder = _Condition(dim, cfg.freq_dim, cfg.text_dim)
                self.blocks = nn.ModuleList([WanTransformerBlock(dim, cfg.ffn_dim) for _ in range(cfg.num_layers)])
                self.proj_out = nn.Linear(dim, cfg.out_channels * math.prod(cfg.patch_size))
                self.scale_shift_table = nn.Parameter(torch.zeros(1, 2, dim))
        finally:
            torch.set_default_dtype(prev)
        self.requires_grad_(False)
        self._engine = None

    @property
    def device(self):
        return self.proj_out.weight.device

    @property
    def dtype(self):
        return self.proj_out.weight.dtype

    def init_synthetic_(self, seed: int = 0, std: float = 0.02):
        g = torch.Generator(device=self.device).manual_seed(seed)
        with torch.no_grad():
            for name, p in self.named_parameters():
                r = torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32)
                if name.endswith("norm_q.weight") or name.endswith("norm_k.weight") or name.endswith("norm2.weight"):
                    p.copy_(1.0 + 0.1 * r)
                elif "scale_shift_table" in name:
                    p.copy_(r * 0.1)
                else:
                    p.copy_(r * std)
        self._engine = None
        return self

    @property
    def engine(self):
        if self._engine is None:
            self._engine = WanEngine(self)
        return self._engine

    def forward(self, hidden_states, timestep, encoder_hidden_states, return_dict=False, **kwargs):
        """diffusers call signature (wan21.py:596-602); returns `(sample,)` [B, 16, F, H, W].  Inference / eager use; the
        training step drives the engine directly (`WanLoRATrainStep`)."""
        B, C, Fr, H, W = hidden_states.shape
        x = hidden_states.to(torch.bfloat16).reshape(B, C, Fr * H, W).contiguous()
        zero = torch.zeros(B, device=x.device, dtype=torch.float32)
        packed = ops.flow_add_noise(x, x, zero, pack=True)  # t = 0: the packing kernel as a pure patchify
        t = timestep.to(device=x.device, dtype=torch.float32).reshape(-1).expand(B).contiguous()
        pred = self.engine.forward(packed, t, encoder_hidden_states.to(torch.bfloat16).contiguous(), (Fr, H // 2, W // 2),
                                   save=False)
        p = pred.view(B, Fr, H // 2, W // 2, C, 2, 2).permute(0, 4, 1, 2, 5, 3, 6).reshape(B, C, Fr, H, W)
        return (p,)


def _empty(shape, like, dtype=torch.bfloat16):
    return torch.empty(shape, device=like.device, dtype=dtype)


class WanEngine:
    def __init__(self, model: WanTransformer3DModel):
        self.model = model
        self.D = model.cfg.inner_dim
        self.H = model.cfg.num_attention_heads
        self._rope = {}
        self._w = None
        self.saved = None

    # ------------------------------------------------------------------------------------------ set-up (once)
    def _weights(self):
        """Frozen-weight views the kernels consume: the patch-embedding conv as a [D, 64] matrix (column order (c, 1, ph, pw)
        = the packed feature order), proj_out permuted to (c, ph, pw) output order, norm2 as (scale, shift) rows."""
        if self._w is not None:
            return self._w
        m = self.model
        C = m.cfg.out_channels
        with torch.no_grad():
            w_pe = m.patch_embedding.weight.reshape(self.D, -1).contiguous()
            idx = torch.tensor([(ph * 2 + pw) * C + c for c in range(C) for ph in range(2) for pw in range(2)],
                               device=m.proj_out.weight.device)
            w_po = m.proj_out.weight.index_select(0, idx).contiguous()
            b_po = m.proj_out.bias.index_select(0, idx).contiguous()
            n2 = [((blk.norm2.weight.float() - 1.0).to(torch.bfloat16).view(1, -1).contiguous(),
                   blk.norm2.bias.to(torch.bfloat16).view(1, -1).contiguous()) for blk in m.blocks]
        self._w = dict(w_pe=w_pe, w_po=w_po, b_po=b_po, n2=n2)
        return self._w

    def _tables(self, B):
        """The frozen `scale_shift_table`s broadcast over the batch ([B, 6 D] per block, [B, D] x 2 for the head), built once per
        batch size so that the per-step modulation is ONE add kernel per block."""
        hit = getattr(self, "_tabs", {}).get(B)
        if hit is not None:
            return hit
        m, D = self.model, self.D
        with torch.no_grad():
            blocks = [blk.scale_shift_table.to(torch.bfloat16).reshape(1, 6 * D).expand(B, 6 * D).contiguous() for blk in m.blocks]
            t2 = m.scale_shift_table.to(torch.bfloat16).reshape(2, D)
            head = (t2[0:1].expand(B, D).contiguous(), t2[1:2].expand(B, D).contiguous())
        if not hasattr(self, "_tabs"):
            self._tabs = {}
        self._tabs[B] = dict(blocks=blocks, head=head)
        return self._tabs[B]

    def rope_tables(self, grid, device):
        """WanRotaryPosEmbed as fp32 cos / sin tables [L, 128] (pairs repeated): axes (t, h, w) with dims
        (128 - 4 (128 // 6), 2 (128 // 6), 2 (128 // 6)) = (44, 42, 42), angles in float64."""
        key = (tuple(grid), str(device))
        hit = self._rope.get(key)
        if hit is not None:
            return hit
        ppf, pph, ppw = grid
        d = 128
        h_dim = w_dim = 2 * (d // 6)
        t_dim = d - h_dim - w_dim
        parts = []
        for dim, n, shape in ((t_dim, ppf, (ppf, 1, 1)), (h_dim, pph, (1, pph, 1)), (w_dim, ppw, (1, 1, ppw))):
            freqs = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float64, device=device)[: dim // 2] / dim))
            ang = torch.outer(torch.arange(n, dtype=torch.float64, device=device), freqs)
            parts.append(ang.view(*shape, -1).expand(ppf, pph, ppw, -1))
        ang = torch.cat(parts, dim=-1).reshape(ppf * pph * ppw, d // 2)
        out = (ang.cos().repeat_interleave(2, dim=1).float().contiguous(), ang.sin().repeat_interleave(2, dim=1).float().contiguous())
        self._rope[key] = out
        return out

    def active_network(self):
        ref = getattr(self.model, "_b200_network", None)
        return ref() if ref is not None else None

    def _register_groups(self, net):
        groups = []
        for blk in self.model.blocks:
            groups.append([live_lora(l) for l in (blk.attn1.to_q, blk.attn1.to_k, blk.attn1.to_v)])
            groups.append([live_lora(l) for l in (blk.attn2.to_k, blk.attn2.to_v)])
        net.register_fused_groups(groups)
        self._groups_for = net

    # ------------------------------------------------------------------------------------------ helpers
    @staticmethod
    def _group_or_each_fwd(lins, x, out):
        loras = [live_lora(l) for l in lins]
        grp = loras[0].network_ref().fused_group(loras) if all(l is not None for l in loras) else None
        if grp is not None:
            return ("g", group_fwd(grp, lins, x, out))
        D = lins[0].out_features
        return ("s", [linear_fwd(lin, x, out[:, j * D:(j + 1) * D], lora=lo) for j, (lin, lo) in enumerate(zip(lins, loras))])

    @staticmethod
    def _group_or_each_bwd(lins, dy, x, saved, dx, accumulate=False):
        loras = [live_lora(l) for l in lins]
        kind, z = saved
        if kind == "g":
            group_bwd(loras[0].network_ref().fused_group(loras), lins, dy, x, z, dx, **({"res": dx} if accumulate else {}))
            return
        D = lins[0].out_features
        for j, (lin, lo) in enumerate(zip(lins, loras)):
            acc = accumulate or j > 0
            linear_bwd(lin, dy[:, j * D:(j + 1) * D], x, z[j], dx, lora=lo, **({"res": dx} if (acc and dx is not None) else {}))

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, packed, timesteps, enc, grid, save=True):
        """packed [B, L, 64] bf16 (patchified noisy latents), timesteps [B] fp32 (0..1000), enc [B, Lt, text_dim] bf16,
        grid = (frames, h/2, w/2).  Returns pred [B*L, 64] bf16 in the packed (c, ph, pw) feature order."""
        m = self.model
        D, H = self.D, self.H
        B, L, Cin = packed.shape
        Lt = enc.shape[1]
        dev = packed.device
        if dev.type != "cuda":
            raise cabi.B200Error("WanEngine needs a B200; there is no CPU / eager fallback")
        assert L == grid[0] * grid[1] * grid[2]
        net = self.active_network()
        if net is not None and net.is_active and not net.is_merged_in and len(net.get_all_modules()):
            if getattr(self, "_groups_for", None) is not net:
                self._register_groups(net)
            if any(mod.has_dropout() or (mod.module_dropout and mod.training) for mod in net.get_all_modules()):
                raise NotImplementedError("dropout variants with the fused Wan engine: use the per-module path")
            net.refresh_packs()
            net.ensure_grad_views()
        W = self._weights()
        cos, sin = self.rope_tables(grid, dev)
        ce = m.condition_embedder
        # --- conditioning (frozen prelude, M = B rows: weight-streaming GEMV kernels)
        t_sin = ops.timestep_embed(timesteps, m.cfg.freq_dim, div=1.0, mult=0.0)  # fp32 timestep, no bf16 re-scaling
        h1, _ = ops.lora_gemv_fwd(t_sin, ce.time_embedder.linear_1.weight, ce.time_embedder.linear_1.bias)
        temb, _ = ops.lora_gemv_fwd(ops.silu(h1), ce.time_embedder.linear_2.weight, ce.time_embedder.linear_2.bias)
        proj, _ = ops.lora_gemv_fwd(ops.silu(temb), ce.time_proj.weight, ce.time_proj.bias)  # [B, 6 D]
        # --- text projection (frozen): GELU-tanh MLP over the B * Lt text tokens
        e1 = _empty((B * Lt, D), packed)
        gemm_bf16(enc.reshape(B * Lt, enc.shape[2]), ce.text_embedder.linear_1.weight, e1, bias=ce.text_embedder.linear_1.bias,
                  act=ACT_GELU_TANH)
        text = _empty((B * Lt, D), packed)
        gemm_bf16(e1, ce.text_embedder.linear_2.weight, text, bias=ce.text_embedder.linear_2.bias)
        # --- patch embedding (frozen): Conv3d (1, 2, 2) stride (1, 2, 2) == Linear over the packed patch features
        x = _empty((B * L, D), packed)
        gemm_bf16(packed.reshape(B * L, Cin), W["w_pe"], x, bias=m.patch_embedding.bias)
        sv = {"B": B, "L": L, "Lt": Lt, "cos": cos, "sin": sin, "blocks": [], "text": text} if save else None
        tabs = self._tables(B)
        for bi, blk in enumerate(m.blocks):
            # six modulation vectors of this block: bf16(scale_shift_table + time projection), [B, 6 D]
            mod = ops.add_bf16(tabs["blocks"][bi], proj)
            x, s = self._block_fwd(blk, W["n2"][bi], x, text, mod, B, L, Lt, cos, sin, save)
            if save:
                sv["blocks"].append(s)
        # --- output head (frozen): LN modulated by (table + temb), Linear to 64 in (c, ph, pw) order
        shift_o = ops.add_bf16(tabs["head"][0], temb)
        scale_o = ops.add_bf16(tabs["head"][1], temb)
        n_out, mean_o, rstd_o = ops.ln_modulate_fwd(x, shift_o, scale_o, L)
        pred = _empty((B * L, Cin), packed)
        gemm_bf16(n_out, W["w_po"], pred, bias=W["b_po"])
        if save:
            sv.update(x_final=x, scale_o=scale_o, mean_o=mean_o, rstd_o=rstd_o)
            self.saved = sv
        return pred

    def _block_fwd(self, blk, n2w, x, text, mod, B, L, Lt, cos, sin, save):
        D, H = self.D, self.H
        a1, a2 = blk.attn1, blk.attn2
        # ---- self-attention
        n1, mean1, rstd1 = ops.ln_modulate_fwd(x, mod[:, 0:D], mod[:, D:2 * D], L)
        qkv = _empty((B * L, 3 * D), x)
        zq = self._group_or_each_fwd((a1.to_q, a1.to_k, a1.to_v), n1, qkv)
        Q, K, V = (_empty((B, H, L, 128), x) for _ in range(3))
        rq = ops.rms_rope_fwd(qkv[:, :D], a1.norm_q.weight, cos, sin, Q, B, L)
        rk = ops.rms_rope_fwd(qkv[:, D:2 * D], a1.norm_k.weight, cos, sin, K, B, L)
        ops.rms_rope_fwd(qkv[:, 2 * D:], None, None, None, V, B, L, mode=0)
        o = _empty((B * L, D), x)
        lse = attention.fwd(Q, K, V, None, o, 0)
        x1 = _empty(x.shape, x)
        z_o = linear_fwd(a1.to_out[0], o, x1, lora=live_lora(a1.to_out[0]), gate=mod[:, 2 * D:3 * D], rows_per_sample=L, res=x)
        # ---- cross-attention on the projected text tokens
        n2, mean2, rstd2 = ops.ln_modulate_fwd(x1, n2w[1], n2w[0], B * L)  # LayerNorm affine: scale = w - 1, shift = b
        q2 = _empty((B * L, D), x)
        z_q2 = linear_fwd(a2.to_q, n2, q2, lora=live_lora(a2.to_q))
        kv2 = _empty((B * Lt, 2 * D), x)
        z_kv2 = self._group_or_each_fwd((a2.to_k, a2.to_v), text, kv2)
        Q2 = _empty((B, H, L, 128), x)
        K2, V2 = (_empty((B, H, Lt, 128), x) for _ in range(2))
        rq2 = ops.rms_rope_fwd(q2, a2.norm_q.weight, None, None, Q2, B, L)
        rk2 = ops.rms_rope_fwd(kv2[:, :D], a2.norm_k.weight, None, None, K2, B, Lt)
        ops.rms_rope_fwd(kv2[:, D:], None, None, None, V2, B, Lt, mode=0)
        o2 = _empty((B * L, D), x)
        lse2 = attention.fwd(Q2, K2, V2, None, o2, 0)
        x2 = _empty(x.shape, x)
        z_o2 = linear_fwd(a2.to_out[0], o2, x2, lora=live_lora(a2.to_out[0]), res=x1)
        # ---- MLP
        n3, mean3, rstd3 = ops.ln_modulate_fwd(x2, mod[:, 3 * D:4 * D], mod[:, 4 * D:5 * D], L)
        inner = blk.ffn.net[0].proj.out_features
        pre = _empty((B * L, inner), x)
        act = _empty((B * L, inner), x)
        z_f1 = linear_fwd(blk.ffn.net[0].proj, n3, act, lora=live_lora(blk.ffn.net[0].proj), act=ACT_GELU_TANH, aux_out=pre)
        x3 = _empty(x.shape, x)
        z_f2 = linear_fwd(blk.ffn.net[2], act, x3, lora=live_lora(blk.ffn.net[2]), gate=mod[:, 5 * D:6 * D], rows_per_sample=L,
                          res=x2)
        s = None
        if save:
            s = dict(x=x, mod=mod, n1=n1, mean1=mean1, rstd1=rstd1, qkv=qkv, zq=zq, Q=Q, K=K, V=V, rq=rq, rk=rk, o=o, lse=lse,
                     z_o=z_o, x1=x1, n2=n2, mean2=mean2, rstd2=rstd2, q2=q2, z_q2=z_q2, kv2=kv2, z_kv2=z_kv2, Q2=Q2, K2=K2, V2=V2,
                     rq2=rq2, rk2=rk2, o2=o2, lse2=lse2, z_o2=z_o2, x2=x2, n3=n3, mean3=mean3, rstd3=rstd3, pre=pre, act=act,
                     z_f1=z_f1, z_f2=z_f2)
        return x3, s

    # ------------------------------------------------------------------------------------------ backward
    def backward(self, dpred):
        """dpred [B*L, 64] bf16 (packed order) -> accumulates dA / dB of every live adapter into the flat gradient buffer.
        Nothing upstream of the blocks is trainable (embedders, modulation tables and the text projection are frozen and
        outside `blocks`), so no gradient is formed for the text tokens' producers or the modulation vectors."""
        sv = self.saved
        assert sv is not None, "WanEngine.backward without a saved forward"
        m = self.model
        D = self.D
        B, L, Lt = sv["B"], sv["L"], sv["Lt"]
        W = self._weights()
        dn = _empty((B * L, D), dpred)
        gemm_bf16(dpred, W["w_po"], dn, trans_b=True)
        dx = ops.ln_modulate_bwd(dn, sv["x_final"], sv["mean_o"], sv["rstd_o"], sv["scale_o"], L)
        nb = len(m.blocks)
        for bi in range(nb - 1, -1, -1):
            dx = self._block_bwd(m.blocks[bi], W["n2"][bi], sv["blocks"][bi], dx, sv["text"], B, L, Lt, sv["cos"], sv["sin"],
                                 need_dx=bi > 0)
        self.saved = None

    def _block_bwd(self, blk, n2w, s, dx3, text, B, L, Lt, cos, sin, need_dx=True):
        D = self.D
        a1, a2 = blk.attn1, blk.attn2
        mod = s["mod"]
        # ---- MLP: x3 = x2 + c_gate * ffn2(gelu(ffn1(n3)))
        dy = _empty(dx3.shape, dx3)
        ops.col_reduce(dx3, L, g=mod[:, 5 * D:6 * D], mul_out=dy)
        dpre = _empty(s["pre"].shape, dx3)
        linear_bwd(blk.ffn.net[2], dy, s["act"], s["z_f2"], dpre, lora=live_lora(blk.ffn.net[2]), aux_in=s["pre"])
        dn3 = _empty(dx3.shape, dx3)
        linear_bwd(blk.ffn.net[0].proj, dpre, s["n3"], s["z_f1"], dn3, lora=live_lora(blk.ffn.net[0].proj))
        dx2 = ops.ln_modulate_bwd(dn3, s["x2"], s["mean3"], s["rstd3"], mod[:, 4 * D:5 * D], L, dres=dx3)
        # ---- cross-attention: x2 = x1 + to_out(attn(q(n2), k(text), v(text)))
        do2 = _empty(dx2.shape, dx2)
        linear_bwd(a2.to_out[0], dx2, s["o2"], s["z_o2"], do2, lora=live_lora(a2.to_out[0]))
        dQ2, dK2, dV2 = attention.bwd(s["Q2"], s["K2"], s["V2"], None, s["o2"], None, do2, s["lse2"], 0)
        dq2 = _empty(s["q2"].shape, dx2)
        ops.rms_rope_bwd(dQ2, s["q2"], a2.norm_q.weight, None, None, s["rq2"], dq2, B, L)
        dkv2 = _empty(s["kv2"].shape, dx2)
        ops.rms_rope_bwd(dK2, s["kv2"][:, :D], a2.norm_k.weight, None, None, s["rk2"], dkv2[:, :D], B, Lt)
        ops.rms_rope_bwd(dV2, None, None, None, None, None, dkv2[:, D:], B, Lt, mode=0)
        self._group_or_each_bwd((a2.to_k, a2.to_v), dkv2, text, s["z_kv2"], None)  # adapters only: the text path is frozen
        dn2 = _empty(dx2.shape, dx2)
        linear_bwd(a2.to_q, dq2, s["n2"], s["z_q2"], dn2, lora=live_lora(a2.to_q))
        dx1 = ops.ln_modulate_bwd(dn2, s["x1"], s["mean2"], s["rstd2"], n2w[0], B * L, dres=dx2)
        # ---- self-attention: x1 = x + gate * to_out(attn(...))
        dya = dy  # reuse
        ops.col_reduce(dx1, L, g=mod[:, 2 * D:3 * D], mul_out=dya)
        do = _empty(dx1.shape, dx1)
        linear_bwd(a1.to_out[0], dya, s["o"], s["z_o"], do, lora=live_lora(a1.to_out[0]))
        dQ, dK, dV = attention.bwd(s["Q"], s["K"], s["V"], None, s["o"], None, do, s["lse"], 0)
        dqkv = _empty(s["qkv"].shape, dx1)
        qkv = s["qkv"]
        ops.rms_rope_bwd(dQ, qkv[:, :D], a1.norm_q.weight, cos, sin, s["rq"], dqkv[:, :D], B, L)
        ops.rms_rope_bwd(dK, qkv[:, D:2 * D], a1.norm_k.weight, cos, sin, s["rk"], dqkv[:, D:2 * D], B, L)
        ops.rms_rope_bwd(dV, None, None, None, None, None, dqkv[:, 2 * D:], B, L, mode=0)
        dn1 = _empty(dx1.shape, dx1) if need_dx else None
        self._group_or_each_bwd((a1.to_q, a1.to_k, a1.to_v), dqkv, s["n1"], s["zq"], dn1)
        if not need_dx:
            return None
        return ops.ln_modulate_bwd(dn1, s["x"], s["mean1"], s["rstd1"], mod[:, D:2 * D], L, dres=dx1)
