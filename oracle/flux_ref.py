"""TEST INFRASTRUCTURE ONLY (CPU/eager restatement; never imported by ai_toolkit_b200/).

Plain-PyTorch restatement of the frozen FLUX.1 DiT forward that ai-toolkit trains LoRAs against.
The arithmetic itself lives in the third-party dependency `diffusers` pinned at git commit
c943837899b16cbae2f619b8dd4f7bb6f07dd81a (reference requirements.txt:5), which is NOT vendored under
/root/reference and not installed here, so this file restates the published algorithm with the
diffusers module names (so LoRA names / state-dict keys match what the reference saves) and anchors on
the in-tree call sites and the in-tree BFL-lineage restatement:

  * call site / input packing ...... toolkit/stable_diffusion_model.py:2154-2222
  * double / single stream block ... extensions_built_in/diffusion_models/chroma/src/layers.py:471-681
  * RoPE + attention ............... extensions_built_in/diffusion_models/chroma/src/math.py:13-51
  * sinusoidal timestep embedding .. extensions_built_in/diffusion_models/chroma/src/layers.py:30-53
  * QK RMSNorm ..................... extensions_built_in/diffusion_models/chroma/src/layers.py:72-91,417-427
  * FLUX.1 dimensions .............. extensions_built_in/diffusion_models/chroma/src/model.py:37-53

PARITY UNPINNED at the diffusers boundary: the reference holds no golden vectors for the DiT forward
 and diffusers itself cannot be run here.  What IS checked (tests/test_oracle_pinned.py,
tests/test_oracle_crosscheck_bfl.py):
  * the double / single stream blocks and the timestep embedding against the reference's OWN in-tree blocks listed
    above, executed from /root/reference with mapped weights (forward 1e-5, backward 1e-4, fp32);
  * the whole model (embedders, modulation order, RoPE, final AdaLN with its scale/shift swap) against an independent
    BFL-lineage implementation installed with torchtitan, through the public BFL -> diffusers weight conversion;
  * the LoRA wrapper (which IS in-tree) by running the reference's own classes on top of this model
    (oracle/make_golden.py).
What stays unpinned is only that diffusers @ c9438378 implements this same published architecture.

Every nn.Linear here is a plain `torch.nn.Linear` so that the reference's unmodified
`LoRASpecialNetwork` (class-name matching, toolkit/lora_special.py:484-489) attaches to it.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class FluxConfig:
    in_channels: int = 64
    num_layers: int = 19
    num_single_layers: int = 38
    attention_head_dim: int = 128
    num_attention_heads: int = 24
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    guidance_embeds: bool = True
    axes_dims_rope: tuple = (16, 56, 56)
    mlp_ratio: float = 4.0

    @property
    def inner_dim(self):
        return self.attention_head_dim * self.num_attention_heads


def flux_dev_config() -> FluxConfig:
    return FluxConfig()


def tiny_config(layers=1, single_layers=1, heads=2, joint_dim=64, pooled=32) -> FluxConfig:
    return FluxConfig(num_layers=layers, num_single_layers=single_layers, num_attention_heads=heads,
                      joint_attention_dim=joint_dim, pooled_projection_dim=pooled)


def get_timestep_embedding(timesteps: torch.Tensor, dim: int = 256, max_period: int = 10000) -> torch.Tensor:
    """diffusers `Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0)` -> [cos | sin]."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def rope_cos_sin(ids: torch.Tensor, axes_dim=(16, 56, 56), theta: float = 10000.0):
    """FluxPosEmbed: per-axis rotary tables (float64 angles), cos/sin repeat-interleaved to head_dim."""
    cos_out, sin_out = [], []
    pos = ids.to(torch.float64)
    for i, d in enumerate(axes_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64, device=ids.device)[: d // 2] / d))
        ang = torch.outer(pos[:, i], freqs)
        cos_out.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin_out.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos_out, dim=-1), torch.cat(sin_out, dim=-1)


def apply_rotary_emb(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x [B,H,S,D]; (x_real, x_imag) are interleaved pairs along D."""
    cos = cos[None, None]
    sin = sin[None, None]
    x_real, x_imag = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rot = torch.stack([-x_imag, x_real], dim=-1).flatten(3)
    return (x.float() * cos + x_rot.float() * sin).to(x.dtype)


class RMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        in_dtype = x.dtype
        var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(var + self.eps)
        if self.weight.dtype in (torch.float16, torch.bfloat16):
            x = x.to(self.weight.dtype)
        x = x * self.weight
        return x if self.weight.dtype in (torch.float16, torch.bfloat16) else x.to(in_dtype)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim, dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class CombinedTimestepGuidanceTextProjEmbeddings(nn.Module):
    def __init__(self, dim, pooled_dim, guidance_embeds=True):
        super().__init__()
        self.timestep_embedder = TimestepEmbedding(256, dim)
        if guidance_embeds:
            self.guidance_embedder = TimestepEmbedding(256, dim)
        self.text_embedder = TimestepEmbedding(pooled_dim, dim)
        self.guidance_embeds = guidance_embeds

    def forward(self, timestep, guidance, pooled):
        t_emb = self.timestep_embedder(get_timestep_embedding(timestep).to(pooled.dtype))
        if self.guidance_embeds:
            g_emb = self.guidance_embedder(get_timestep_embedding(guidance).to(pooled.dtype))
            t_emb = t_emb + g_emb
        return t_emb + self.text_embedder(pooled)


class AdaLayerNormZero(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.linear = nn.Linear(dim, 6 * dim)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, emb):
        emb = self.linear(F.silu(emb))
        shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = emb.chunk(6, dim=1)
        x = self.norm(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
        return x, gate_msa, shift_mlp, scale_mlp, gate_mlp


class AdaLayerNormZeroSingle(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.linear = nn.Linear(dim, 3 * dim)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, emb):
        emb = self.linear(F.silu(emb))
        shift_msa, scale_msa, gate_msa = emb.chunk(3, dim=1)
        x = self.norm(x) * (1 + scale_msa[:, None]) + shift_msa[:, None]
        return x, gate_msa


class AdaLayerNormContinuous(nn.Module):
    def __init__(self, dim, cond_dim):
        super().__init__()
        self.linear = nn.Linear(cond_dim, 2 * dim)
        self.norm = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)

    def forward(self, x, cond):
        emb = self.linear(F.silu(cond).to(x.dtype))
        scale, shift = torch.chunk(emb, 2, dim=1)
        return self.norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]


class GELU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4.0):
        super().__init__()
        inner = int(dim * mult)
        self.net = nn.ModuleList([GELU(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class Attention(nn.Module):
    """Joint attention; `pre_only=True` (single blocks) has no output projection."""

    def __init__(self, dim, heads, head_dim, added_kv=False, pre_only=False):
        super().__init__()
        self.heads, self.head_dim = heads, head_dim
        self.to_q = nn.Linear(dim, dim)
        self.to_k = nn.Linear(dim, dim)
        self.to_v = nn.Linear(dim, dim)
        self.norm_q = RMSNorm(head_dim)
        self.norm_k = RMSNorm(head_dim)
        self.added_kv = added_kv
        if added_kv:
            self.add_q_proj = nn.Linear(dim, dim)
            self.add_k_proj = nn.Linear(dim, dim)
            self.add_v_proj = nn.Linear(dim, dim)
            self.norm_added_q = RMSNorm(head_dim)
            self.norm_added_k = RMSNorm(head_dim)
            self.to_add_out = nn.Linear(dim, dim)
        if not pre_only:
            self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
        self.pre_only = pre_only

    def _heads(self, x):
        b, s, _ = x.shape
        return x.view(b, s, self.heads, self.head_dim).transpose(1, 2)

    def forward(self, hidden, encoder_hidden=None, rope=None):
        q = self.norm_q(self._heads(self.to_q(hidden)))
        k = self.norm_k(self._heads(self.to_k(hidden)))
        v = self._heads(self.to_v(hidden))
        if encoder_hidden is not None:
            eq = self.norm_added_q(self._heads(self.add_q_proj(encoder_hidden)))
            ek = self.norm_added_k(self._heads(self.add_k_proj(encoder_hidden)))
            ev = self._heads(self.add_v_proj(encoder_hidden))
            q = torch.cat([eq, q], dim=2)
            k = torch.cat([ek, k], dim=2)
            v = torch.cat([ev, v], dim=2)
        if rope is not None:
            q = apply_rotary_emb(q, *rope)
            k = apply_rotary_emb(k, *rope)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=0.0, is_causal=False)
        b, h, s, d = o.shape
        o = o.transpose(1, 2).reshape(b, s, h * d).to(q.dtype)
        if encoder_hidden is not None:
            t = encoder_hidden.shape[1]
            enc_o, o = o[:, :t], o[:, t:]
            o = self.to_out[0](o)
            enc_o = self.to_add_out(enc_o)
            return o, enc_o
        return o


class FluxTransformerBlock(nn.Module):
    def __init__(self, dim, heads, head_dim, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = AdaLayerNormZero(dim)
        self.norm1_context = AdaLayerNormZero(dim)
        self.attn = Attention(dim, heads, head_dim, added_kv=True)
        self.norm2 = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff = FeedForward(dim, mlp_ratio)
        self.norm2_context = nn.LayerNorm(dim, elementwise_affine=False, eps=1e-6)
        self.ff_context = FeedForward(dim, mlp_ratio)

    def forward(self, hidden, encoder_hidden, temb, rope):
        n_h, gate_msa, shift_mlp, scale_mlp, gate_mlp = self.norm1(hidden, temb)
        n_e, c_gate_msa, c_shift_mlp, c_scale_mlp, c_gate_mlp = self.norm1_context(encoder_hidden, temb)
        attn_o, ctx_o = self.attn(n_h, n_e, rope)
        hidden = hidden + gate_msa.unsqueeze(1) * attn_o
        n_h = self.norm2(hidden)
        n_h = n_h * (1 + scale_mlp[:, None]) + shift_mlp[:, None]
        hidden = hidden + gate_mlp.unsqueeze(1) * self.ff(n_h)
        encoder_hidden = encoder_hidden + c_gate_msa.unsqueeze(1) * ctx_o
        n_e = self.norm2_context(encoder_hidden)
        n_e = n_e * (1 + c_scale_mlp[:, None]) + c_shift_mlp[:, None]
        encoder_hidden = encoder_hidden + c_gate_mlp.unsqueeze(1) * self.ff_context(n_e)
        return encoder_hidden, hidden


class FluxSingleTransformerBlock(nn.Module):
    def __init__(self, dim, heads, head_dim, mlp_ratio=4.0):
        super().__init__()
        self.mlp_hidden_dim = int(dim * mlp_ratio)
        self.norm = AdaLayerNormZeroSingle(dim)
        self.proj_mlp = nn.Linear(dim, self.mlp_hidden_dim)
        self.proj_out = nn.Linear(dim + self.mlp_hidden_dim, dim)
        self.attn = Attention(dim, heads, head_dim, pre_only=True)

    def forward(self, hidden, temb, rope):
        residual = hidden
        n_h, gate = self.norm(hidden, temb)
        mlp_h = F.gelu(self.proj_mlp(n_h), approximate="tanh")
        attn_o = self.attn(n_h, None, rope)
        hidden = torch.cat([attn_o, mlp_h], dim=2)
        hidden = gate.unsqueeze(1) * self.proj_out(hidden)
        return residual + hidden


class FluxTransformer2DModel(nn.Module):
    """Same class name as diffusers': the reference targets LoRA by this name (lora_special.py:692-693)."""

    def __init__(self, cfg: FluxConfig):
        super().__init__()
        self.cfg = cfg
        d = cfg.inner_dim
        self.time_text_embed = CombinedTimestepGuidanceTextProjEmbeddings(d, cfg.pooled_projection_dim,
                                                                          cfg.guidance_embeds)
        self.context_embedder = nn.Linear(cfg.joint_attention_dim, d)
        self.x_embedder = nn.Linear(cfg.in_channels, d)
        self.transformer_blocks = nn.ModuleList(
            [FluxTransformerBlock(d, cfg.num_attention_heads, cfg.attention_head_dim, cfg.mlp_ratio)
             for _ in range(cfg.num_layers)])
        self.single_transformer_blocks = nn.ModuleList(
            [FluxSingleTransformerBlock(d, cfg.num_attention_heads, cfg.attention_head_dim, cfg.mlp_ratio)
             for _ in range(cfg.num_single_layers)])
        self.norm_out = AdaLayerNormContinuous(d, d)
        self.proj_out = nn.Linear(d, cfg.in_channels)

    def forward(self, hidden_states, timestep, encoder_hidden_states, pooled_projections, txt_ids, img_ids,
                guidance=None):
        hidden = self.x_embedder(hidden_states)
        timestep = timestep.to(hidden.dtype) * 1000
        if guidance is not None:
            guidance = guidance.to(hidden.dtype) * 1000
        temb = self.time_text_embed(timestep, guidance, pooled_projections)
        enc = self.context_embedder(encoder_hidden_states)
        ids = torch.cat((txt_ids, img_ids), dim=0)
        rope = rope_cos_sin(ids, self.cfg.axes_dims_rope)
        ckpt = getattr(self, "gradient_checkpointing", False) and torch.is_grad_enabled()
        if ckpt:
            from torch.utils.checkpoint import checkpoint
        for blk in self.transformer_blocks:
            if ckpt:  # the reference's default (config_modules.py:413): re-run the block's forward in backward
                enc, hidden = checkpoint(blk, hidden, enc, temb, rope, use_reentrant=False)
            else:
                enc, hidden = blk(hidden, enc, temb, rope)
        hidden = torch.cat([enc, hidden], dim=1)
        for blk in self.single_transformer_blocks:
            hidden = checkpoint(blk, hidden, temb, rope, use_reentrant=False) if ckpt else blk(hidden, temb, rope)
        hidden = hidden[:, enc.shape[1]:, ...]
        hidden = self.norm_out(hidden, temb)
        return self.proj_out(hidden)


# ---------------------------------------------------------------------------------------------
# input packing — toolkit/stable_diffusion_model.py:2157-2219
# ---------------------------------------------------------------------------------------------
def pack_latents(latents: torch.Tensor) -> torch.Tensor:
    """`rearrange(x, "b c (h ph) (w pw) -> b (h w) (c ph pw)", ph=2, pw=2)` (:2166-2172)."""
    b, c, h, w = latents.shape
    x = latents.view(b, c, h // 2, 2, w // 2, 2)
    return x.permute(0, 2, 4, 1, 3, 5).reshape(b, (h // 2) * (w // 2), c * 4)


def unpack_latents(x: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """inverse rearrange (:2211-2219); h, w are LATENT height/width."""
    b, _, cpp = x.shape
    c = cpp // 4
    x = x.view(b, h // 2, w // 2, c, 2, 2)
    return x.permute(0, 3, 1, 4, 2, 5).reshape(b, c, h, w)


def make_img_ids(h: int, w: int, device=None) -> torch.Tensor:
    """img_ids[..., 1] = row, [..., 2] = col over the packed (h/2, w/2) grid (:2174-2178)."""
    ids = torch.zeros(h // 2, w // 2, 3, device=device)
    ids[..., 1] = ids[..., 1] + torch.arange(h // 2, device=device)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(w // 2, device=device)[None, :]
    return ids.reshape(-1, 3)


def init_synthetic_(model: nn.Module, seed: int = 0, std: float = 0.02):
    """Synthetic frozen weights N(0, std^2); biases small; RMSNorm weights near 1."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 2:
                p.copy_(torch.randn(p.shape, generator=g) * std)
            elif name.endswith("norm_q.weight") or name.endswith("norm_k.weight") or "norm_added" in name:
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return model
