"""TEST INFRASTRUCTURE ONLY -- eager PyTorch restatement of the frozen Wan2.1 video DiT (BASELINE.json configs[3]) with
diffusers' module names, so that `LoRASpecialNetwork(target_lin_modules=['WanTransformer3DModel'], transformer_only=True)`
produces the reference's adapter names (`transformer.blocks.N.attn1.to_q`, ...; toolkit/models/wan21/wan21.py:330).

PARITY UNPINNED at the diffusers boundary: `WanTransformer3DModel` is third-party diffusers @ c943837899b16cbae2f619b8dd4f7bb6f07dd81a
(reference requirements.txt:5), absent here; the reference holds no golden vectors for it.  What IS in-tree and pinned:
  * the attention arithmetic -- `WanAttnProcessor2_0` (toolkit/models/wan21/wan_attn.py:8-84): q/k/v projections, RMSNorm
    across heads BEFORE the head split, RoPE as a complex multiply in float64, torch SDPA, output projection.
    tests/test_wan.py runs that unmodified class on this file's `WanAttention` modules and compares.
  * the call site (`Wan21.get_noise_prediction`, wan21.py:578-603: raw 0..1000 timestep, `encoder_hidden_states` = UMT5
    embeddings) and the loss target (`get_loss_target`, :717-724: noise - latents).
The block / embedder structure restates the published model: per block
    shift, scale, gate, c_shift, c_scale, c_gate = (scale_shift_table + temb6).chunk(6)          [fp32]
    x = x + gate * attn1( LN(x) (1 + scale) + shift , rope)                                       self-attention
    x = x + attn2( LN_affine(x), text )                                                           cross-attention
    x = x + c_gate * ffn( LN(x) (1 + c_scale) + c_shift )                                         GELU-tanh MLP
with FP32 LayerNorms (eps 1e-6), patchify Conv3d(1,2,2), sinusoidal(256) -> MLP time embedding, SiLU -> Linear(6 dim) time
projection, GELU-tanh text projection, final LN modulated by (scale_shift_table + temb), Linear to 64 = (1,2,2,16) and
unpatchify.  Dimensions of Wan2.1-T2V-1.3B (public model card): dim 1536, 12 heads x 128, ffn 8960, 30 layers, text 4096.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


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
    rope_max_seq_len: int = 1024

    @property
    def inner_dim(self):
        return self.num_attention_heads * self.attention_head_dim


def wan_1_3b_config() -> WanConfig:
    return WanConfig()


class FP32LayerNorm(nn.LayerNorm):
    def forward(self, x):
        return F.layer_norm(x.float(), self.normalized_shape, self.weight.float() if self.weight is not None else None,
                            self.bias.float() if self.bias is not None else None, self.eps).to(x.dtype)


class RMSNorm(nn.Module):
    """diffusers RMSNorm: fp32 variance, the normalised value cast to the (half) weight dtype, then times weight."""

    def __init__(self, dim, eps):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        dt = x.dtype
        var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(var + self.eps)
        if self.weight.dtype in (torch.float16, torch.bfloat16):
            x = x.to(self.weight.dtype)
        x = x * self.weight
        return x.to(dt) if self.weight.dtype == torch.float32 else x


def apply_rotary_emb(x, freqs):
    """wan_attn.py:48-53: x [B, H, L, 128] as complex pairs (float64) times freqs [1, 1, L, 64] complex."""
    xr = torch.view_as_complex(x.to(torch.float64).unflatten(3, (-1, 2)))
    return torch.view_as_real(xr * freqs).flatten(3, 4).type_as(x)


class WanAttention(nn.Module):
    """diffusers `Attention` as Wan builds it (bias everywhere, qk_norm across heads) + WanAttnProcessor2_0's forward."""

    def __init__(self, dim, heads, eps, cross=False):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim)
        self.to_k = nn.Linear(dim, dim)
        self.to_v = nn.Linear(dim, dim)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
        self.norm_q = RMSNorm(dim, eps)
        self.norm_k = RMSNorm(dim, eps)
        self.add_k_proj = None  # T2V: no image branch (wan_attn.py:25-29)

    def forward(self, hidden_states, encoder_hidden_states=None, rotary_emb=None):
        enc = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q = self.norm_q(self.to_q(hidden_states))
        k = self.norm_k(self.to_k(enc))
        v = self.to_v(enc)
        q = q.unflatten(2, (self.heads, -1)).transpose(1, 2)
        k = k.unflatten(2, (self.heads, -1)).transpose(1, 2)
        v = v.unflatten(2, (self.heads, -1)).transpose(1, 2)
        if rotary_emb is not None:
            q = apply_rotary_emb(q, rotary_emb)
            k = apply_rotary_emb(k, rotary_emb)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).flatten(2, 3).type_as(q)
        return self.to_out[1](self.to_out[0](o))


class _GELUProj(nn.Module):
    def __init__(self, d_in, d_out):
        super().__init__()
        self.proj = nn.Linear(d_in, d_out)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class FeedForward(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class WanTransformerBlock(nn.Module):
    def __init__(self, dim, ffn_dim, heads, eps):
        super().__init__()
        self.norm1 = FP32LayerNorm(dim, eps, elementwise_affine=False)
        self.attn1 = WanAttention(dim, heads, eps)
        self.attn2 = WanAttention(dim, heads, eps, cross=True)
        self.norm2 = FP32LayerNorm(dim, eps, elementwise_affine=True)  # cross_attn_norm=True
        self.ffn = FeedForward(dim, ffn_dim)
        self.norm3 = FP32LayerNorm(dim, eps, elementwise_affine=False)
        self.scale_shift_table = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)

    def forward(self, x, enc, temb6, rotary_emb):
        shift, scale, gate, c_shift, c_scale, c_gate = (self.scale_shift_table + temb6.float()).chunk(6, dim=1)
        n = (self.norm1(x.float()) * (1 + scale) + shift).type_as(x)
        a = self.attn1(n, rotary_emb=rotary_emb)
        x = (x.float() + a * gate).type_as(x)
        n = self.norm2(x.float()).type_as(x)
        a = self.attn2(n, encoder_hidden_states=enc)
        x = x + a
        n = (self.norm3(x.float()) * (1 + c_scale) + c_shift).type_as(x)
        f = self.ffn(n)
        x = (x.float() + f.float() * c_gate).type_as(x)
        return x


def timestep_sinusoid(t, dim=256, max_period=10000.0):
    """diffusers Timesteps(num_channels=256, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class _TimestepEmbedding(nn.Module):
    def __init__(self, d_in, dim):
        super().__init__()
        self.linear_1 = nn.Linear(d_in, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class _TextProjection(nn.Module):
    def __init__(self, d_in, dim):
        super().__init__()
        self.linear_1 = nn.Linear(d_in, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.gelu(self.linear_1(x), approximate="tanh"))


class WanTimeTextEmbedding(nn.Module):
    def __init__(self, dim, freq_dim, text_dim):
        super().__init__()
        self.freq_dim = freq_dim
        self.time_embedder = _TimestepEmbedding(freq_dim, dim)
        self.time_proj = nn.Linear(dim, dim * 6)
        self.text_embedder = _TextProjection(text_dim, dim)

    def forward(self, timestep, enc):
        t = timestep_sinusoid(timestep, self.freq_dim)
        temb = self.time_embedder(t.to(self.time_embedder.linear_1.weight.dtype)).type_as(enc)
        proj = self.time_proj(F.silu(temb))
        return temb, proj, self.text_embedder(enc)


def rope_freqs(cfg: WanConfig, ppf, pph, ppw, device):
    """WanRotaryPosEmbed: per-axis complex rotations (float64), axes dims (t, h, w) = (d - 4 (d // 6), 2 (d // 6), 2 (d // 6))."""
    d = cfg.attention_head_dim
    h_dim = w_dim = 2 * (d // 6)
    t_dim = d - h_dim - w_dim
    parts = []
    for dim, n, shape in ((t_dim, ppf, (ppf, 1, 1)), (h_dim, pph, (1, pph, 1)), (w_dim, ppw, (1, 1, ppw))):
        freqs = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float64, device=device)[: dim // 2] / dim))
        ang = torch.outer(torch.arange(n, dtype=torch.float64, device=device), freqs)
        cis = torch.polar(torch.ones_like(ang), ang)
        parts.append(cis.view(*shape, -1).expand(ppf, pph, ppw, -1))
    return torch.cat(parts, dim=-1).reshape(1, 1, ppf * pph * ppw, -1)


class WanTransformer3DModel(nn.Module):
    def __init__(self, cfg: WanConfig = None):
        super().__init__()
        cfg = cfg or wan_1_3b_config()
        self.cfg = cfg
        dim = cfg.inner_dim
        self.patch_embedding = nn.Conv3d(cfg.in_channels, dim, kernel_size=cfg.patch_size, stride=cfg.patch_size)
        self.condition_embedder = WanTimeTextEmbedding(dim, cfg.freq_dim, cfg.text_dim)
        self.blocks = nn.ModuleList([WanTransformerBlock(dim, cfg.ffn_dim, cfg.num_attention_heads, cfg.eps)
                                     for _ in range(cfg.num_layers)])
        self.norm_out = FP32LayerNorm(dim, cfg.eps, elementwise_affine=False)
        self.proj_out = nn.Linear(dim, cfg.out_channels * math.prod(cfg.patch_size))
        self.scale_shift_table = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)
        self.gradient_checkpointing = False

    def forward(self, hidden_states, timestep, encoder_hidden_states, return_dict=False, **kw):
        b, c, f, h, w = hidden_states.shape
        pt, ph, pw = self.cfg.patch_size
        ppf, pph, ppw = f // pt, h // ph, w // pw
        rotary = rope_freqs(self.cfg, ppf, pph, ppw, hidden_states.device)
        x = self.patch_embedding(hidden_states).flatten(2).transpose(1, 2)
        temb, proj, enc = self.condition_embedder(timestep, encoder_hidden_states)
        temb6 = proj.unflatten(1, (6, -1))
        for blk in self.blocks:
            if self.gradient_checkpointing and torch.is_grad_enabled():
                x = torch.utils.checkpoint.checkpoint(blk, x, enc, temb6, rotary, use_reentrant=False)
            else:
                x = blk(x, enc, temb6, rotary)
        shift, scale = (self.scale_shift_table + temb.unsqueeze(1)).chunk(2, dim=1)
        x = (self.norm_out(x.float()) * (1 + scale) + shift).type_as(x)
        x = self.proj_out(x)
        x = x.reshape(b, ppf, pph, ppw, pt, ph, pw, -1).permute(0, 7, 1, 4, 2, 5, 3, 6)
        out = x.flatten(6, 7).flatten(4, 5).flatten(2, 3)
        return (out,)


def init_synthetic_(model: nn.Module, seed: int = 0, std: float = 0.02):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("norm_q.weight") or name.endswith("norm_k.weight") or name.endswith("norm2.weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif "scale_shift_table" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return model


def wan_predict(model, noisy_latents, timesteps, text_embeds):
    """`Wan21.get_noise_prediction` (toolkit/models/wan21/wan21.py:578-603): raw timesteps, UMT5 embeddings."""
    return model(hidden_states=noisy_latents, timestep=timesteps, encoder_hidden_states=text_embeds, return_dict=False)[0]
