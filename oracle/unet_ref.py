"""TEST INFRASTRUCTURE ONLY -- eager PyTorch restatement of the UNet blocks of SD1.5 / SDXL (BASELINE.json configs[0], [1])
with diffusers' module names, so that `LoRASpecialNetwork` (default target `Transformer2DModel`, toolkit/kohya_lora.py:750;
`ResnetBlock2D` / `Downsample2D` / `Upsample2D` with conv_lora_dim, :751) produces the reference's kohya adapter names
(`lora_unet_down_blocks_1_attentions_0_transformer_blocks_0_attn1_to_q`, ...).

PARITY UNPINNED at the diffusers boundary: `UNet2DConditionModel` is third-party diffusers @ c943837899b16cbae2f619b8dd4f7bb6f07dd81a
(reference requirements.txt:5; called at toolkit/stable_diffusion_model.py:2049-2055 (SDXL, `added_cond_kwargs`) and :2260-2265
(SD1.5)), absent here, no golden vectors in the reference.  This file restates the published blocks:
  Transformer2DModel   GroupNorm(32, eps 1e-6) -> proj_in (1x1 conv: SD1.5 / Linear: SDXL `use_linear_projection`) ->
                       N x BasicTransformerBlock -> proj_out -> + residual
  BasicTransformerBlock  x += attn1(LN(x));  x += attn2(LN(x), context);  x += ff(LN(x)),  ff = GEGLU(dim, 4 dim) -> Linear
  Attention            to_q / to_k / to_v without bias, heads split, torch SDPA, to_out.0 with bias
  ResnetBlock2D        GN-SiLU-conv3x3 (+ time_emb_proj(SiLU(emb))) GN-SiLU-conv3x3 + shortcut
and the UNet wiring (down / mid / up blocks with skip connections, sinusoidal time embedding, SDXL `text_time` embedding).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn as nn
import torch.nn.functional as F


class Attention(nn.Module):
    def __init__(self, query_dim, cross_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, x, context=None):
        ctx = x if context is None else context
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)
        q, k, v = (t.unflatten(2, (self.heads, -1)).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).flatten(2, 3).to(q.dtype)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, d_in, d_out):
        super().__init__()
        self.proj = nn.Linear(d_in, d_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context) + x
        x = self.ff(self.norm3(x)) + x
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, num_layers=1, cross_dim=768, use_linear_projection=False, groups=32):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Linear(in_channels, inner) if use_linear_projection else nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_dim) for _ in range(num_layers)])
        self.proj_out = nn.Linear(inner, in_channels) if use_linear_projection else nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, context):
        B, C, H, W = x.shape
        res = x
        h = self.norm(x)
        if self.use_linear_projection:
            h = self.proj_in(h.permute(0, 2, 3, 1).reshape(B, H * W, C))
        else:
            h = self.proj_in(h)
            h = h.permute(0, 2, 3, 1).reshape(B, H * W, h.shape[1])
        for blk in self.transformer_blocks:
            h = blk(h, context)
        if self.use_linear_projection:
            h = self.proj_out(h).reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()
        else:
            h = self.proj_out(h.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous())
        return h + res


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_ch, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class Downsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Block(nn.Module):
    """down / up block: resnets (+ attentions) (+ down / up sampler); names as in diffusers."""

    def __init__(self, in_chs, cout, temb_ch, n_attn_layers, heads, cross_dim, linear, sampler, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ci, cout, temb_ch, groups) for ci in in_chs])
        if n_attn_layers:
            self.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout, n_attn_layers, cross_dim, linear, groups)
                                             for _ in in_chs])
        else:
            self.attentions = None
        if sampler == "down":
            self.downsamplers = nn.ModuleList([Downsample2D(cout)])
        elif sampler == "up":
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: tuple = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attn_layers: tuple = (1, 1, 1, 0)       # transformer depth per down level (0 = plain DownBlock2D)
    heads: tuple = (8, 8, 8, 8)             # diffusers `attention_head_dim` (really the number of heads) per level
    cross_attention_dim: int = 768
    use_linear_projection: bool = False
    addition_embed: bool = False            # SDXL `text_time`
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    norm_groups: int = 32
    extra: dict = field(default_factory=dict)


def sd15_config():
    return UNetConfig()


def sdxl_config():
    return UNetConfig(block_out_channels=(320, 640, 1280), attn_layers=(0, 2, 10), heads=(5, 10, 20), cross_attention_dim=2048,
                      use_linear_projection=True, addition_embed=True)


def timestep_embedding(t, dim, flip_sin_to_cos=True, shift=0.0, max_period=10000.0):
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - shift)
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class _MLP(nn.Module):
    def __init__(self, d_in, dim):
        super().__init__()
        self.linear_1 = nn.Linear(d_in, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig = None):
        super().__init__()
        cfg = cfg or sd15_config()
        self.cfg = cfg
        ch = cfg.block_out_channels
        temb_ch = ch[0] * 4
        g = cfg.norm_groups
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = _MLP(ch[0], temb_ch)
        if cfg.addition_embed:
            self.add_embedding = _MLP(cfg.projection_class_embeddings_input_dim, temb_ch)
        n = len(ch)
        self.down_blocks = nn.ModuleList()
        cin = ch[0]
        for i, cout in enumerate(ch):
            in_chs = [cin] + [cout] * (cfg.layers_per_block - 1)
            self.down_blocks.append(_Block(in_chs, cout, temb_ch, cfg.attn_layers[i], cfg.heads[i], cfg.cross_attention_dim,
                                           cfg.use_linear_projection, "down" if i < n - 1 else None, g))
            cin = cout
        self.mid_block = _Block([ch[-1], ch[-1]], ch[-1], temb_ch, max(cfg.attn_layers[-1], 1) if cfg.attn_layers[-1] or True else 0,
                                cfg.heads[-1], cfg.cross_attention_dim, cfg.use_linear_projection, None, g)
        self.mid_block.attentions = nn.ModuleList([self.mid_block.attentions[0]])  # UNetMidBlock2DCrossAttn: res, attn, res
        rch = list(reversed(ch))
        rattn, rheads = list(reversed(cfg.attn_layers)), list(reversed(cfg.heads))
        self.up_blocks = nn.ModuleList()
        prev = rch[0]
        for i, cout in enumerate(rch):
            skip_in = rch[min(i + 1, n - 1)]
            in_chs = []
            for j in range(cfg.layers_per_block + 1):
                res_skip = skip_in if j == cfg.layers_per_block else cout
                res_in = prev if j == 0 else cout
                in_chs.append(res_in + res_skip)
            self.up_blocks.append(_Block(in_chs, cout, temb_ch, rattn[i], rheads[i], cfg.cross_attention_dim,
                                         cfg.use_linear_projection, "up" if i < n - 1 else None, g))
            prev = cout
        self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=1e-5)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs=None, return_dict=False, **kw):
        cfg = self.cfg
        t = timestep.reshape(-1).expand(sample.shape[0])
        temb = self.time_embedding(timestep_embedding(t, cfg.block_out_channels[0]).to(sample.dtype))
        if cfg.addition_embed:
            tid = added_cond_kwargs["time_ids"]
            te = timestep_embedding(tid.flatten(), cfg.addition_time_embed_dim).reshape(tid.shape[0], -1)
            add = torch.cat([added_cond_kwargs["text_embeds"], te.to(sample.dtype)], dim=-1)
            temb = temb + self.add_embedding(add.to(sample.dtype))
        ctx = encoder_hidden_states
        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down_blocks:
            for j, res in enumerate(blk.resnets):
                h = res(h, temb)
                if blk.attentions is not None:
                    h = blk.attentions[j](h, ctx)
                skips.append(h)
            if hasattr(blk, "downsamplers"):
                h = blk.downsamplers[0](h)
                skips.append(h)
        mb = self.mid_block
        h = mb.resnets[0](h, temb)
        h = mb.attentions[0](h, ctx)
        h = mb.resnets[1](h, temb)
        for blk in self.up_blocks:
            for j, res in enumerate(blk.resnets):
                h = res(torch.cat([h, skips.pop()], dim=1), temb)
                if blk.attentions is not None:
                    h = blk.attentions[j](h, ctx)
            if hasattr(blk, "upsamplers"):
                h = blk.upsamplers[0](h)
        h = self.conv_out(F.silu(self.conv_norm_out(h)))
        return (h,)


def init_synthetic_(model: nn.Module, seed: int = 0, std: float = 0.02):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if ("norm" in name) and name.endswith(".weight") and p.dim() == 1:
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(torch.randn(p.shape, generator=g) * std)
    return model
