"""Frozen FLUX.1 DiT with diffusers' module names, executed by the B200 engine.

The reference trains LoRAs against `diffusers.FluxTransformer2DModel` (third-party, pinned at
c943837899b16cbae2f619b8dd4f7bb6f07dd81a in requirements.txt:5; loaded at
toolkit/stable_diffusion_model.py:667, called at :2192-2205).  This class has the SAME class name, module
tree and parameter names, so that (a) diffusers checkpoints load with `load_state_dict`, (b) the
reference's module discovery (`toolkit/lora_special.py:484-489, 531-567`) finds the same 494 Linear layers
and produces the same `lora_name`s / saved keys.  The modules hold parameters only; the arithmetic is the
hand-written kernels behind `flux_engine.FluxEngine` (no eager PyTorch math on the hot path).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn


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
    """FLUX.1-dev (dims as in extensions_built_in/diffusion_models/chroma/src/model.py:37-53)."""
    return FluxConfig()


class _Holder(nn.Module):
    """Parameter container: its `forward` is never used (the engine reads the weights)."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container of the B200 FLUX engine; call FluxTransformer2DModel.forward")


class _RMSNormW(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = 1e-6


class _TimestepEmbedding(_Holder):
    def __init__(self, in_dim, dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.linear_2 = nn.Linear(dim, dim)


class _CombinedEmbeddings(_Holder):
    def __init__(self, dim, pooled_dim, guidance):
        super().__init__()
        self.timestep_embedder = _TimestepEmbedding(256, dim)
        if guidance:
            self.guidance_embedder = _TimestepEmbedding(256, dim)
        self.text_embedder = _TimestepEmbedding(pooled_dim, dim)


class _AdaNorm(_Holder):
    def __init__(self, dim, mult):
        super().__init__()
        self.linear = nn.Linear(dim, mult * dim)


class _GELUProj(_Holder):
    def __init__(self, d_in, d_out):
        super().__init__()
        self.proj = nn.Linear(d_in, d_out)


class _FeedForward(_Holder):
    def __init__(self, dim, inner):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim)])


class _Attention(_Holder):
    def __init__(self, dim, head_dim, added_kv, pre_only):
        super().__init__()
        self.to_q = nn.Linear(dim, dim)
        self.to_k = nn.Linear(dim, dim)
        self.to_v = nn.Linear(dim, dim)
        self.norm_q = _RMSNormW(head_dim)
        self.norm_k = _RMSNormW(head_dim)
        if added_kv:
            self.add_q_proj = nn.Linear(dim, dim)
            self.add_k_proj = nn.Linear(dim, dim)
            self.add_v_proj = nn.Linear(dim, dim)
            self.norm_added_q = _RMSNormW(head_dim)
            self.norm_added_k = _RMSNormW(head_dim)
            self.to_add_out = nn.Linear(dim, dim)
        if not pre_only:
            self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])


class FluxTransformerBlock(_Holder):
    def __init__(self, dim, head_dim, inner):
        super().__init__()
        self.norm1 = _AdaNorm(dim, 6)
        self.norm1_context = _AdaNorm(dim, 6)
        self.attn = _Attention(dim, head_dim, added_kv=True, pre_only=False)
        self.ff = _FeedForward(dim, inner)
        self.ff_context = _FeedForward(dim, inner)


class FluxSingleTransformerBlock(_Holder):
    def __init__(self, dim, head_dim, inner):
        super().__init__()
        self.norm = _AdaNorm(dim, 3)
        self.proj_mlp = nn.Linear(dim, inner)
        self.proj_out = nn.Linear(dim + inner, dim)
        self.attn = _Attention(dim, head_dim, added_kv=False, pre_only=True)


class FluxTransformer2DModel(nn.Module):
    """Same name as diffusers' class: `LoRASpecialNetwork(is_flux=True)` targets it by name (lora_special.py:692-693)."""

    def __init__(self, cfg: FluxConfig = None, device=None, dtype=torch.bfloat16):
        super().__init__()
        cfg = cfg or flux_dev_config()
        self.cfg = cfg
        d, hd = cfg.inner_dim, cfg.attention_head_dim
        inner = int(d * cfg.mlp_ratio)
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)  # build the 12 B parameters directly in bf16 (no fp32 staging copy)
        try:
            self._build(cfg, d, hd, inner, device)
        finally:
            torch.set_default_dtype(prev)
        self.requires_grad_(False)  # frozen base (BaseSDTrainProcess.py:1900)
        self._engine = None

    def _build(self, cfg, d, hd, inner, device):
        with torch.device(device if device is not None else "cpu"):
            self.time_text_embed = _CombinedEmbeddings(d, cfg.pooled_projection_dim, cfg.guidance_embeds)
            self.context_embedder = nn.Linear(cfg.joint_attention_dim, d)
            self.x_embedder = nn.Linear(cfg.in_channels, d)
            self.transformer_blocks = nn.ModuleList([FluxTransformerBlock(d, hd, inner) for _ in range(cfg.num_layers)])
            self.single_transformer_blocks = nn.ModuleList(
                [FluxSingleTransformerBlock(d, hd, inner) for _ in range(cfg.num_single_layers)])
            self.norm_out = _AdaNorm(d, 2)
            self.proj_out = nn.Linear(d, cfg.in_channels)

    @property
    def dtype(self):
        return self.x_embedder.weight.dtype

    @property
    def device(self):
        return self.x_embedder.weight.device

    def init_synthetic_(self, seed: int = 0, std: float = 0.02):
        """Random weights of the true shapes, N(0, std^2): chunked so that the 12 B
        parameter model initialises on the device without a 48 GB fp32 temporary."""
        g = torch.Generator(device=self.device).manual_seed(seed)
        with torch.no_grad():
            for name, p in self.named_parameters():
                if name.endswith("norm_q.weight") or name.endswith("norm_k.weight") or "norm_added" in name:
                    p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32))
                else:
                    p.copy_(torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32) * std)
        return self

    @property
    def engine(self):
        if self._engine is None:
            from .flux_engine import FluxEngine

            self._engine = FluxEngine(self)
        return self._engine

    def forward(self, hidden_states, timestep, encoder_hidden_states, pooled_projections, txt_ids=None, img_ids=None,
                guidance=None, return_dict=False, **kwargs):
        """diffusers call signature (toolkit/stable_diffusion_model.py:2192-2205): `timestep` is t/1000,
        `guidance` the raw scale; returns `(sample,)`.  Differentiable w.r.t. the LoRA parameters through
        `flux_engine.FluxFunction` (the backward is the engine's hand-written backward)."""
        from .flux_engine import flux_apply

        out = flux_apply(self, hidden_states, timestep, encoder_hidden_states, pooled_projections, txt_ids, img_ids, guidance)
        return (out,)
