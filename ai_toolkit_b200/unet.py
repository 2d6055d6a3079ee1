"""SD1.5 / SDXL `UNet2DConditionModel` host for the B200 Transformer2D engine (BASELINE.json configs[0], [1]).

HYBRID, and labelled as such everywhere it is measured: the adapter-bearing `Transformer2DModel` stacks (the reference's
default LoRA target, toolkit/kohya_lora.py:750 -- in SDXL they hold ~80 % of the UNet's FLOPs) run on the hand-written
kernels (`unet_blocks.Transformer2DEngine`); the FROZEN body that carries no adapters under the default target list
(ResnetBlock2D, Down/Upsample2D, conv_in / conv_out, time / text-time embedding MLPs) is ordinary eager PyTorch (cuDNN /
cuBLAS library kernels) with diffusers' module names, and autograd connects the two.  A full UNet engine (GroupNorm-SiLU-conv
fusions, the body under CUDA graphs) is not built; see DESIGN.md section 7.  With `conv_lora_dim` the ResNet / sampler convs get
k x k adapters through the per-module seam (`LoRAModule.forward` -> fused GEMM over im2col rows).

`UNetLoRATrainStep` is `SDTrainer.hook_train_loop` for this model: DDPM add_noise with integer timesteps (kernel) ->
`predict_noise` (toolkit/stable_diffusion_model.py:1968-2070 SDXL `added_cond_kwargs` / :2260-2265 SD1.5) -> eps / v MSE with the
per-sample weights of `calculate_loss` (kernel) -> backward -> clip + AdamW (+ EMA) over the flat LoRA buffers (kernels).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import calc_loss, ops
from .samplers import DDPMTable
from .unet_blocks import Transformer2DModel


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: tuple = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attn_layers: tuple = (1, 1, 1, 0)   # transformer depth per down level (0: DownBlock2D / UpBlock2D without attention)
    heads: tuple = (8, 8, 8, 8)         # diffusers `attention_head_dim` = number of heads per level
    cross_attention_dim: int = 768
    use_linear_projection: bool = False
    addition_embed: bool = False        # SDXL `text_time`
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    norm_groups: int = 32


def sdxl_config() -> UNetConfig:
    """SDXL-base (public config): 3 levels, transformer depth (0, 2, 10), heads (5, 10, 20) -> head dim 64 everywhere."""
    return UNetConfig(block_out_channels=(320, 640, 1280), attn_layers=(0, 2, 10), heads=(5, 10, 20), cross_attention_dim=2048,
                      use_linear_projection=True, addition_embed=True)


def sd15_config() -> UNetConfig:
    """SD1.5: 8 heads per level -> head dims 40 / 80 (tcgen05 kernels, zero-padded to 128) and 160 at the 1280-channel levels
    (CUDA-core kernel csrc/small_attn.cu: 256 / 64 tokens at 512^2)."""
    return UNetConfig()


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
    def __init__(self, in_chs, cout, temb_ch, n_attn_layers, heads, cross_dim, linear, sampler, groups, device):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ci, cout, temb_ch, groups) for ci in in_chs])
        self.attentions = None
        if n_attn_layers:
            self.attentions = nn.ModuleList([Transformer2DModel(heads, cout // heads, cout, n_attn_layers, cross_dim, linear, groups,
                                                                device=device) for _ in in_chs])
        if sampler == "down":
            self.downsamplers = nn.ModuleList([Downsample2D(cout)])
        elif sampler == "up":
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])


class _MLP(nn.Module):
    def __init__(self, d_in, dim):
        super().__init__()
        self.linear_1 = nn.Linear(d_in, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


def timestep_embedding(t, dim, max_period=10000.0):
    """diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class UNet2DConditionModel(nn.Module):
    """Same class / module / parameter names as diffusers' (checkpoints load with `load_state_dict`)."""

    def __init__(self, cfg: UNetConfig = None, device=None, dtype=torch.bfloat16):
        super().__init__()
        cfg = cfg or sdxl_config()
        self.cfg = cfg
        ch = cfg.block_out_channels
        temb_ch = ch[0] * 4
        g = cfg.norm_groups
        n = len(ch)
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            with torch.device(device if device is not None else "cpu"):
                self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
                self.time_embedding = _MLP(ch[0], temb_ch)
                if cfg.addition_embed:
                    self.add_embedding = _MLP(cfg.projection_class_embeddings_input_dim, temb_ch)
                self.down_blocks = nn.ModuleList()
                cin = ch[0]
                for i, cout in enumerate(ch):
                    in_chs = [cin] + [cout] * (cfg.layers_per_block - 1)
                    self.down_blocks.append(_Block(in_chs, cout, temb_ch, cfg.attn_layers[i], cfg.heads[i], cfg.cross_attention_dim,
                                                   cfg.use_linear_projection, "down" if i < n - 1 else None, g, device))
                    cin = cout
                self.mid_block = _Block([ch[-1], ch[-1]], ch[-1], temb_ch, max(cfg.attn_layers[-1], 1), cfg.heads[-1],
                                        cfg.cross_attention_dim, cfg.use_linear_projection, None, g, device)
                self.mid_block.attentions = nn.ModuleList([self.mid_block.attentions[0]])  # resnet, attention, resnet
                rch, rattn, rheads = list(reversed(ch)), list(reversed(cfg.attn_layers)), list(reversed(cfg.heads))
                self.up_blocks = nn.ModuleList()
                prev_ch = rch[0]
                for i, cout in enumerate(rch):
                    skip_in = rch[min(i + 1, n - 1)]
                    in_chs = [(prev_ch if j == 0 else cout) + (skip_in if j == cfg.layers_per_block else cout)
                              for j in range(cfg.layers_per_block + 1)]
                    self.up_blocks.append(_Block(in_chs, cout, temb_ch, rattn[i], rheads[i], cfg.cross_attention_dim,
                                                 cfg.use_linear_projection, "up" if i < n - 1 else None, g, device))
                    prev_ch = cout
                self.conv_norm_out = nn.GroupNorm(g, ch[0], eps=1e-5)
                self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)
        finally:
            torch.set_default_dtype(prev)
        self.requires_grad_(False)

    @property
    def device(self):
        return self.conv_in.weight.device

    def init_synthetic_(self, seed: int = 0, std: float = 0.02):
        g = torch.Generator(device=self.device).manual_seed(seed)
        with torch.no_grad():
            for name, p in self.named_parameters():
                r = torch.randn(p.shape, generator=g, device=p.device, dtype=torch.float32)
                p.copy_(1.0 + 0.1 * r if ("norm" in name and name.endswith(".weight") and p.dim() == 1) else r * std)
        return self

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs=None, return_dict=False, **kw):
        cfg = self.cfg
        dt = self.conv_in.weight.dtype
        t = timestep.reshape(-1).expand(sample.shape[0])
        temb = self.time_embedding(timestep_embedding(t, cfg.block_out_channels[0]).to(dt))
        if cfg.addition_embed:
            tid = added_cond_kwargs["time_ids"]
            te = timestep_embedding(tid.flatten(), cfg.addition_time_embed_dim).reshape(tid.shape[0], -1)
            temb = temb + self.add_embedding(torch.cat([added_cond_kwargs["text_embeds"].to(dt), te.to(dt)], dim=-1))
        ctx = encoder_hidden_states.to(dt)
        h = self.conv_in(sample.to(dt))
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
        h = mb.attentions[0](mb.resnets[0](h, temb), ctx)
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


# "SD1.5/SDXL: take F_step from torch.utils.flop_counter.FlopCounterMode over the oracle fwd+bwd (no
# checkpointing)".  Counted on the meta device over oracle/unet_ref.py at C2 (bs=2, 128x128 latents, 77 text tokens): forward
# 13.522 TF, forward + backward (dX through the frozen layers) 28.443 TF -> per sample; tests/test_unet_blocks.py recounts it.
SDXL_STEP_FLOPS_PER_SAMPLE = 28.4431220736e12 / 2
# the same count for SD1.5 at C1 (bs=1, 64x64 latents = 512^2, 77 text tokens of width 768): forward 0.8033 TF, fwd + bwd 1.7260 TF
SD15_STEP_FLOPS_PER_SAMPLE = 1.72603375616e12


def unet_flops(cfg: UNetConfig, B, H, W, Lc=77):
    """Algorithmic FLOPs of ONE forward (2 per MAC): convs / Linears of the body, and per Transformer2DModel the projections,
    the 10 block Linears and the two attentions.  The step counts 3x (frozen layers: dX only = 1x forward; attention 2.5x;
    adapters ~0): reported as 2 * body_and_linears + 3.5 * attention like SURVEY.md section 8d."""
    ch = cfg.block_out_channels
    temb = ch[0] * 4
    lin = attn = 0.0

    def conv(ci, co, k, h, w):
        return 2.0 * B * h * w * ci * co * k * k

    def t2d(c, depth, heads, h, w):
        L = h * w
        f_lin = 2 * 2.0 * B * L * c * c  # proj_in / proj_out
        f_lin += depth * (2.0 * B * L * c * c * 6 + 2.0 * B * Lc * cfg.cross_attention_dim * c * 2 + 2.0 * B * L * c * 8 * c
                          + 2.0 * B * L * 4 * c * c)
        f_att = depth * (4.0 * B * L * L * c + 4.0 * B * L * Lc * c)
        return f_lin, f_att

    def resnet(ci, co, h, w):
        return conv(ci, co, 3, h, w) + conv(co, co, 3, h, w) + (conv(ci, co, 1, h, w) if ci != co else 0) + 2.0 * B * temb * co

    h, w = H, W
    lin += conv(cfg.in_channels, ch[0], 3, h, w)
    cin = ch[0]
    n = len(ch)
    for i, co in enumerate(ch):
        for j in range(cfg.layers_per_block):
            lin += resnet(cin if j == 0 else co, co, h, w)
            if cfg.attn_layers[i]:
                a, b = t2d(co, cfg.attn_layers[i], cfg.heads[i], h, w)
                lin, attn = lin + a, attn + b
        cin = co
        if i < n - 1:
            lin += conv(co, co, 3, h // 2, w // 2)
            h, w = h // 2, w // 2
    lin += 2 * resnet(ch[-1], ch[-1], h, w)
    a, b = t2d(ch[-1], max(cfg.attn_layers[-1], 1), cfg.heads[-1], h, w)
    lin, attn = lin + a, attn + b
    rch, rattn, rheads = list(reversed(ch)), list(reversed(cfg.attn_layers)), list(reversed(cfg.heads))
    prev = rch[0]
    for i, co in enumerate(rch):
        skip_in = rch[min(i + 1, n - 1)]
        for j in range(cfg.layers_per_block + 1):
            ci = (prev if j == 0 else co) + (skip_in if j == cfg.layers_per_block else co)
            lin += resnet(ci, co, h, w)
            if rattn[i]:
                a, b = t2d(co, rattn[i], rheads[i], h, w)
                lin, attn = lin + a, attn + b
        prev = co
        if i < n - 1:
            h, w = h * 2, w * 2
            lin += conv(co, co, 3, h, w)
    lin += conv(ch[0], cfg.out_channels, 3, h, w)
    return 2 * lin + 3.5 * attn, lin, attn


class UNetLoRATrainStep:
    """One optimizer step of UNet (SDXL / SD1.5-style) LoRA training: see the module docstring.

    `use_cuda_graph=True`: the batch lives in static device buffers and the WHOLE step (zero-grad, add_noise, eager frozen body +
    engine blocks forward, loss, autograd backward, clip + AdamW + re-pack) is captured into one CUDA graph after two eager
    steps -- SDXL has ~5,500 launches of this repo's kernels plus the eager body's per step, and the host cannot issue them
    as fast as the GPU retires them (measured: 132 ms eager-launched).  The per-sample loss coefficients are host table
    gathers (as in the reference) copied into static device vectors before the replay."""

    def __init__(self, unet, network, optimizer, *, prediction_type="epsilon", min_snr_gamma=None, snr_gamma=None,
                 use_cuda_graph=False):
        self.unet, self.network, self.optimizer = unet, network, optimizer
        self.dev = getattr(unet, "device", None) or next(unet.parameters()).device
        cfg = getattr(unet, "cfg", None)  # this package's host; a diffusers UNet (adopt_unet_transformers) carries `.config`
        self.addition_embed = bool(cfg.addition_embed) if cfg is not None else \
            getattr(getattr(unet, "config", None), "addition_embed_type", None) == "text_time"
        self.table = DDPMTable(prediction_type=prediction_type, device=self.dev)
        self.min_snr_gamma, self.snr_gamma = min_snr_gamma, snr_gamma
        self.use_cuda_graph = use_cuda_graph
        self.loss_host = torch.zeros(1, dtype=torch.float32)
        if torch.device(self.dev).type == "cuda":
            self.loss_host = self.loss_host.pin_memory()
        self.buf = None
        self._graph = None
        self._warm = 0

    def time_ids(self, B, H, W):
        """`get_time_ids_from_latents` (stable_diffusion_model.py:1824-1852): (h, w, 0, 0, h, w) in pixels."""
        return torch.tensor([[H * 8, W * 8, 0, 0, H * 8, W * 8]] * B, device=self.dev, dtype=torch.float32)

    # -- static batch buffers ---------------------------------------------------------------------------------------
    def load_batch(self, latents, noise, timesteps, text_embeds, pooled_embeds=None, loss_multiplier=None):
        """Host (pinned) or device tensors -> the static device buffers (asynchronous copies on the current stream).
        `timesteps` int64 [B]: pass a HOST tensor to keep the step free of device->host syncs."""
        B, _, H, W = latents.shape
        if self.buf is None:
            dev = self.dev
            self.buf = {
                "latents": torch.empty(latents.shape, device=dev, dtype=torch.bfloat16),
                "noise": torch.empty(noise.shape, device=dev, dtype=torch.bfloat16),
                "timesteps": torch.empty(B, device=dev, dtype=torch.int64),
                "text": torch.empty(text_embeds.shape, device=dev, dtype=torch.bfloat16),
                "pooled": None if pooled_embeds is None else torch.empty(pooled_embeds.shape, device=dev, dtype=torch.bfloat16),
                "time_ids": self.time_ids(B, H, W),
                "coef_noise": torch.ones(B, device=dev), "coef_latent": torch.zeros(B, device=dev),
                "sample_weight": torch.ones(B, device=dev), "loss_ws": torch.zeros(B + 1, device=dev),
            }
        b = self.buf
        if tuple(latents.shape) != tuple(b["latents"].shape):
            raise ValueError(f"batch shape {tuple(latents.shape)} differs from the step's static buffers {tuple(b['latents'].shape)}")
        b["latents"].copy_(latents, non_blocking=True)
        b["noise"].copy_(noise, non_blocking=True)
        if timesteps.is_floating_point():  # the trainer hands the scheduler's float timesteps over (integers for DDPM)
            timesteps = timesteps.round().to(torch.int64)
        b["timesteps"].copy_(timesteps, non_blocking=True)
        b["text"].copy_(text_embeds, non_blocking=True)
        if pooled_embeds is not None:
            b["pooled"].copy_(pooled_embeds, non_blocking=True)
        v = calc_loss.loss_vectors(timesteps, is_flow_matching=False, prediction_type=self.table.prediction_type,
                                   ddpm_table=self.table, min_snr_gamma=self.min_snr_gamma, snr_gamma=self.snr_gamma,
                                   loss_multiplier=loss_multiplier, device="cpu")
        for k in ("coef_noise", "coef_latent", "sample_weight"):
            if v[k] is None:
                b[k].fill_(1.0)
            else:
                src = v[k].pin_memory() if torch.device(self.dev).type == "cuda" else v[k]
                b[k].copy_(src, non_blocking=True)

    def _step(self, zero=True, step=True):
        net, opt, b = self.network, self.optimizer, self.buf
        if zero:
            opt.zero_grad()
        noisy = ops.ddpm_add_noise(b["latents"], b["noise"], b["timesteps"], self.table.device_table)
        added = None
        if self.addition_embed:
            added = {"text_embeds": b["pooled"], "time_ids": b["time_ids"]}
        net.is_active = True
        try:
            pred = self.unet(noisy, b["timesteps"].float(), b["text"], added_cond_kwargs=added)[0]
            _, _, dpred = ops.train_loss(pred.contiguous(), b["latents"], b["noise"], coef_noise=b["coef_noise"],
                                         coef_latent=b["coef_latent"], sample_weight=b["sample_weight"], pack=False,
                                         loss_ws=b["loss_ws"])
            pred.backward(dpred)
        finally:
            net.is_active = False
        if step:
            opt.step()

    def run(self, latents=None, noise=None, timesteps=None, text_embeds=None, pooled_embeds=None, loss_multiplier=None, *,
            first_micro_batch=True, last_micro_batch=True):
        """latents / noise [B, 4, H, W] bf16, timesteps int64 [B], text_embeds [B, 77, Dc] bf16 (+ pooled [B, 1280] for SDXL);
        with no arguments the resident batch is stepped again.  Returns the device loss scalar (no host sync).
        Gradient accumulation as in `FluxLoRATrainStep.run`: gradients of the micro-batches are summed, the optimizer runs after
        the last one (those partial steps are launched eagerly; the CUDA graph holds the whole single-batch step)."""
        if latents is not None:
            self.load_batch(latents, noise, timesteps, text_embeds, pooled_embeds, loss_multiplier)
        B = self.buf["latents"].shape[0]
        self.optimizer.sync_hyper()
        whole = first_micro_batch and last_micro_batch
        if not self.use_cuda_graph or self._warm < 2 or not whole:
            self._step(zero=first_micro_batch, step=last_micro_batch)
            self._warm += 1 if whole else 0
            return self.buf["loss_ws"][B:B + 1]
        if self._graph is None:
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._step()
        self._graph.replay()
        return self.buf["loss_ws"][B:B + 1]

    # -- save / resume from the device flat buffers (BaseSDTrainProcess.save :505-721, run :2046-2060, :2189-2222) ----
    def save(self, save_root, name, step, epoch=0, dtype=torch.float16, **kw):
        """`{name}_{step:09d}.safetensors` with the kohya keys (`lora_unet_*.lora_down/.lora_up/.alpha`) + `optimizer.pt`."""
        from . import checkpoint

        if torch.device(self.dev).type == "cuda":
            torch.cuda.synchronize()
        return checkpoint.save_checkpoint(self.network, self.optimizer, save_root, name, step=step, epoch=epoch, dtype=dtype, **kw)

    def resume(self, save_root, name, **kw):
        """-> (path, step, epoch); the captured graph stays valid (same flat buffers), the operand packs are refreshed."""
        from . import checkpoint

        out = checkpoint.resume(self.network, self.optimizer, save_root, name, **kw)
        self.network.mark_params_changed()
        if torch.device(self.dev).type == "cuda":
            self.network.refresh_packs(force=True)
        return out

    def hook_train_loop(self, batch) -> OrderedDict:
        loss = self.run(batch["latents"], batch["noise"], batch["timesteps"], batch["text_embeds"], batch.get("pooled_embeds"))
        self.loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return OrderedDict(loss=float(self.loss_host[0]))
