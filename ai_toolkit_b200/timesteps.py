"""Training-timestep tables, index sampling and per-timestep loss weights of the flow-matching path (host logic; index
ops bit-exact, no host synchronisation).

Mirrors `CustomFlowMatchEulerDiscreteScheduler` (toolkit/samplers/custom_flowmatch_sampler.py):
  * `set_train_timesteps` :107-219 for `linear` / `weighted`, `sigmoid` (TrainConfig default, config_modules.py:556),
    `lognorm_blend`, and the `shift` family (`flux_shift`, `lumina2_shift`, `shift`);
  * the bell-shaped loss weights of `__init__` :24-57 and `get_weights_for_timesteps` :59-76 (the reference finds each
    index with `(self.timesteps == t).nonzero().item()`, one host sync per sample; here one vectorised compare);
  * `calculate_shift` :10-20;
and the `balanced` index draw of `process_general_training_batch` (jobs/process/BaseSDTrainProcess.py:1301-1323).
RNG is torch's own (Philox on CUDA), taken as given.

The `shift` family leans on diffusers' `FlowMatchEulerDiscreteScheduler` (not vendored by the reference, PARITY UNPINNED
for these three lines): `_sigma_to_t(s) = s * num_train_timesteps`, `sigma_max/min` = first/last of
`linspace(1, N, N)[::-1] / N` (statically shifted when `use_dynamic_shifting` is off) and the exponential
`time_shift(mu, 1, t) = e^mu / (e^mu + (1/t - 1))`.  `tests/test_timesteps.py` runs the reference class with exactly
these three definitions supplied.  The karras / exponential / beta sigma conversions, `shift_terminal` and
`invert_sigmas` are off in every flow-matching config the reference ships and raise here.  The `weighted` loss weights
are a 1000-entry table of the reference (toolkit/timestep_weighing/default_weighing_scheme.py) that is data, not
logic; pass it in as `table_weights`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


@dataclass
class FlowMatchSchedulerConfig:
    """The diffusers scheduler-config fields `set_train_timesteps` reads (FLUX.1-dev values as defaults)."""
    num_train_timesteps: int = 1000
    shift: float = 3.0
    use_dynamic_shifting: bool = True
    base_shift: float = 0.5
    max_shift: float = 1.15
    base_image_seq_len: int = 256
    max_image_seq_len: int = 4096
    shift_terminal: Optional[float] = None
    use_karras_sigmas: bool = False
    use_exponential_sigmas: bool = False
    use_beta_sigmas: bool = False
    invert_sigmas: bool = False


def calculate_shift(image_seq_len, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.16):
    """custom_flowmatch_sampler.py:10-20 (note the reference's own default max_shift of 1.16)."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def _sigma_range(cfg: FlowMatchSchedulerConfig):
    n = cfg.num_train_timesteps
    sig = np.linspace(1, n, n, dtype=np.float32)[::-1].copy() / n
    if not cfg.use_dynamic_shifting:
        sig = cfg.shift * sig / (1 + (cfg.shift - 1) * sig)
    return float(sig[0]), float(sig[-1])  # sigma_max, sigma_min


def set_train_timesteps(num_timesteps: int, device, timestep_type: str = "linear", generator=None, latents=None,
                        patch_size: int = 1, config: Optional[FlowMatchSchedulerConfig] = None) -> torch.Tensor:
    if timestep_type in ("linear", "weighted"):
        return torch.linspace(1000, 1, num_timesteps, device=device)
    if timestep_type == "sigmoid":
        t = torch.sigmoid(torch.randn((num_timesteps,), device=device, generator=generator))
        timesteps = (1 - t) * 1000
        timesteps, _ = torch.sort(timesteps, descending=True)
        return timesteps
    if timestep_type in ("flux_shift", "lumina2_shift", "shift"):
        cfg = config or FlowMatchSchedulerConfig()
        if cfg.shift_terminal or cfg.use_karras_sigmas or cfg.use_exponential_sigmas or cfg.use_beta_sigmas \
                or cfg.invert_sigmas:
            raise NotImplementedError("shift_terminal / karras / exponential / beta / inverted sigmas")
        n = cfg.num_train_timesteps
        sigma_max, sigma_min = _sigma_range(cfg)
        timesteps = np.linspace(sigma_max * n, sigma_min * n, num_timesteps)
        sigmas = timesteps / n
        if cfg.use_dynamic_shifting:
            if latents is None:
                raise ValueError("latents is None")
            image_seq_len = latents.shape[2] * latents.shape[3] // (patch_size ** 2)
            mu = calculate_shift(image_seq_len, cfg.base_image_seq_len, cfg.max_image_seq_len, cfg.base_shift,
                                 cfg.max_shift)
            sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1) ** 1.0)
        else:
            sigmas = cfg.shift * sigmas / (1 + (cfg.shift - 1) * sigmas)
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32, device=device)
        return sigmas * n
    if timestep_type == "lognorm_blend":
        alpha = 0.75
        lognormal = torch.distributions.LogNormal(loc=0, scale=0.333)
        t1 = lognormal.sample((int(num_timesteps * alpha),)).to(device)
        t1 = (1 - t1 / t1.max()) * 1000
        t2 = torch.linspace(1000, 1, int(num_timesteps * (1 - alpha)), device=device)
        timesteps, _ = torch.sort(torch.cat((t1, t2)), descending=True)
        return timesteps.to(torch.int)
    raise ValueError(f"Invalid timestep type: {timestep_type}")


def sample_timestep_indices(batch_size: int, device, min_noise_steps: int = 0, max_noise_steps: int = 999,
                            flowmatch: bool = True, generator=None) -> torch.Tensor:
    """`content_or_style == 'balanced'` (BaseSDTrainProcess.py:1301-1318): flowmatch draws indices in [min, max),
    other schedulers in [min + 1, max - 1)."""
    if min_noise_steps == max_noise_steps:
        return (torch.ones((batch_size,), device=device) * min_noise_steps).long()
    lo, hi = (min_noise_steps, max_noise_steps) if flowmatch else (min_noise_steps + 1, max_noise_steps - 1)
    return torch.randint(lo, hi, (batch_size,), device=device, generator=generator).long()


def timesteps_for_batch(table: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """`timesteps = noise_scheduler.timesteps[timestep_indices.long()]` (:1323)."""
    return table[indices.long()]


def bell_weights(num_timesteps: int = 1000):
    """(bsmntw, hbsmntw) of the scheduler constructor (:28-57): bell-shaped, minimum shifted to 0, mean 1; the second
    one is flat at its maximum over the second half."""
    x = torch.arange(num_timesteps, dtype=torch.float32)
    y = torch.exp(-2 * ((x - num_timesteps / 2) / num_timesteps) ** 2)
    y_shifted = y - y.min()
    bsmntw = y_shifted * (num_timesteps / y_shifted.sum())
    hbsmntw = y_shifted * (num_timesteps / y_shifted.sum())
    hbsmntw[num_timesteps // 2:] = hbsmntw[num_timesteps // 2:].max()
    return bsmntw, hbsmntw


def weights_for_timesteps(table: torch.Tensor, timesteps: torch.Tensor, v2: bool = False,
                          timestep_type: str = "linear", table_weights=None) -> torch.Tensor:
    """`get_weights_for_timesteps` (:59-76): weight of each batch timestep by its index in the training table.  One
    vectorised equality compare on the device instead of a `.nonzero().item()` per sample; like the reference it
    requires every timestep to be an element of the table."""
    eq = table.to(timesteps.device)[None, :] == timesteps[:, None]
    if eq.device.type == "cpu" and not eq.any(dim=1).all():  # (no such check on CUDA: it would be a host sync)
        raise ValueError("timestep not in the training table")
    idx = eq.to(torch.uint8).argmax(dim=1)
    if timestep_type == "weighted":
        if table_weights is None:
            raise NotImplementedError("pass the reference's default_weighing_scheme table as `table_weights`")
        w = torch.as_tensor(table_weights, dtype=timesteps.dtype, device=timesteps.device)
        return w[idx]
    w1, w2 = bell_weights(1000)  # the constructor's tables always have 1000 entries (:29)
    return (w2 if v2 else w1).to(timesteps.device)[idx].flatten()
