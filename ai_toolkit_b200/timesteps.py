"""Training-timestep tables and index sampling of the flow-matching path (host logic, index ops bit-exact).

Mirrors `CustomFlowMatchEulerDiscreteScheduler.set_train_timesteps` (toolkit/samplers/custom_flowmatch_sampler.py:107-219)
for the `linear` / `weighted` and `sigmoid` (TrainConfig default, config_modules.py:556) types and the `balanced`
index draw of `process_general_training_batch` (jobs/process/BaseSDTrainProcess.py:1301-1323).  RNG is torch's own
(Philox on CUDA), taken as given.  The `shift` family needs diffusers' scheduler config and is
not implemented here.
"""
from __future__ import annotations

import torch


def set_train_timesteps(num_timesteps: int, device, timestep_type: str = "linear", generator=None) -> torch.Tensor:
    if timestep_type in ("linear", "weighted"):
        return torch.linspace(1000, 1, num_timesteps, device=device)
    if timestep_type == "sigmoid":
        t = torch.sigmoid(torch.randn((num_timesteps,), device=device, generator=generator))
        timesteps = (1 - t) * 1000
        timesteps, _ = torch.sort(timesteps, descending=True)
        return timesteps
    raise NotImplementedError(f"timestep_type {timestep_type!r} (needs the diffusers scheduler config)")


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
