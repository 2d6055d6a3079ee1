"""Host-side noise-schedule tables for the eps / v-prediction models (SD1.5, SDXL) -- SURVEY.md section 8 rows a5', a12.

The reference builds a diffusers `DDPMScheduler` from `sd_config` (toolkit/sampler.py:31-50 via `get_sampler`, :120-185):
`beta_schedule="scaled_linear"`, beta 0.00085 -> 0.012, 1000 train steps.  diffusers is third-party and absent here, so the
published algorithm is restated:

    betas          = linspace(beta_start ** 0.5, beta_end ** 0.5, N, dtype=float32) ** 2          ("scaled_linear")
    alphas_cumprod = cumprod(1 - betas)
    add_noise      : sqrt(ac[t]) x0 + sqrt(1 - ac[t]) noise          (table cast to the sample dtype first)
    get_velocity   : sqrt(ac[t]) noise - sqrt(1 - ac[t]) sample

The per-sample SNR weights are in-tree (toolkit/train_tools.py:642-654 `get_all_snr`, :720-749 `apply_snr_weight`) and are
pinned against the live reference in tests/test_samplers.py.  Everything here is a [1000]-entry table or a [B]-entry
gather: host / tiny device tensors, no per-sample `.item()` syncs; the element-wise work is in csrc/batch_ops.cu.
"""
from __future__ import annotations

import torch

SCHEDULER_LINEAR_START = 0.00085  # toolkit/sampler.py:25-28
SCHEDULER_LINEAR_END = 0.0120
SCHEDULER_TIMESTEPS = 1000


class DDPMTable:
    """alphas_cumprod of the reference's training scheduler + the gathers the trainer does with it."""

    def __init__(self, num_train_timesteps=SCHEDULER_TIMESTEPS, beta_start=SCHEDULER_LINEAR_START, beta_end=SCHEDULER_LINEAR_END,
                 beta_schedule="scaled_linear", prediction_type="epsilon", device="cpu"):
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(f"beta_schedule {beta_schedule!r}")
        self.num_train_timesteps = int(num_train_timesteps)
        self.prediction_type = prediction_type
        self.betas = betas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        # `timesteps` of a freshly built DDPMScheduler: N-1 ... 0 (what `noise_scheduler.timesteps[0] == 1000` looks at)
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)
        self.device_table = self.alphas_cumprod.to(device).contiguous()

    def to(self, device):
        self.device_table = self.alphas_cumprod.to(device).contiguous()
        return self

    # -- BaseSDTrainProcess.process_general_training_batch, ddpm branch (:1301-1323) ---------------------------------
    def sample_timesteps(self, batch_size, min_step=0, max_step=999, generator=None, device="cpu"):
        """`torch.randint(min + 1, max - 1, (B,))` (:1306-1307) -> int64 timesteps (the scheduler's timestep VALUES equal
        their indices for DDPM, so index and value coincide)."""
        return torch.randint(min_step + 1, max_step - 1, (batch_size,), generator=generator, device=device).long()

    # -- torch restatements (the oracle side of tests; the device work is ops.ddpm_add_noise / ops.train_loss) -------
    def add_noise(self, x0, noise, timesteps):
        ac = self.alphas_cumprod.to(device=x0.device, dtype=x0.dtype)
        t = timesteps.to(x0.device)
        sa = (ac[t] ** 0.5).flatten()
        sb = ((1 - ac[t]) ** 0.5).flatten()
        while sa.dim() < x0.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * x0 + sb * noise

    def get_velocity(self, sample, noise, timesteps):
        ac = self.alphas_cumprod.to(device=sample.device, dtype=sample.dtype)
        t = timesteps.to(sample.device)
        sa = (ac[t] ** 0.5).flatten()
        sb = ((1 - ac[t]) ** 0.5).flatten()
        while sa.dim() < sample.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * noise - sb * sample

    def target_coefficients(self, timesteps, dtype=torch.bfloat16):
        """-> (coef_noise [B] fp32, coef_latent [B] fp32) for `ops.train_loss`: eps -> (1, 0); v_prediction -> the
        `get_velocity` coefficients rounded as the `dtype` tensor expression rounds them."""
        B = timesteps.numel()
        if self.prediction_type == "epsilon":
            return torch.ones(B), torch.zeros(B)
        if self.prediction_type != "v_prediction":
            raise NotImplementedError(f"prediction_type {self.prediction_type!r}")
        ac = self.alphas_cumprod.to(dtype)[timesteps.cpu()]
        return (ac ** 0.5).float(), ((1 - ac) ** 0.5).float()

    # -- toolkit/train_tools.py:642-654, :720-749 -------------------------------------------------------------------
    def all_snr(self):
        a = torch.sqrt(self.alphas_cumprod)
        s = torch.sqrt(1.0 - self.alphas_cumprod)
        return (a / s) ** 2

    def snr_weights(self, timesteps, gamma, fixed=False):
        """Per-sample factor of `apply_snr_weight`: gamma / snr[t] (fixed: `snr_gamma`) or min(gamma / snr[t], 1)
        (`min_snr_gamma`).  One vectorised gather instead of the reference's per-sample python loop."""
        offset = 1 if int(self.timesteps[0]) == 1000 else 0
        snr = self.all_snr()[(timesteps.cpu() - offset).int().long()]
        g = torch.ones_like(snr) * gamma / snr
        return g.float() if fixed else torch.minimum(g, torch.ones_like(g)).float()
