"""Host side of `SDTrainer.calculate_loss` (extensions_built_in/sd_trainer/SDTrainer.py:522-1052, default 'mse' path): turns
the trainer's per-step scalars into the three [B] fp32 vectors the fused loss kernel (`ops.train_loss`,
csrc/batch_ops.cu) consumes.  Everything the reference multiplies the loss by is a per-sample scalar:

    timestep weights   :923-943  `noise_scheduler.get_weights_for_timesteps` (linear_timesteps / linear_timesteps2 /
                                 timestep_type == 'weighted'), applied before the mean over [1, 2, 3]
    loss_multiplier    :994      `batch.loss_multiplier_list`, after the mean
    SNR weights        :1001-1011 `apply_snr_weight(..., gamma, fixed)` for snr_gamma / min_snr_gamma (eps / v models)

so they collapse into ONE `sample_weight[b]`; the target (flow :644-646, eps :650, v :623-625) is two coefficients.
Pinned to the unmodified reference method by tests/golden/calc_loss.pt (oracle/make_golden_loss.py)."""
from __future__ import annotations

from typing import Optional

import torch

from . import timesteps as ts


def loss_vectors(timesteps: torch.Tensor, *, is_flow_matching: bool, prediction_type: str = "epsilon", ddpm_table=None,
                 flow_table: Optional[torch.Tensor] = None, linear_timesteps: bool = False, linear_timesteps2: bool = False,
                 timestep_type: str = "sigmoid", snr_gamma: Optional[float] = None, min_snr_gamma: Optional[float] = None,
                 loss_multiplier=None, device=None):
    """-> dict(coef_noise, coef_latent, sample_weight): fp32 [B] tensors on `device` (None entries = all ones)."""
    t_cpu = timesteps.detach().to("cpu")
    B = t_cpu.numel()
    w = torch.ones(B, dtype=torch.float32)
    weighted = False
    if is_flow_matching and (linear_timesteps or linear_timesteps2 or timestep_type == "weighted"):
        if flow_table is None:
            raise ValueError("timestep weights need the scheduler's timestep table of this step")
        w = w * ts.weights_for_timesteps(flow_table.cpu(), t_cpu.float(), v2=linear_timesteps2, timestep_type=timestep_type).float()
        weighted = True
    if loss_multiplier is not None:
        lm = torch.as_tensor(loss_multiplier, dtype=torch.float32).reshape(-1)
        if not bool((lm == 1).all()):
            w = w * lm
            weighted = True
    if not is_flow_matching or ddpm_table is not None:
        # (the reference applies SNR weighting whenever the config sets it; it only makes sense with a DDPM table)
        if snr_gamma is not None and snr_gamma > 0.000001:
            w = w * ddpm_table.snr_weights(t_cpu.long(), snr_gamma, fixed=True)
            weighted = True
        elif min_snr_gamma is not None and min_snr_gamma > 0.000001:
            w = w * ddpm_table.snr_weights(t_cpu.long(), min_snr_gamma, fixed=False)
            weighted = True
    cn = cl = None
    if not is_flow_matching:
        if prediction_type == "v_prediction":
            cn, cl = ddpm_table.target_coefficients(t_cpu.long())
        else:
            cn, cl = torch.ones(B), torch.zeros(B)
    dev = device if device is not None else timesteps.device
    mv = lambda x: None if x is None else x.to(dev, torch.float32).contiguous()  # noqa: E731
    return dict(coef_noise=mv(cn), coef_latent=mv(cl), sample_weight=mv(w) if weighted else None)
