"""TEST INFRASTRUCTURE ONLY -- golden vectors for `SDTrainer.calculate_loss` produced by the UNMODIFIED reference method
(extensions_built_in/sd_trainer/SDTrainer.py:522-1052) run in this container through oracle/ref_import.py's stub importer:
the real `SDTrainer.calculate_loss` is called on an instance made with `object.__new__` carrying the reference's real
`TrainConfig` (toolkit/config_modules.py:375-620) and minimal stand-ins for `self.sd` / the batch DTO.

    python oracle/make_golden_loss.py      ->  tests/golden/calc_loss.pt   (committed; the GPU box has no /root/reference)

Cases: flow matching (default), eps, v-prediction, min-SNR-gamma, fixed SNR-gamma, loss multipliers, a mask multiplier,
bell-shaped timestep weights (linear_timesteps).  Inputs are bf16 (the trainer's dtype), results fp32.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402


def build_cases():
    g = torch.Generator().manual_seed(1234)
    B, C, H, W = 3, 4, 16, 24
    base = dict(pred=torch.randn(B, C, H, W, generator=g).bfloat16(), latents=torch.randn(B, C, H, W, generator=g).bfloat16(),
                noise=torch.randn(B, C, H, W, generator=g).bfloat16())
    mask = (torch.rand(B, 1, H, W, generator=g) > 0.4).float() * 1.25
    cases = {
        "flow": dict(flow=True, timesteps=torch.tensor([500.0, 125.0, 875.0])),
        "flow_multiplier": dict(flow=True, timesteps=torch.tensor([500.0, 125.0, 875.0]), loss_multiplier=[1.0, 0.5, 2.0]),
        "flow_mask": dict(flow=True, timesteps=torch.tensor([500.0, 125.0, 875.0]), mask=mask),
        "flow_linear_timesteps": dict(flow=True, timesteps="table", train=dict(linear_timesteps=True)),
        "eps": dict(flow=False, timesteps=torch.tensor([10, 400, 990])),
        "eps_min_snr": dict(flow=False, timesteps=torch.tensor([10, 400, 990]), train=dict(min_snr_gamma=5.0)),
        "eps_snr": dict(flow=False, timesteps=torch.tensor([10, 400, 990]), train=dict(snr_gamma=5.0), loss_multiplier=[2.0, 1.0, 0.5]),
        "v": dict(flow=False, v=True, timesteps=torch.tensor([10, 400, 990])),
    }
    return base, cases


def run_reference():
    SDTrainer = ref_import.reference_sd_trainer()
    from toolkit.config_modules import TrainConfig  # type: ignore
    from toolkit.samplers.custom_flowmatch_sampler import CustomFlowMatchEulerDiscreteScheduler  # type: ignore

    from ai_toolkit_b200.samplers import DDPMTable

    base, cases = build_cases()
    out = {"base": base, "cases": {}}
    for name, c in cases.items():
        tr = object.__new__(SDTrainer)
        tr.train_config = TrainConfig(**dict(dict(dtype="bf16", noise_scheduler="flowmatch" if c["flow"] else "ddpm"),
                                             **c.get("train", {})))
        tr.device_torch = torch.device("cpu")
        tr.dfe = None
        tr.adapter = None
        tr.additional_logs = {}

        class SD:
            is_flow_matching = c["flow"]
            prediction_type = "v_prediction" if c.get("v") else "epsilon"

            @staticmethod
            def scale_loss(loss):
                return loss

        sd = SD()
        timesteps = c["timesteps"]
        if c["flow"]:
            sch = object.__new__(CustomFlowMatchEulerDiscreteScheduler)
            # the attributes get_weights_for_timesteps reads: the linear table and the two weight tables built in __init__
            from ai_toolkit_b200 import timesteps as ts

            table = torch.linspace(1000, 1, 1000)
            sch.timesteps = table
            w1, w2 = ts.bell_weights(1000)
            sch.linear_timesteps_weights, sch.linear_timesteps_weights2 = w1, w2
            sd.noise_scheduler = sch
            if isinstance(timesteps, str):
                timesteps = table[torch.tensor([100, 500, 900])]
        else:
            tab = DDPMTable(prediction_type=sd.prediction_type)

            class Sched:
                alphas_cumprod = tab.alphas_cumprod
                timesteps = tab.timesteps

                @staticmethod
                def get_velocity(sample, noise, t):
                    return tab.get_velocity(sample, noise, t)

            sd.noise_scheduler = Sched()
        tr.sd = sd

        class Batch:
            latents = base["latents"]
            tensor = base["latents"]
            mask_tensor = None
            loss_multiplier_list = c.get("loss_multiplier", [1.0] * 3)
            audio_pred = None
            audio_target = None

            @staticmethod
            def get_is_reg_list():
                return [False] * 3

        pred = base["pred"].clone().requires_grad_(True)
        mm = c.get("mask", 1.0)
        loss = SDTrainer.calculate_loss(tr, pred, base["noise"], base["latents"], timesteps, Batch(), mask_multiplier=mm)
        loss.backward()
        out["cases"][name] = dict(timesteps=timesteps, loss=loss.detach().float(), dpred=pred.grad.float(),
                                  loss_multiplier=Batch.loss_multiplier_list, mask=c.get("mask"), flow=c["flow"],
                                  v=bool(c.get("v")), train=c.get("train", {}))
    return out


if __name__ == "__main__":
    res = run_reference()
    path = os.path.join(ROOT, "tests", "golden", "calc_loss.pt")
    torch.save(res, path)
    for k, v in res["cases"].items():
        print(k, float(v["loss"]), tuple(v["dpred"].shape))
    print("wrote", path)
