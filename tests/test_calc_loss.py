"""`SDTrainer.calculate_loss`: tests/golden/calc_loss.pt was produced by the UNMODIFIED
reference method (oracle/make_golden_loss.py).  CPU: the oracle restatement and the host-side weight vectors reproduce it.
GPU: the fused kernel `ops.train_loss` fed by `calc_loss.loss_vectors` reproduces it (loss 1e-5 rel., dpred at bf16 rounding)."""
import os

import pytest
import torch

from ai_toolkit_b200 import calc_loss
from ai_toolkit_b200.samplers import DDPMTable

GOLD = torch.load(os.path.join(os.path.dirname(__file__), "golden", "calc_loss.pt"), weights_only=False)
CASES = sorted(GOLD["cases"].keys())


def _vectors(case, device="cpu"):
    c = GOLD["cases"][case]
    tab = None if c["flow"] else DDPMTable(prediction_type="v_prediction" if c["v"] else "epsilon")
    tr = c["train"]
    return tab, calc_loss.loss_vectors(c["timesteps"], is_flow_matching=c["flow"], prediction_type="v_prediction" if c["v"] else "epsilon",
                                       ddpm_table=tab, flow_table=torch.linspace(1000, 1, 1000),
                                       linear_timesteps=tr.get("linear_timesteps", False), timestep_type="linear",
                                       snr_gamma=tr.get("snr_gamma"), min_snr_gamma=tr.get("min_snr_gamma"),
                                       loss_multiplier=c["loss_multiplier"], device=device)


@pytest.mark.parametrize("case", CASES)
def test_oracle_restatement_matches_reference_golden(case):
    from oracle import lora_ref
    c, b = GOLD["cases"][case], GOLD["base"]
    tab, v = _vectors(case)
    pred = b["pred"].clone().requires_grad_(True)
    # the single sample-weight vector is the product of the reference's three multipliers: feed it as `loss_multiplier`
    loss = lora_ref.calculate_loss_ref(pred, b["latents"], b["noise"], c["timesteps"], is_flow_matching=c["flow"],
                                       prediction_type="v_prediction" if c["v"] else "epsilon",
                                       get_velocity=None if tab is None else tab.get_velocity,
                                       loss_multiplier=v["sample_weight"], mask_multiplier=1.0 if c["mask"] is None else c["mask"])
    loss.backward()
    assert abs(float(loss) - float(c["loss"])) <= 2e-6 * abs(float(c["loss"]))
    torch.testing.assert_close(pred.grad.float(), c["dpred"], rtol=2e-2, atol=1e-6)  # the reference's grad is bf16-rounded
    if not c["flow"] and not c["v"]:
        assert v["coef_noise"].tolist() == [1.0] * 3 and v["coef_latent"].tolist() == [0.0] * 3


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_fused_loss_kernel_matches_reference_golden(case):
    from ai_toolkit_b200 import ops
    dev = "cuda:0"
    c, b = GOLD["cases"][case], GOLD["base"]
    _, v = _vectors(case, dev)
    mask = None if c["mask"] is None else c["mask"].to(dev).float().contiguous()
    tot, per, dpred = ops.train_loss(b["pred"].to(dev), b["latents"].to(dev), b["noise"].to(dev), mask=mask, pack=False, **v)
    assert abs(tot.item() - float(c["loss"])) <= 1e-5 * abs(float(c["loss"]))
    g = c["dpred"].to(dev)
    assert ((dpred.float() - g).norm() / g.norm()).item() < 6e-3
