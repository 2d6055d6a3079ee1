"""DDPM tables / SNR weights (SURVEY.md section 8 rows a5', a12): the in-tree parts against the LIVE reference, the
diffusers part (absent third-party) against its published algorithm restated with torch ops in float64."""
import pytest
import torch

from ai_toolkit_b200.samplers import DDPMTable
from oracle import ref_import


def test_scaled_linear_table_matches_published_algorithm():
    tab = DDPMTable()
    b = (torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2)
    ac = torch.cumprod(1 - b, 0)
    assert tab.alphas_cumprod.dtype == torch.float32 and tab.alphas_cumprod.shape == (1000,)
    torch.testing.assert_close(tab.alphas_cumprod.double(), ac, rtol=2e-5, atol=1e-7)
    assert abs(float(tab.alphas_cumprod[0]) - 0.99915) < 1e-5 and float(tab.alphas_cumprod[-1]) < 5e-3
    assert tab.timesteps[0] == 999 and tab.timesteps[-1] == 0


def test_timestep_draw_is_the_reference_randint():
    tab = DDPMTable()
    g1, g2 = torch.Generator().manual_seed(7), torch.Generator().manual_seed(7)
    t = tab.sample_timesteps(5, 0, 999, generator=g1)
    assert torch.equal(t, torch.randint(1, 998, (5,), generator=g2).long())  # BaseSDTrainProcess.py:1306-1307
    assert t.dtype == torch.int64 and int(t.min()) >= 1 and int(t.max()) <= 997


def test_add_noise_and_velocity_identities():
    tab = DDPMTable(prediction_type="v_prediction")
    x, n = torch.randn(3, 4, 8, 8), torch.randn(3, 4, 8, 8)
    t = torch.tensor([1, 500, 997])
    noisy, v = tab.add_noise(x, n, t), tab.get_velocity(x, n, t)
    sa = tab.alphas_cumprod[t].sqrt().view(3, 1, 1, 1)
    sb = (1 - tab.alphas_cumprod[t]).sqrt().view(3, 1, 1, 1)
    torch.testing.assert_close(sa * noisy - sb * v, x, rtol=1e-4, atol=1e-5)  # x0 = sqrt(ac) x_t - sqrt(1-ac) v
    cn, cl = tab.target_coefficients(t, torch.float32)
    torch.testing.assert_close(cn.view(3, 1, 1, 1) * n - cl.view(3, 1, 1, 1) * x, v)
    e = DDPMTable().target_coefficients(t)
    assert e[0].tolist() == [1, 1, 1] and e[1].tolist() == [0, 0, 0]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_snr_weights_identical_to_live_reference():
    ref_import.install()
    from toolkit.train_tools import apply_snr_weight, get_all_snr

    tab = DDPMTable()

    class Sched:  # the two attributes the reference functions read
        alphas_cumprod = tab.alphas_cumprod
        timesteps = tab.timesteps

    assert torch.equal(get_all_snr(Sched(), "cpu"), tab.all_snr())
    t = torch.tensor([3, 250, 640, 997])
    loss = torch.rand(4) + 0.5
    for gamma, fixed in ((5.0, False), (5.0, True), (1.0, False)):
        want = apply_snr_weight(loss.clone(), t, Sched(), gamma, fixed=fixed)
        assert torch.equal(want, loss * tab.snr_weights(t, gamma, fixed=fixed))
