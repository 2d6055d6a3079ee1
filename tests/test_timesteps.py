"""Timestep table / index ops vs the reference statements restated inline (bit-exact on the same RNG state)."""
import torch

from ai_toolkit_b200 import timesteps as ts


def test_linear_table_and_indexing_bit_exact():
    table = ts.set_train_timesteps(1000, "cpu", "linear")
    assert torch.equal(table, torch.linspace(1000, 1, 1000))  # custom_flowmatch_sampler.py:117
    g1, g2 = torch.Generator().manual_seed(7), torch.Generator().manual_seed(7)
    idx = ts.sample_timestep_indices(5, "cpu", 0, 999, flowmatch=True, generator=g1)
    ref = torch.randint(0, 999, (5,), generator=g2).long()    # BaseSDTrainProcess.py:1312-1317 with flowmatch bounds
    assert torch.equal(idx, ref)
    assert torch.equal(ts.timesteps_for_batch(table, idx), table[ref])
    idx2 = ts.sample_timestep_indices(4, "cpu", 10, 990, flowmatch=False, generator=torch.Generator().manual_seed(1))
    assert torch.equal(idx2, torch.randint(11, 989, (4,), generator=torch.Generator().manual_seed(1)))
    assert torch.equal(ts.sample_timestep_indices(3, "cpu", 500, 500), torch.full((3,), 500))


def test_sigmoid_table_matches_reference_statements():
    g1, g2 = torch.Generator().manual_seed(3), torch.Generator().manual_seed(3)
    table = ts.set_train_timesteps(1000, "cpu", "sigmoid", generator=g1)
    t = torch.sigmoid(torch.randn((1000,), generator=g2))   # custom_flowmatch_sampler.py:123-131
    ref, _ = torch.sort((1 - t) * 1000, descending=True)
    assert torch.equal(table, ref)
    assert (table[:-1] >= table[1:]).all() and table.max() <= 1000 and table.min() >= 0


# ---------------------------------------------------------------------------------------------------------------------
# The same functions against the LIVE reference scheduler class (run through the stub importer; skipped on the GPU box)
# ---------------------------------------------------------------------------------------------------------------------
import math  # noqa: E402

import numpy as np  # noqa: E402
import pytest  # noqa: E402

from oracle import ref_import  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")


def _ref_scheduler(cfg=None):
    """The UNMODIFIED CustomFlowMatchEulerDiscreteScheduler.  Its diffusers base class is a stub here, so the three
    base-class members the `shift` family reads are supplied as diffusers publishes them (the unpinned part)."""
    ref_import.install()
    from toolkit.samplers.custom_flowmatch_sampler import CustomFlowMatchEulerDiscreteScheduler as Ref

    class Cfg(dict):
        __getattr__ = dict.__getitem__

    class WithBase(Ref):
        def __init__(self, c):
            super().__init__()
            self.config = Cfg(num_train_timesteps=c.num_train_timesteps, use_dynamic_shifting=c.use_dynamic_shifting,
                              base_image_seq_len=c.base_image_seq_len, max_image_seq_len=c.max_image_seq_len,
                              base_shift=c.base_shift, max_shift=c.max_shift, shift_terminal=None, use_karras_sigmas=False,
                              use_exponential_sigmas=False, use_beta_sigmas=False, invert_sigmas=False)
            n = c.num_train_timesteps
            sig = torch.from_numpy(np.linspace(1, n, n, dtype=np.float32)[::-1].copy()) / n
            self.shift = c.shift
            if not c.use_dynamic_shifting:
                sig = c.shift * sig / (1 + (c.shift - 1) * sig)
            self.sigma_min, self.sigma_max = sig[-1].item(), sig[0].item()

        def _sigma_to_t(self, sigma):
            return sigma * self.config.num_train_timesteps

        def time_shift(self, mu, sigma, t):
            return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    return WithBase(cfg or ts.FlowMatchSchedulerConfig())


@needs_ref
def test_tables_identical_to_live_reference_linear_sigmoid_lognorm():
    ref = _ref_scheduler()
    assert torch.equal(ts.set_train_timesteps(1000, "cpu", "linear"), ref.set_train_timesteps(1000, "cpu", "linear"))
    for kind in ("sigmoid", "lognorm_blend"):
        torch.manual_seed(11)
        want = ref.set_train_timesteps(1000, "cpu", kind)
        torch.manual_seed(11)
        got = ts.set_train_timesteps(1000, "cpu", kind)
        assert got.dtype == want.dtype and torch.equal(got, want), kind
    with pytest.raises(ValueError):
        ts.set_train_timesteps(1000, "cpu", "no_such_type")


@needs_ref
@pytest.mark.parametrize("dynamic,hw,patch", [(True, (128, 128), 2), (True, (64, 96), 2), (False, (128, 128), 1)])
def test_shift_family_identical_to_live_reference(dynamic, hw, patch):
    cfg = ts.FlowMatchSchedulerConfig(use_dynamic_shifting=dynamic)
    ref = _ref_scheduler(cfg)
    latents = torch.zeros(1, 16, *hw)
    for kind in ("flux_shift", "shift"):
        want = ref.set_train_timesteps(1000, "cpu", kind, latents=latents, patch_size=patch)
        got = ts.set_train_timesteps(1000, "cpu", kind, latents=latents, patch_size=patch, config=cfg)
        assert torch.equal(got, want)
    assert got[0] == 1000 and (got[:-1] > got[1:]).all()
    if dynamic:  # FLUX at 1024^2: 4096 image tokens -> mu = max_shift
        assert abs(ts.calculate_shift(4096, 256, 4096, 0.5, 1.15) - 1.15) < 1e-12


@needs_ref
def test_loss_weights_identical_to_live_reference():
    ref = _ref_scheduler()
    w1, w2 = ts.bell_weights(1000)
    assert torch.equal(w1, ref.linear_timesteps_weights) and torch.equal(w2, ref.linear_timesteps_weights2)
    table = ref.set_train_timesteps(1000, "cpu", "linear")
    t = table[torch.tensor([0, 17, 499, 500, 998])]
    for v2 in (False, True):
        assert torch.equal(ts.weights_for_timesteps(table, t, v2=v2), ref.get_weights_for_timesteps(t, v2=v2))
    from toolkit.timestep_weighing.default_weighing_scheme import default_weighing_scheme
    got = ts.weights_for_timesteps(table, t, timestep_type="weighted", table_weights=default_weighing_scheme)
    assert torch.equal(got, ref.get_weights_for_timesteps(t, timestep_type="weighted"))
    with pytest.raises(ValueError):
        ts.weights_for_timesteps(table, torch.tensor([123.456]))
