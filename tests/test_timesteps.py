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
