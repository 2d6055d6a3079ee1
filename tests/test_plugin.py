"""The shipped plugin (extensions/b200_lora + ai_toolkit_b200/plugin.py), SURVEY.md section 8b rows "plugin seam" and
"trainer hooks".  With /root/reference present (this container) the UNMODIFIED registry (`toolkit.extension`) discovers the
extension and `SDTrainerB200` is built on the UNMODIFIED `SDTrainer`; its hooks are then driven on an instance made with
`object.__new__` (the constructor needs a real job / diffusers models).  On the GPU box the same behaviour checks run on a
stand-in base class."""
import os
import sys
import types
from collections import OrderedDict

import pytest
import torch

from ai_toolkit_b200 import LoRASpecialNetwork, plugin
from ai_toolkit_b200.flux import FluxConfig, FluxTransformer2DModel
from ai_toolkit_b200.optimizer import B200AdamW
from oracle import ref_import

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CFG = dict(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=64, pooled_projection_dim=32)


class _StandInTrainer:  # the hook surface of BaseSDTrainProcess / SDTrainer that the plugin builds on
    def hook_before_model_load(self):
        pass

    def hook_after_model_load(self):
        pass

    def hook_before_train_loop(self):
        self.base_before_train_loop_called = True

    def preprocess_batch(self, batch):
        return batch

    def end_of_training_loop(self):
        self.end_calls = getattr(self, "end_calls", 0) + 1


def _trainer_class():
    if ref_import.available():
        SDTrainer = ref_import.reference_sd_trainer()
        from toolkit.scheduler import get_lr_scheduler

        return plugin.make_trainer_class(SDTrainer, get_lr_scheduler), True
    return plugin.make_trainer_class(_StandInTrainer, None), False


def _network(seed=0):
    torch.manual_seed(seed)
    model = FluxTransformer2DModel(FluxConfig(**CFG), dtype=torch.float32)
    net = LoRASpecialNetwork(None, model, lora_dim=4, alpha=4, train_text_encoder=False, is_flux=True, transformer_only=True)
    net.force_to("cpu", torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, model, False, True)
    return model, net


class _FakeStep:
    def __init__(self, log):
        self.log = log
        self.loss_host = torch.zeros(1)
        self.k = 0

    def load_batch(self, latents, noise, timesteps, text, pooled):
        self.log.append(("load", tuple(latents.shape), float(timesteps[0])))

    def run(self, first_micro_batch=True, last_micro_batch=True):
        self.k += 1
        self.log.append(("run", first_micro_batch, last_micro_batch))
        return torch.tensor([float(self.k)])


def _instance(monkeypatch, use_ema=True, real=False):
    cls, is_real = _trainer_class()
    tr = object.__new__(cls)
    model, net = _network()
    tr.network = net
    tr.network_config = types.SimpleNamespace(type="lora")
    tr.train_config = types.SimpleNamespace(lr=3e-4, optimizer="adamw", optimizer_params={"weight_decay": 0.02}, max_grad_norm=0.5,
                                            ema_config=types.SimpleNamespace(use_ema=use_ema, ema_decay=0.97),
                                            lr_scheduler="constant", lr_scheduler_params={}, steps=10, cfg_scale=1.0,
                                            loss_type="mse")
    tr.sd = types.SimpleNamespace(is_flux=True, unet=model, ema=None, pipeline=None)
    tr.device_torch = torch.device("cpu")
    tr.is_grad_accumulation_step = False
    tr.adapter = None
    tr.embedding = None
    tr.ema = object() if use_ema else None
    tr.optimizer = torch.optim.AdamW(net.prepare_optimizer_params(None, 3e-4, 3e-4), lr=3e-4, eps=1e-6)
    tr.lr_scheduler = torch.optim.lr_scheduler.ConstantLR(tr.optimizer, factor=1.0, total_iters=10)

    class Acc:  # the stock loop would call this: the fused loop must not
        def clip_grad_norm_(self, *a, **k):
            raise AssertionError("clip_grad_norm_ called on top of the fused clip")

    tr.accelerator = Acc()
    log = []
    fake = _FakeStep(log)
    monkeypatch.setattr(tr, "_b200_step_for", lambda latents, text: fake, raising=False)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: types.SimpleNamespace(synchronize=lambda: None))

    def prep(batch):
        B = batch.latents.shape[0]
        return batch.latents * 0.5, torch.zeros_like(batch.latents), torch.full((B,), batch.t), ["p"] * B, None

    monkeypatch.setattr(tr, "process_general_training_batch", prep, raising=False)
    if not hasattr(tr, "end_of_training_loop") or is_real:
        monkeypatch.setattr(tr, "end_of_training_loop", lambda: None, raising=False)
    return tr, net, log, is_real


def _batch(t=500.0, B=1):
    pe = types.SimpleNamespace(text_embeds=torch.zeros(B, 8, 64), pooled_embeds=torch.zeros(B, 32))
    return types.SimpleNamespace(latents=torch.zeros(B, 16, 8, 8), prompt_embeds=pe, t=t)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_reference_registry_discovers_the_extension(monkeypatch):
    """`toolkit.extension.get_all_extensions_process_dict()` (unmodified) scanning this repository's `extensions/` folder
    returns {uid: process class}; the class is an `SDTrainer` (so `BaseSDTrainProcess.run` drives it unchanged)."""
    ref_import.install()
    monkeypatch.syspath_prepend(ROOT)
    import toolkit.extension as ext

    monkeypatch.setattr(ext, "TOOLKIT_ROOT", ROOT)
    for name in [m for m in sys.modules if m == "extensions" or m.startswith("extensions.")]:
        monkeypatch.delitem(sys.modules, name)
    procs = ext.get_all_extensions_process_dict()
    assert list(procs.keys()) == [plugin.UID]
    cls = procs[plugin.UID]
    SDTrainer = ref_import.reference_sd_trainer()
    assert issubclass(cls, SDTrainer) and cls.__name__ == "SDTrainerB200"
    for hook in ("hook_before_model_load", "hook_after_model_load", "hook_before_train_loop", "hook_train_loop"):
        assert getattr(cls, hook) is not getattr(SDTrainer, hook)
    import extensions.b200_lora as pkg

    e = pkg.AI_TOOLKIT_EXTENSIONS[0]
    assert issubclass(e, ext.Extension) and e.uid == plugin.UID and e.name


def test_setup_hands_over_optimizer_scheduler_and_ema(monkeypatch):
    tr, net, log, is_real = _instance(monkeypatch)
    old = tr.optimizer
    # a resumed run: the torch optimizer already carries state (run() loaded optimizer.pt into it)
    for p in old.param_groups[0]["params"]:
        old.state[p] = {"step": torch.tensor(5.0), "exp_avg": torch.full_like(p, 0.25), "exp_avg_sq": torch.full_like(p, 0.5)}
    tr.b200_setup()
    opt = tr.optimizer
    assert isinstance(opt, B200AdamW) and opt.network is net
    assert opt.param_groups[0]["lr"] == 3e-4 and opt.max_grad_norm == 0.5 and opt.ema_decay == 0.97
    assert opt.param_groups[0]["weight_decay"] == 0.02 and opt.param_groups[0]["eps"] == 1e-6
    assert int(opt.state_buf[0]) == 5 and float(opt.exp_avg[0]) == 0.25 and float(opt.exp_avg_sq[3]) == 0.5
    assert isinstance(tr.ema, plugin.FusedEMA) and tr.sd.ema is tr.ema
    if is_real:  # re-bound through the reference's own scheduler factory
        assert tr.lr_scheduler.optimizer is opt
    # FusedEMA: eval() swaps the shadow in (sampling / saving), train() restores the training weights
    with torch.no_grad():
        opt.ema.fill_(7.0)
    before = net.flat_params.clone()
    tr.ema.eval()
    assert float(net.flat_params[0]) == 7.0 and net._pack_dirty
    tr.ema.train()
    assert torch.equal(net.flat_params, before)
    assert tr.ema.update() is None


def test_hook_train_loop_does_not_double_clip_or_double_ema(monkeypatch):
    tr, net, log, is_real = _instance(monkeypatch)
    tr.b200_setup()
    sched_steps = []
    tr.lr_scheduler = types.SimpleNamespace(step=lambda: sched_steps.append(1))

    class Ema(plugin.FusedEMA):
        def update(self, *a, **k):
            raise AssertionError("ema.update called on top of the fused EMA")

    tr.ema = Ema(tr.optimizer)
    out = tr.hook_train_loop(_batch(t=250.0))
    assert isinstance(out, OrderedDict) and out["loss"] == 1.0
    assert log == [("load", (1, 16, 8, 8), 250.0), ("run", True, True)] and sched_steps == [1]
    # a batch list = gradient accumulation inside one call (SDTrainer.py:2250-2268): zero once, step once, mean loss
    log.clear()
    out = tr.hook_train_loop([_batch(100.0), _batch(200.0), _batch(300.0)])
    assert [e for e in log if e[0] == "run"] == [("run", True, False), ("run", False, False), ("run", False, True)]
    assert out["loss"] == pytest.approx((2.0 + 3.0 + 4.0) / 3)
    # accumulation ACROSS calls (is_grad_accumulation_step): no optimizer in the first call, no zeroing in the second
    log.clear()
    tr.is_grad_accumulation_step = True
    tr.hook_train_loop(_batch())
    tr.is_grad_accumulation_step = False
    tr.hook_train_loop(_batch())
    assert [e for e in log if e[0] == "run"] == [("run", True, False), ("run", False, True)]


def test_unsupported_configurations_fail_loudly(monkeypatch):
    tr, net, log, _ = _instance(monkeypatch)
    tr.b200_setup()
    tr.train_config.loss_type = "mae"
    with pytest.raises(NotImplementedError):
        tr.hook_train_loop(_batch())
    tr.train_config.loss_type = "mse"
    tr.sd.is_flux = False
    with pytest.raises(NotImplementedError):
        tr.hook_train_loop(_batch())
    tr.train_config.optimizer = "prodigy"
    with pytest.raises(NotImplementedError):
        tr.b200_setup()


def test_adopt_flux_transformer_shares_storage():
    """The diffusers transformer is replaced by the engine's parameter container over the SAME tensors."""
    from oracle import flux_ref

    src = flux_ref.init_synthetic_(flux_ref.FluxTransformer2DModel(flux_ref.FluxConfig(**CFG)), seed=1)
    src.config = types.SimpleNamespace(in_channels=64, num_layers=1, num_single_layers=1, attention_head_dim=128,
                                       num_attention_heads=2, joint_attention_dim=64, pooled_projection_dim=32,
                                       guidance_embeds=True, axes_dims_rope=(16, 56, 56))
    dst = plugin.adopt_flux_transformer(src)
    assert isinstance(dst, FluxTransformer2DModel) and plugin.adopt_flux_transformer(dst) is dst
    a, b = dict(src.named_parameters()), dict(dst.named_parameters())
    assert a.keys() == b.keys()
    assert all(a[k].data_ptr() == b[k].data_ptr() for k in a) and not any(p.requires_grad for p in dst.parameters())


def test_unet_route_adopts_the_transformers_and_builds_the_unet_step(monkeypatch):
    """SD1.5 / SDXL under `sd_trainer_b200`: `hook_after_model_load` moves the Transformer2DModels of the loaded UNet onto the
    engine (same tensors), the supported-configuration check lets the UNet through, and the step class is `UNetLoRATrainStep`
    with the model's prediction type and the configured min-SNR gamma."""
    from ai_toolkit_b200.unet import UNetLoRATrainStep
    from ai_toolkit_b200.unet_blocks import Transformer2DModel
    from oracle import unet_ref

    cls, _ = _trainer_class()
    tr = object.__new__(cls)
    cfg = unet_ref.UNetConfig(block_out_channels=(64, 128), attn_layers=(1, 1), heads=(1, 2), cross_attention_dim=96)
    unet = unet_ref.init_synthetic_(unet_ref.UNet2DConditionModel(cfg), seed=2)
    tr.sd = types.SimpleNamespace(is_flux=False, unet=unet, pipeline=None, prediction_type="v_prediction")
    monkeypatch.setattr(type(tr).__mro__[1], "hook_after_model_load", lambda self: None, raising=False)
    tr.hook_after_model_load()
    assert sum(isinstance(m, Transformer2DModel) for m in unet.modules()) == 2 + 2 + 1 + 3 + 3
    net = LoRASpecialNetwork(None, unet, lora_dim=4, alpha=4, train_text_encoder=False)
    net.force_to("cpu", torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, unet, False, True)
    tr.network = net
    tr.optimizer = types.SimpleNamespace()
    tr.train_config = types.SimpleNamespace(loss_type="mse", min_snr_gamma=5.0, snr_gamma=None, cfg_scale=1.0)
    tr.adapter = tr.embedding = None
    tr._b200_steps = {}
    tr._b200_check_supported(None)  # does not raise for the UNet
    step = tr._b200_step_for(torch.zeros(1, 4, 8, 8), torch.zeros(1, 77, 96))
    assert isinstance(step, UNetLoRATrainStep) and step.table.prediction_type == "v_prediction" and step.min_snr_gamma == 5.0
    assert tr._b200_step_for(torch.zeros(1, 4, 8, 8), torch.zeros(1, 77, 96)) is step


def test_preservation_runs_as_a_second_micro_batch_of_the_same_step(monkeypatch):
    """diff_output_preservation / blank_prompt_preservation (SDTrainer.py:2182-2219): normal pass, then the preservation pass
    (prior-prediction target on the class / blank embeddings, x multiplier) with gradients summed and ONE optimizer step; the
    loss is normal + preservation and both are logged."""
    tr, net, log, _ = _instance(monkeypatch)
    tr.b200_setup()
    tr.lr_scheduler = types.SimpleNamespace(step=lambda: None)
    tr.train_config.diff_output_preservation = True
    tr.train_config.diff_output_preservation_multiplier = 0.5
    tr.train_config.diff_output_preservation_class = "person"
    tr.train_config.blank_prompt_preservation = False
    tr.additional_logs = {}
    class_pe = types.SimpleNamespace(text_embeds=torch.ones(1, 8, 64), pooled_embeds=torch.ones(1, 32))
    class_pe.expand_to_batch = lambda n: class_pe
    tr.cached_dop_class_embeds = class_pe
    calls = []
    fake_main, fake_pres = _FakeStep(log), _FakeStep(log)
    fake_pres.loss_multiplier = 0.5

    def step_for(latents, text, preservation=False):
        calls.append((preservation, float(text.flatten()[0])))
        return fake_pres if preservation else fake_main

    monkeypatch.setattr(tr, "_b200_step_for", step_for, raising=False)
    out = tr.hook_train_loop(_batch(t=250.0))
    assert calls == [(False, 0.0), (True, 1.0)]                       # second pass on the class embeddings
    runs = [e for e in log if e[0] == "run"]
    assert runs == [("run", True, False), ("run", False, True)]         # zero once, optimizer once
    assert out["loss"] == pytest.approx(1.0 + 0.5 * 1.0)               # normal + multiplier x preservation
    assert tr.additional_logs == {"loss/normal": 1.0, "loss/preservation": 0.5}
    # the UNet step has no prior-target mode: fail loudly
    tr.sd.is_flux = False
    with pytest.raises(NotImplementedError):
        tr.hook_train_loop(_batch())
