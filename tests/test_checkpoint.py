"""Save / resume conventions: file naming, metadata, optimizer.pt, pruning of old saves, latest-path
selection — and interchange with the LIVE reference (its metadata reader, its network loader, torch AdamW)."""
import os
import time

import pytest
import torch

from ai_toolkit_b200 import LoRASpecialNetwork, checkpoint as ck
from ai_toolkit_b200.flux import FluxConfig, FluxTransformer2DModel
from ai_toolkit_b200.optimizer import B200AdamW
from oracle import ref_import

CFG = dict(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=64, pooled_projection_dim=32)
KW = dict(text_encoder=None, alpha=None, train_unet=True, train_text_encoder=False, is_flux=True, network_type="lora",
          transformer_only=True)


def _net(rank=4, seed=0):
    torch.manual_seed(seed)
    model = FluxTransformer2DModel(FluxConfig(**CFG), dtype=torch.float32)
    kw = dict(KW, alpha=rank)
    net = LoRASpecialNetwork(unet=model, lora_dim=rank, **kw)
    net.force_to("cpu", torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, model, False, True)
    with torch.no_grad():
        for l in net.unet_loras:
            l.lora_up.weight.normal_(0, 0.1)
    return model, net


def test_save_naming_metadata_pruning_and_latest(tmp_path):
    root = str(tmp_path / "out")
    _, net = _net()
    opt = B200AdamW(net, lr=2e-4)
    opt.state_buf[0] = 3
    net.multiplier = 0.5
    paths = []
    for step in (100, 200, 300):
        paths.append(ck.save_checkpoint(net, opt, root, "my_lora", step=step, epoch=1, dtype=torch.float16,
                                        max_step_saves_to_keep=2))
        time.sleep(0.02)  # distinct ctimes
    assert net.multiplier == 0.5  # restored after saving at 1.0 (BaseSDTrainProcess.py:544-556)
    assert os.path.basename(paths[0]) == "my_lora_000000100.safetensors"
    assert sorted(os.listdir(root)) == ["my_lora_000000200.safetensors", "my_lora_000000300.safetensors", "optimizer.pt"]
    meta = ck.load_metadata_from_safetensors(paths[2])
    assert meta["training_info"] == {"step": 300, "epoch": 1} and meta["ss_output_name"] == "my_lora"
    assert meta["format"] == "pt" and meta["software"]["name"] and len(meta["sshs_legacy_hash"]) == 8
    assert ck.get_latest_save_path(root, "my_lora") == paths[2]
    assert ck.load_training_state_from_metadata(paths[2]) == (300, 1)
    # false positives: another run's `_LoRA` files share the prefix and must not be picked up (:843-851)
    final = ck.save_checkpoint(net, None, root, "my_lora", step=None, named_lora=True)
    assert os.path.basename(final) == "my_lora_LoRA.safetensors"
    assert ck.get_latest_save_path(root, "my_lora") == paths[2]
    assert ck.get_latest_save_path(root, "my_lora_LoRA") == final
    # nothing saved yet -> the pretrained LoRA is the start point, and its metadata is NOT a resume state (:868-872)
    empty = str(tmp_path / "empty")
    assert ck.get_latest_save_path(empty, "x") is None
    assert ck.get_latest_save_path(empty, "x", pretrained_lora_path=paths[2]) == paths[2]
    assert ck.load_training_state_from_metadata(paths[2], pretrained_lora_path=paths[2]) is None


def test_resume_restores_weights_step_and_optimizer_but_keeps_configured_lr(tmp_path):
    root = str(tmp_path / "run")
    _, net = _net(seed=1)
    opt = B200AdamW(net, lr=2e-4)
    opt.exp_avg.normal_()
    opt.exp_avg_sq.uniform_()
    opt.state_buf[0] = 41
    ck.save_checkpoint(net, opt, root, "r", step=41, epoch=2, dtype=torch.float32)
    _, net2 = _net(seed=2)
    opt2 = B200AdamW(net2, lr=5e-5)
    path, step, epoch = ck.resume(net2, opt2, root, "r")
    assert (step, epoch) == (41, 2) and path.endswith("r_000000041.safetensors")
    for a, b in zip(net.unet_loras, net2.unet_loras):
        assert torch.equal(a.lora_up.weight, b.lora_up.weight) and torch.equal(a.lora_down.weight, b.lora_down.weight)
    assert torch.equal(opt2.exp_avg, opt.exp_avg) and int(opt2.state_buf[0]) == 41
    assert opt2.param_groups[0]["lr"] == 5e-5 and abs(float(opt2.hyper[0]) - 5e-5) < 1e-11  # :2215-2218
    # a rank change on load invalidates the optimizer state: it must not be loaded (:2200-2205)
    _, net3 = _net(rank=8, seed=3)
    opt3 = B200AdamW(net3, lr=1e-4)
    _, step3, _ = ck.resume(net3, opt3, root, "r")
    assert net3.did_change_weights and step3 == 41 and float(opt3.exp_avg.abs().sum()) == 0.0


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_checkpoint_interchange_with_live_reference(tmp_path):
    from oracle import flux_ref

    ref_import.install()
    from toolkit.metadata import load_metadata_from_safetensors as ref_load_meta  # the reference's own reader

    RefNet, _ = ref_import.reference_lora()
    root = str(tmp_path / "x")
    _, net = _net(seed=5)
    opt = B200AdamW(net, lr=1e-4, eps=1e-6)
    opt.exp_avg.normal_()
    opt.exp_avg_sq.uniform_()
    opt.state_buf[0] = 9
    path = ck.save_checkpoint(net, opt, root, "job", step=9, dtype=torch.float32)
    # 1. the reference reads our metadata the way load_training_state_from_metadata does (:883-889)
    meta = ref_load_meta(path)
    assert meta["training_info"]["step"] == 9 and dict(meta) == dict(ck.load_metadata_from_safetensors(path))
    # 2. the reference network loads our file
    m1 = flux_ref.FluxTransformer2DModel(flux_ref.FluxConfig(**CFG))
    r = RefNet(unet=m1, lora_dim=4, **dict(KW, alpha=4))
    r.force_to("cpu", torch.float32); r._update_torch_multiplier(); r.apply_to(None, m1, False, True)
    r.load_weights(path)
    for a, b in zip(net.unet_loras, r.unet_loras):
        assert a.lora_name == b.lora_name
        assert torch.equal(a.lora_up.weight, b.lora_up.weight) and torch.equal(a.lora_down.weight, b.lora_down.weight)
    # 3. torch.optim.AdamW over the reference network's parameter groups loads our optimizer.pt (:2209-2211)
    ref_opt = torch.optim.AdamW(r.prepare_optimizer_params(1e-4, 1e-4, 1e-4), lr=1e-4, eps=1e-6)
    ref_opt.load_state_dict(torch.load(os.path.join(root, "optimizer.pt"), weights_only=True))
    p0 = ref_opt.param_groups[0]["params"][0]
    assert float(ref_opt.state[p0]["step"]) == 9.0
    assert torch.equal(ref_opt.state[p0]["exp_avg"].reshape(-1), opt.exp_avg[:p0.numel()])
    # 4. and back: a file written by the reference network resumes here
    r.save_weights(os.path.join(root, "job_000000010.safetensors"), dtype=torch.float32,
                   metadata=ck.get_meta_for_safetensors(ck.training_metadata("job", 10), "job"))
    time.sleep(0.02)
    _, net2 = _net(seed=6)
    p, step, _ = ck.resume(net2, None, root, "job")
    assert p.endswith("job_000000010.safetensors") and step == 10
    for a, b in zip(net2.unet_loras, r.unet_loras):
        assert torch.equal(a.lora_up.weight, b.lora_up.weight)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_metadata_and_hashes_identical_to_live_reference():
    """`get_meta_for_safetensors` flattening and the sshs hashes (toolkit/metadata.py:13-46, train_tools.py:162-185)
    computed by the reference's own functions on the same state dict."""
    import copy
    from collections import OrderedDict

    ref_import.install()
    from toolkit import metadata as ref_meta

    from ai_toolkit_b200 import metadata as my_meta

    torch.manual_seed(0)
    sd = OrderedDict((f"transformer.blocks.{i}.q.lora_{ab}.weight", torch.randn(*shape).half())
                     for i in range(3) for ab, shape in (("A", (4, 4096)), ("B", (4096, 4))))  # > 1 MiB: legacy hash window
    meta = OrderedDict(training_info=OrderedDict(step=7, epoch=1), ss_output_name="[name]_v1", ss_base_model_version="flux.1",
                       nested=dict(a=[1, 2], b="x"))
    want = ref_meta.get_meta_for_safetensors(copy.deepcopy(meta), name="job", add_software_info=False)
    got = my_meta.get_meta_for_safetensors(copy.deepcopy(meta), name="job", add_software_info=False)
    assert got == want and got["ss_output_name"] == "job_v1" and got["format"] == "pt"
    want = ref_meta.add_model_hash_to_meta(sd, copy.deepcopy(want))
    got = my_meta.add_model_hash_to_meta(sd, copy.deepcopy(got))
    assert got["sshs_model_hash"] == want["sshs_model_hash"] and got["sshs_legacy_hash"] == want["sshs_legacy_hash"]
    assert ck.parse_metadata_from_safetensors(got) == ref_meta.parse_metadata_from_safetensors(want)


def test_resume_resets_ema_and_reads_both_optimizer_layouts_and_overwrites_stale_training_info(tmp_path):
    """(1) EMA shadow after resume = the LOADED weights (the reference builds its EMA after load_weights,
    BaseSDTrainProcess.py:2053 then :2229); (2) an optimizer.pt written through B200AdamW.state_dict() (what the reference
    trainer's own save() calls) loads as well as the torch.optim.AdamW layout; (3) `meta` carrying a previous run's
    training_info is overwritten with the current step (`self.meta.update`, :388-409)."""
    root = str(tmp_path / "run")
    _, net = _net(seed=1)
    opt = B200AdamW(net, lr=2e-4, ema_decay=0.99)
    opt.exp_avg.normal_()
    opt.state_buf[0] = 7
    stale = {"training_info": {"step": 3, "epoch": 0}, "ss_output_name": "old"}
    path = ck.save_checkpoint(net, opt, root, "r", step=7, epoch=1, dtype=torch.float32, meta=stale)
    assert ck.load_training_state_from_metadata(path) == (7, 1)
    assert ck.load_metadata_from_safetensors(path)["ss_output_name"] == "r"
    _, net2 = _net(seed=2)
    opt2 = B200AdamW(net2, lr=1e-4, ema_decay=0.99)
    assert not torch.equal(opt2.ema, net.flat_params)
    ck.resume(net2, opt2, root, "r")
    assert torch.equal(net2.flat_params, net.flat_params)
    assert torch.equal(opt2.ema, net2.flat_params)
    assert torch.equal(opt2.exp_avg, opt.exp_avg) and int(opt2.state_buf[0]) == 7
    # flat layout on disk
    torch.save(opt.state_dict(), os.path.join(root, "optimizer.pt"))
    _, net3 = _net(seed=3)
    opt3 = B200AdamW(net3, lr=1e-4, ema_decay=0.99)
    ck.resume(net3, opt3, root, "r")
    assert torch.equal(opt3.exp_avg, opt.exp_avg) and int(opt3.state_buf[0]) == 7
    assert opt3.param_groups[0]["lr"] == 1e-4


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_weight_mapped_keymap_identical_to_live_reference(monkeypatch):
    """`get_keymap(force_weight_mapping=True)` / the ssd and vega paths (toolkit/network_mixins.py:524-566,
    toolkit/saving.py:279-330) on the reference's own key-map data files."""
    import types

    ref_import.install()
    from toolkit.saving import get_lora_keymap_from_model_keymap
    import toolkit
    from ai_toolkit_b200 import keymaps

    monkeypatch.setenv("AITK_KEYMAPS_ROOT", os.path.join(os.path.dirname(toolkit.__file__), "keymaps"))
    for flags in (dict(is_sdxl=True), dict(), dict(is_ssd=True), dict(is_vega=True)):
        net = types.SimpleNamespace(is_ssd=False, is_vega=False, is_sdxl=False, is_v2=False)
        net.__dict__.update(flags)
        tail = "ssd" if net.is_ssd else "vega" if net.is_vega else "sdxl" if net.is_sdxl else "sd1"
        import json
        with open(os.path.join(os.environ["AITK_KEYMAPS_ROOT"], f"stable_diffusion_{tail}.json")) as f:
            want = get_lora_keymap_from_model_keymap(json.load(f)["ldm_diffusers_keymap"])
        got = keymaps.load_keymap(net, force_weight_mapping=True)
        assert got is not None and list(got.items()) == list(want.items()) and len(got) > 100
        if not (net.is_ssd or net.is_vega):
            assert keymaps.load_keymap(net) is None  # no stable_diffusion_locon_<tail>.json in the reference tree
