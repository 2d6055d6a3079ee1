"""Host-side mirror of the reference's network API (no GPU): names, state dict, peft save format, load with
rank expand / shrink, merge_in / merge_out, multiplier / is_active semantics, optimizer param groups, and the
'no CPU fallback' rule for an active network."""
import os

import pytest
import torch

from ai_toolkit_b200 import LoRASpecialNetwork, cabi, get_network
from ai_toolkit_b200.flux import FluxConfig, FluxTransformer2DModel
from oracle import ref_import

GOLD = os.path.join(os.path.dirname(__file__), "golden", "lora_tiny.pt")
CFG = dict(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=64, pooled_projection_dim=32)


def _net(rank=4, **kw):
    model = FluxTransformer2DModel(FluxConfig(**CFG), dtype=torch.float32)
    net = LoRASpecialNetwork(text_encoder=None, unet=model, lora_dim=rank, alpha=rank, train_unet=True, train_text_encoder=False,
                             is_flux=True, transformer_only=True, **kw)
    net.force_to("cpu", torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, model, False, True)
    return model, net


def test_names_and_saved_keys_match_reference_golden():
    gold = torch.load(GOLD, weights_only=False)
    model, net = _net()
    assert [l.lora_name for l in net.unet_loras] == gold["lora_names"]
    assert list(net.state_dict().keys()) == list(gold["init_state_dict"].keys())
    assert list(net.get_state_dict(dtype=torch.float16).keys()) == gold["saved_keys"]
    assert len(net.unet_loras) == 14 + 6  # every Linear under transformer_blocks / single_transformer_blocks
    assert net.peft_format and all(l.scale == 1.0 for l in net.unet_loras)


def test_flat_buffers_are_views():
    model, net = _net()
    n = sum(p.numel() for p in net.parameters())
    assert net.n_params == n
    w = net.unet_loras[3].lora_down.weight
    w.data.fill_(0.25)
    assert (net.flat_params == 0.25).sum() == w.numel()
    assert w.grad.data_ptr() >= net.flat_grads.data_ptr()
    groups = net.prepare_optimizer_params(None, 1e-4, 1e-4)
    assert len(groups) == 1 and groups[0]["lr"] == 1e-4 and len(groups[0]["params"]) == 2 * len(net.unet_loras)


def test_save_load_roundtrip_and_rank_change(tmp_path):
    model, net = _net(rank=4)
    with torch.no_grad():
        for l in net.unet_loras:
            l.lora_up.weight.normal_(0, 0.1)
    f = str(tmp_path / "lora.safetensors")
    net.save_weights(f, dtype=torch.float32, metadata={"ss_output_name": "t"})
    from safetensors import safe_open

    with safe_open(f, "pt") as sf:
        keys = list(sf.keys())
        meta = sf.metadata()
    assert all(k.startswith("transformer.") and (".lora_A.weight" in k or ".lora_B.weight" in k) for k in keys)
    assert "sshs_model_hash" in meta and len(meta["sshs_legacy_hash"]) == 8
    model2, net2 = _net(rank=4)
    assert net2.load_weights(f) is None
    for a, b in zip(net.unet_loras, net2.unet_loras):
        assert torch.equal(a.lora_up.weight, b.lora_up.weight) and torch.equal(a.lora_down.weight, b.lora_down.weight)
    assert net2.flat_params.data_ptr() == net2.unet_loras[0].lora_down.weight.data_ptr()  # still views after loading
    model3, net3 = _net(rank=8)  # expand (network_mixins.py:737-775)
    net3.load_weights(f)
    assert net3.did_change_weights
    a, c = net.unet_loras[2], net3.unet_loras[2]
    assert torch.equal(c.lora_down.weight[:4], a.lora_down.weight) and c.lora_down.weight[4:].abs().sum() == 0
    assert torch.equal(c.lora_up.weight[:, :4], a.lora_up.weight) and c.lora_up.weight[:, 4:].abs().sum() == 0


def test_merge_in_out_and_multiplier():
    model, net = _net()
    lin = model.transformer_blocks[0].attn.to_q
    w0 = lin.weight.detach().clone()
    lora = lin._b200_lora()
    with torch.no_grad():
        lora.lora_up.weight.normal_(0, 0.1)
    net.merge_in(0.5)
    want = w0 + 0.5 * lora.scale * (lora.lora_up.weight @ lora.lora_down.weight)
    torch.testing.assert_close(lin.weight, want)
    assert net.is_merged_in and not lora.is_live()
    net.merge_out(0.5)
    torch.testing.assert_close(lin.weight, w0, atol=1e-6, rtol=1e-5)
    net.multiplier = [1.0, -1.0]
    assert net.torch_multiplier.tolist() == [1.0, -1.0]
    net.multiplier = 0
    with net:
        assert not lora.is_live()
    net.multiplier = 1.0
    assert not lora.is_live()  # inactive outside `with network:`
    x = torch.randn(3, 5, lin.in_features)
    assert torch.equal(lin(x), torch.nn.functional.linear(x, lin.weight, lin.bias))  # falls through to org_forward


def test_active_network_has_no_cpu_fallback():
    model, net = _net()
    lin = model.transformer_blocks[0].attn.to_q
    with net, pytest.raises(cabi.B200Error):
        lin(torch.randn(4, lin.in_features))


def test_get_network_factory_defaults():
    class NC:
        linear, linear_alpha, type, transformer_only, conv, conv_alpha, dropout, network_kwargs = 16, 16, "lora", True, None, None, None, {}

    class MC:
        is_flux = True

    model = FluxTransformer2DModel(FluxConfig(**CFG), dtype=torch.float32)
    net = get_network(model, network_config=NC(), model_config=MC(), device="cpu")
    assert net.lora_dim == 16 and len(net.unet_loras) == 20 and net.flat_params is not None
    with pytest.raises(NotImplementedError):
        LoRASpecialNetwork(None, model, network_type="dora", is_flux=True)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_state_dict_identical_to_live_reference():
    from oracle import flux_ref

    RefNet, _ = ref_import.reference_lora()
    m1 = flux_ref.FluxTransformer2DModel(flux_ref.FluxConfig(**CFG))
    m2 = FluxTransformer2DModel(FluxConfig(**CFG), dtype=torch.float32)
    kw = dict(text_encoder=None, lora_dim=4, alpha=4, train_unet=True, train_text_encoder=False, is_flux=True,
              network_type="lora", transformer_only=True)
    torch.manual_seed(1)
    r = RefNet(unet=m1, **kw)
    r.force_to("cpu", torch.float32); r._update_torch_multiplier(); r.apply_to(None, m1, False, True)
    torch.manual_seed(1)
    n = LoRASpecialNetwork(unet=m2, **kw)
    n.force_to("cpu", torch.float32); n._update_torch_multiplier(); n.apply_to(None, m2, False, True)
    sr, sn = r.state_dict(), n.state_dict()
    assert list(sr.keys()) == list(sn.keys())
    for k in sr:
        assert torch.equal(sr[k], sn[k]), k
    assert list(r.get_state_dict(dtype=torch.float16).keys()) == list(n.get_state_dict(dtype=torch.float16).keys())


def test_optimizer_state_interchange_with_torch_adamw():
    """B200AdamW state <-> torch.optim.AdamW state dict layout (what the reference stores in optimizer.pt)."""
    from ai_toolkit_b200.optimizer import B200AdamW
    model, net = _net()
    opt = B200AdamW(net, lr=3e-4, eps=1e-6, weight_decay=1e-2)
    opt.exp_avg.normal_()
    opt.exp_avg_sq.uniform_()
    opt.state_buf[0] = 7
    sd = opt.torch_state_dict()
    ref = torch.optim.AdamW(net.prepare_optimizer_params(None, 3e-4, 3e-4), lr=3e-4, eps=1e-6)
    ref.load_state_dict(sd)  # torch accepts it
    some = next(iter(ref.state.values()))
    assert float(some["step"]) == 7.0 and ref.param_groups[0]["eps"] == 1e-6
    opt2 = B200AdamW(net, lr=1e-4)
    opt2.load_torch_state_dict(ref.state_dict())
    assert torch.equal(opt2.exp_avg, opt.exp_avg) and torch.equal(opt2.exp_avg_sq, opt.exp_avg_sq)
    assert int(opt2.state_buf[0]) == 7 and opt2.param_groups[0]["lr"] == 3e-4


def test_flux_container_matches_oracle_parameter_names():
    """The parameter container has exactly the oracle's (= diffusers') parameter names and shapes, so checkpoints load."""
    from oracle import flux_ref

    cfg = dict(CFG)
    o = flux_ref.FluxTransformer2DModel(flux_ref.FluxConfig(**cfg))
    m = FluxTransformer2DModel(FluxConfig(**cfg), dtype=torch.float32)
    so, sm = o.state_dict(), m.state_dict()
    assert set(so.keys()) == set(sm.keys())
    for k in so:
        assert so[k].shape == sm[k].shape, k
    m.load_state_dict(so, strict=True)


def test_to_rebuilds_flat_views():
    model, net = _net()
    w_before = net.unet_loras[0].lora_down.weight.detach().clone()
    net.to(torch.float32)  # a no-op move still goes through _apply -> _flatten
    assert torch.equal(net.unet_loras[0].lora_down.weight, w_before)
    assert net.unet_loras[0].lora_down.weight.data_ptr() == net.flat_params.data_ptr()


class Transformer2DModel(torch.nn.Module):  # the class name the reference targets inside a UNet (kohya_lora.py:750)
    def __init__(self):
        super().__init__()
        self.proj_in = torch.nn.Conv2d(8, 16, 1)
        self.to_q = torch.nn.Linear(16, 16, bias=False)
        self.ff = torch.nn.Linear(16, 32)
        self.conv3 = torch.nn.Conv2d(16, 16, 3, padding=1)


def _toy_unet():
    cls = type("UNet2DConditionModel", (torch.nn.Module,), {})

    def init(self):
        torch.nn.Module.__init__(self)
        self.down_blocks = torch.nn.ModuleList([Transformer2DModel(), Transformer2DModel()])
        self.conv_in = torch.nn.Conv2d(4, 8, 3, padding=1)      # outside Transformer2DModel: never wrapped
        self.time_proj = torch.nn.Linear(8, 8)

    cls.__init__ = init
    return cls()


def test_kohya_format_unet_naming_alpha_and_scale():
    """SD1.5 / SDXL path (BASELINE configs[0], [1], plumbing): kohya names `lora_unet_<path_with_underscores>`, alpha saved,
    scale = alpha / rank, Linear + 1x1 Conv wrapped, 3x3 Conv skipped when conv_lora_dim is None (lora_special.py:456-640)."""
    unet = _toy_unet()
    net = LoRASpecialNetwork(text_encoder=None, unet=unet, lora_dim=4, alpha=2, train_unet=True, train_text_encoder=False)
    net.force_to("cpu", torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, unet, False, True)
    names = [l.lora_name for l in net.unet_loras]
    assert names == ["lora_unet_down_blocks_0_proj_in", "lora_unet_down_blocks_0_to_q", "lora_unet_down_blocks_0_ff",
                     "lora_unet_down_blocks_1_proj_in", "lora_unet_down_blocks_1_to_q", "lora_unet_down_blocks_1_ff"]
    assert not net.peft_format and all(abs(l.scale - 0.5) < 1e-9 for l in net.unet_loras)
    sd = net.get_state_dict(dtype=torch.float32)
    assert "lora_unet_down_blocks_0_to_q.alpha" in sd and float(sd["lora_unet_down_blocks_0_to_q.alpha"]) == 2.0
    assert sd["lora_unet_down_blocks_0_proj_in.lora_down.weight"].shape == (4, 8, 1, 1)
    assert sd["lora_unet_down_blocks_0_ff.lora_up.weight"].shape == (32, 4)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_kohya_format_identical_to_live_reference():
    RefNet, _ = ref_import.reference_lora()
    u1, u2 = _toy_unet(), _toy_unet()
    kw = dict(text_encoder=None, lora_dim=4, alpha=2, train_unet=True, train_text_encoder=False)
    torch.manual_seed(2)
    r = RefNet(unet=u1, **kw)
    r.force_to("cpu", torch.float32); r._update_torch_multiplier(); r.apply_to(None, u1, False, True)
    torch.manual_seed(2)
    n = LoRASpecialNetwork(unet=u2, **kw)
    n.force_to("cpu", torch.float32); n._update_torch_multiplier(); n.apply_to(None, u2, False, True)
    assert [l.lora_name for l in r.unet_loras] == [l.lora_name for l in n.unet_loras]
    sr, sn = r.get_state_dict(dtype=torch.float32), n.get_state_dict(dtype=torch.float32)
    assert list(sr.keys()) == list(sn.keys())
    for k in sr:
        assert torch.equal(sr[k], sn[k]), k
    assert [l.scale for l in r.unet_loras] == [l.scale for l in n.unet_loras]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_rank_change_merge_and_multiplier_identical_to_live_reference(tmp_path):
    """Numbers, not just shapes: a rank-4 file loaded into rank-8 / rank-2 networks, merge_in / merge_out of the frozen
    weights, and the multiplier tensor, next to the UNMODIFIED reference network doing the same
    (toolkit/network_mixins.py:737-775, :370-462, :790-846)."""
    from oracle import flux_ref

    RefNet, _ = ref_import.reference_lora()
    kw = dict(text_encoder=None, train_unet=True, train_text_encoder=False, is_flux=True, network_type="lora",
              transformer_only=True)

    def pair(rank, seed):
        torch.manual_seed(seed)
        m1 = flux_ref.FluxTransformer2DModel(flux_ref.FluxConfig(**CFG))
        m2 = FluxTransformer2DModel(FluxConfig(**CFG), dtype=torch.float32)
        m2.load_state_dict(m1.state_dict())
        torch.manual_seed(seed + 1)
        r = RefNet(unet=m1, lora_dim=rank, alpha=rank, **kw)
        r.force_to("cpu", torch.float32); r._update_torch_multiplier(); r.apply_to(None, m1, False, True)
        torch.manual_seed(seed + 1)
        n = LoRASpecialNetwork(unet=m2, lora_dim=rank, alpha=rank, **kw)
        n.force_to("cpu", torch.float32); n._update_torch_multiplier(); n.apply_to(None, m2, False, True)
        return m1, r, m2, n

    _, r4, _, n4 = pair(4, 10)
    with torch.no_grad():
        for a, b in zip(r4.unet_loras, n4.unet_loras):
            a.lora_up.weight.normal_(0, 0.1)
            b.lora_up.weight.copy_(a.lora_up.weight)
    f = str(tmp_path / "r4.safetensors")
    r4.save_weights(f, dtype=torch.float32)
    for rank in (8, 2):  # expand with zero rows / columns, shrink by truncation
        m1, r, m2, n = pair(rank, 20 + rank)
        r.load_weights(f)
        n.load_weights(f)
        assert r.did_change_weights and n.did_change_weights
        for a, b in zip(r.unet_loras, n.unet_loras):
            assert torch.equal(a.lora_down.weight, b.lora_down.weight) and torch.equal(a.lora_up.weight, b.lora_up.weight)
        # merge into the frozen weights and back out
        for mw in (1.0, 0.35):
            r.merge_in(mw)
            n.merge_in(mw)
            assert r.is_merged_in and n.is_merged_in
            sd1, sd2 = m1.state_dict(), m2.state_dict()
            for k in sd1:
                assert torch.equal(sd1[k], sd2[k]), (rank, mw, k)
            r.merge_out(mw)
            n.merge_out(mw)
            sd1, sd2 = m1.state_dict(), m2.state_dict()
            for k in sd1:
                assert torch.equal(sd1[k], sd2[k]), (rank, mw, k)
        # multiplier -> tensor
        for val in (0.5, [1.0, -1.0, 0.25], torch.tensor([2.0, 0.0])):
            r.multiplier = val
            n.multiplier = val
            assert torch.equal(r.torch_multiplier, n.torch_multiplier) and r.torch_multiplier.dtype == n.torch_multiplier.dtype


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_text_encoder_and_unet_construction_identical_to_live_reference():
    """CLIP text encoder(s) + UNet with DEFAULT target lists (the reference takes kohya's: CLIPAttention / CLIPMLP,
    Transformer2DModel): adapter names (`lora_te_`, `lora_te1_/te2_`, `lora_unet_`), kaiming draws, scale = alpha / rank,
    optimizer parameter groups with their learning rates (toolkit/lora_special.py:346-470, :644-700)."""
    from transformers import CLIPTextConfig, CLIPTextModel

    RefNet, _ = ref_import.reference_lora()
    cfg = CLIPTextConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=2, num_attention_heads=2, vocab_size=100,
                         max_position_embeddings=16)
    kw = dict(lora_dim=4, alpha=2, train_unet=True, train_text_encoder=True)

    def build(cls, n_te):
        torch.manual_seed(0)
        tes = [CLIPTextModel(cfg) for _ in range(n_te)]
        unet = _toy_unet()
        torch.manual_seed(1)
        return cls(text_encoder=tes[0] if n_te == 1 else tes, unet=unet, **kw)

    for n_te in (1, 2):
        r, n = build(RefNet, n_te), build(LoRASpecialNetwork, n_te)
        ra, na = r.text_encoder_loras + r.unet_loras, n.text_encoder_loras + n.unet_loras
        assert [l.lora_name for l in ra] == [l.lora_name for l in na] and len(r.text_encoder_loras) == 12 * n_te
        assert na[0].lora_name.startswith("lora_te_" if n_te == 1 else "lora_te1_")
        for a, b in zip(ra, na):
            assert torch.equal(a.lora_down.weight, b.lora_down.weight) and a.scale == b.scale == 0.5
        g1, g2 = r.prepare_optimizer_params(1e-5, 2e-4, 1e-4), n.prepare_optimizer_params(1e-5, 2e-4, 1e-4)
        assert [(g["lr"], len(g["params"])) for g in g1] == [(g["lr"], len(g["params"])) for g in g2]


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("opts", [
    dict(transformer_only=True), dict(transformer_only=False), dict(transformer_only=True, attn_only=True),
    dict(transformer_only=True, only_if_contains=["transformer_blocks.1.", "single_transformer_blocks.0."]),
    dict(transformer_only=True, ignore_if_contains=["norm", "ff"]), dict(transformer_only=True, parameter_threshold=5000.0),
    dict(transformer_only=False, only_if_contains=["x_embedder", "proj_out"]),
    dict(transformer_only=True, target_lin_modules=["FluxTransformerBlock"]), dict(transformer_only=True, is_transformer=True),
    dict(transformer_only=True, peft_format=False)], ids=lambda o: ",".join(f"{k}" for k in o))
def test_module_selection_options_identical_to_live_reference(opts):
    """Which Linear layers get an adapter, in which order, under the selection options of the constructor
    (toolkit/lora_special.py:346-470): identical name lists, format flag and scales next to the live reference."""
    from oracle import flux_ref

    RefNet, _ = ref_import.reference_lora()
    cfg = dict(CFG, num_layers=2, num_single_layers=2)
    base = dict(text_encoder=None, lora_dim=4, alpha=4, train_unet=True, train_text_encoder=False, is_flux=True,
                network_type="lora")
    torch.manual_seed(0)
    r = RefNet(unet=flux_ref.FluxTransformer2DModel(flux_ref.FluxConfig(**cfg)), **base, **opts)
    torch.manual_seed(0)
    n = LoRASpecialNetwork(unet=FluxTransformer2DModel(FluxConfig(**cfg), dtype=torch.float32), **base, **opts)
    assert [l.lora_name for l in r.unet_loras] == [l.lora_name for l in n.unet_loras]
    assert r.peft_format == n.peft_format and [float(l.scale) for l in r.unet_loras] == [float(l.scale) for l in n.unet_loras]


def test_lr_scheduler_drives_the_device_hyper_buffer():
    """`lr_scheduler.step()` after every optimizer step (SDTrainer.py:2299-2301): torch schedulers write
    `param_groups[i]['lr']`; the fused optimizer reads its hyper-parameters from a device buffer, which `sync_hyper()`
    (called by `step()` and by `FluxLoRATrainStep.run`) refreshes only when a value changed."""
    from ai_toolkit_b200.optimizer import B200AdamW
    model, net = _net()
    opt = B200AdamW(net, lr=1e-3, eps=1e-6, weight_decay=0.0)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda step: 0.5 ** step)
    assert abs(float(opt.hyper[0]) - 1e-3) < 1e-9
    for k in range(1, 4):
        opt._step_count = getattr(opt, "_step_count", 0) + 1  # (silences torch's "scheduler before optimizer" warning)
        sched.step()
        before = opt.hyper.data_ptr()
        opt.sync_hyper()
        assert abs(float(opt.hyper[0]) - 1e-3 * 0.5 ** k) < 1e-9 and opt.hyper.data_ptr() == before  # same buffer (CUDA graphs)
    host = list(opt._hyper_host)
    opt.sync_hyper()
    assert opt._hyper_host == host


class ResnetBlock2D(torch.nn.Module):  # kohya_lora.py:751: the 3x3-conv targets
    def __init__(self):
        super().__init__()
        self.conv1 = torch.nn.Conv2d(8, 16, 3, padding=1)
        self.conv2 = torch.nn.Conv2d(16, 16, 3, stride=2, padding=1)
        self.conv_shortcut = torch.nn.Conv2d(8, 16, 1)


def _toy_unet_conv():
    cls = type("UNet2DConditionModel", (torch.nn.Module,), {})

    def init(self):
        torch.nn.Module.__init__(self)
        self.down_blocks = torch.nn.ModuleList([Transformer2DModel(), ResnetBlock2D()])

    cls.__init__ = init
    return cls()


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_conv3x3_lora_extract_and_merge_identical_to_live_reference():
    """conv_lora_dim (toolkit/lora_special.py:95-104, :585-587, :679-681): k x k adapters on the ResNet convs -- names,
    shapes (down = Conv k x k with the wrapped conv's stride / padding, up = 1x1), kaiming draws, saved keys, merge_in and
    extract_weight next to the UNMODIFIED reference."""
    RefNet, _ = ref_import.reference_lora()
    u1, u2 = _toy_unet_conv(), _toy_unet_conv()
    u2.load_state_dict(u1.state_dict())
    kw = dict(text_encoder=None, lora_dim=4, alpha=2, conv_lora_dim=4, conv_alpha=1, train_unet=True, train_text_encoder=False,
              target_lin_modules=["Transformer2DModel"], target_conv_modules=["ResnetBlock2D"])
    torch.manual_seed(5)
    r = RefNet(unet=u1, **dict(kw, target_lin_modules=["Transformer2DModel"]))
    r.force_to("cpu", torch.float32); r._update_torch_multiplier(); r.apply_to(None, u1, False, True)
    torch.manual_seed(5)
    n = LoRASpecialNetwork(unet=u2, **dict(kw, target_lin_modules=["Transformer2DModel"]))
    n.force_to("cpu", torch.float32); n._update_torch_multiplier(); n.apply_to(None, u2, False, True)
    assert [l.lora_name for l in r.unet_loras] == [l.lora_name for l in n.unet_loras]
    assert any("conv2" in l.lora_name for l in n.unet_loras)
    sr, sn = r.get_state_dict(dtype=torch.float32), n.get_state_dict(dtype=torch.float32)
    assert list(sr.keys()) == list(sn.keys())
    for k in sr:
        assert sr[k].shape == sn[k].shape and torch.equal(sr[k], sn[k]), k
    assert [l.scale for l in r.unet_loras] == [l.scale for l in n.unet_loras]
    c2r = [l for l in r.unet_loras if l.lora_name.endswith("conv2")][0]
    c2n = [l for l in n.unet_loras if l.lora_name.endswith("conv2")][0]
    assert tuple(c2n.lora_down.weight.shape) == tuple(c2r.lora_down.weight.shape) == (4, 16, 3, 3)
    assert c2n.lora_down.stride == c2r.lora_down.stride == (2, 2) and c2n.lora_down.padding == c2r.lora_down.padding
    # merge_in on a conv adapter
    with torch.no_grad():
        for a, b in zip(r.unet_loras, n.unet_loras):
            a.lora_up.weight.normal_(0, 0.1)
            b.lora_up.weight.copy_(a.lora_up.weight)
    r.merge_in(0.7); n.merge_in(0.7)
    for (k1, p1), (k2, p2) in zip(u1.state_dict().items(), u2.state_dict().items()):
        torch.testing.assert_close(p1, p2, rtol=1e-6, atol=1e-6, msg=k1)
    r.merge_out(0.7); n.merge_out(0.7)
    # extract_weight: truncated SVD of the wrapped layer (linear and conv), rank / alpha / runtime scale updates
    for a, b in zip(r.unet_loras, n.unet_loras):
        a.extract_weight(extract_mode="fixed", extract_mode_param=2)
        b.extract_weight(extract_mode="fixed", extract_mode_param=2)
        assert a.lora_dim == b.lora_dim and a.scale == b.scale == 1.0 and float(b._runtime_scale) == 1.0
        assert tuple(a.lora_down.weight.shape) == tuple(b.lora_down.weight.shape), a.lora_name
        # SVD factors are unique up to sign per component: compare the product up @ down
        pa = a.lora_up.weight.flatten(1) @ a.lora_down.weight.flatten(1)
        pb = b.lora_up.weight.flatten(1) @ b.lora_down.weight.flatten(1)
        torch.testing.assert_close(pa, pb, rtol=1e-4, atol=1e-5)
    assert n.flat_params.data_ptr() == n.unet_loras[0].lora_down.weight.data_ptr()  # flat views rebuilt after the rank change


def test_module_and_rank_dropout_flags():
    """dropout / rank_dropout / module_dropout are accepted (reference default None); masks are only drawn in training mode
    (toolkit/network_mixins.py:197-226); an inactive / eval network never consults them."""
    unet = _toy_unet()
    net = LoRASpecialNetwork(text_encoder=None, unet=unet, lora_dim=4, alpha=2, train_unet=True, train_text_encoder=False,
                             dropout=0.1, rank_dropout=0.2, module_dropout=1.0)
    net.force_to("cpu", torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, unet, False, True)
    lora = net.unet_loras[1]
    net.train()
    assert lora.has_dropout()
    lin = lora.org_module[0]
    x = torch.randn(2, 3, 16)
    with net:  # module_dropout = 1.0 -> the adapter contributes 0.0: the frozen layer's output, even on CPU
        assert torch.equal(lin(x), torch.nn.functional.linear(x, lin.weight))
    net.eval()
    assert not lora.has_dropout()
