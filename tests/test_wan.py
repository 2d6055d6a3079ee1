"""Wan2.1 (BASELINE.json configs[3]).  CPU: the oracle's attention against the reference's IN-TREE attention processor
(toolkit/models/wan21/wan_attn.py, unmodified), adapter naming / saved keys against the live reference network, the
original-name key conversion against toolkit/models/wan21/wan_lora_convert.py, container == oracle parameter names.
GPU: norm / rope kernels vs torch, and the engine step (forward, loss, backward) vs the oracle in fp32 and bf16."""
import copy

import pytest
import torch

from ai_toolkit_b200 import LoRASpecialNetwork, wan_keys
from ai_toolkit_b200.wan import WanConfig, WanTransformer3DModel
from oracle import lora_ref, ref_import, wan_ref

TOY = dict(num_attention_heads=2, text_dim=64, ffn_dim=512, num_layers=2)
DEV = "cuda:0"


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_container_matches_oracle_parameter_names():
    o = wan_ref.WanTransformer3DModel(wan_ref.WanConfig(**TOY))
    m = WanTransformer3DModel(WanConfig(**TOY), dtype=torch.float32)
    so, sm = o.state_dict(), m.state_dict()
    assert set(so.keys()) == set(sm.keys())
    for k in so:
        assert so[k].shape == sm[k].shape, k
    m.load_state_dict(so, strict=True)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_oracle_attention_matches_in_tree_processor():
    """The reference's own `WanAttnProcessor2_0.__call__` (unmodified) on the oracle's attention module: self-attention with
    RoPE and cross-attention, fp32."""
    ref_import.install()
    from toolkit.models.wan21.wan_attn import WanAttnProcessor2_0

    torch.manual_seed(0)
    cfg = wan_ref.WanConfig(**TOY)
    attn = wan_ref.WanAttention(cfg.inner_dim, cfg.num_attention_heads, cfg.eps)
    wan_ref.init_synthetic_(attn, seed=1, std=0.05)
    proc = WanAttnProcessor2_0()
    x = torch.randn(2, 48, cfg.inner_dim)
    enc = torch.randn(2, 24, cfg.inner_dim)
    rot = wan_ref.rope_freqs(cfg, 3, 4, 4, "cpu")
    torch.testing.assert_close(attn(x, rotary_emb=rot), proc(attn, x, rotary_emb=rot))
    torch.testing.assert_close(attn(x, encoder_hidden_states=enc), proc(attn, x, encoder_hidden_states=enc))


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_adapter_names_and_saved_keys_identical_to_live_reference():
    RefNet, _ = ref_import.reference_lora()
    ref_import.install()
    from toolkit.models.wan21 import wan_lora_convert as wlc

    o = wan_ref.WanTransformer3DModel(wan_ref.WanConfig(**TOY))
    m = WanTransformer3DModel(WanConfig(**TOY), dtype=torch.float32)
    base = wan_keys.WanLoRABaseModel()
    kw = dict(text_encoder=None, lora_dim=4, alpha=4, train_unet=True, train_text_encoder=False, network_type="lora",
              transformer_only=True, is_transformer=True, target_lin_modules=["WanTransformer3DModel"], base_model=base)
    torch.manual_seed(3)
    r = RefNet(unet=o, **kw)
    r.force_to("cpu", torch.float32); r._update_torch_multiplier(); r.apply_to(None, o, False, True)
    torch.manual_seed(3)
    n = LoRASpecialNetwork(unet=m, **kw)
    n.force_to("cpu", torch.float32); n._update_torch_multiplier(); n.apply_to(None, m, False, True)
    assert [l.lora_name for l in r.unet_loras] == [l.lora_name for l in n.unet_loras]
    assert len(n.unet_loras) == 10 * TOY["num_layers"]  # attn1 q k v o, attn2 q k v o, ffn.0, ffn.2 (300 at 30 blocks)
    sr, sn = r.get_state_dict(dtype=torch.float32), n.get_state_dict(dtype=torch.float32)
    assert list(sr.keys()) == list(sn.keys())
    assert "diffusion_model.blocks.0.self_attn.q.lora_A.weight" in sn and "diffusion_model.blocks.1.ffn.2.lora_B.weight" in sn
    for k in sr:
        assert torch.equal(sr[k], sn[k]), k
    # the two converters themselves
    assert list(wlc.convert_to_diffusers(sn).keys()) == list(wan_keys.convert_to_diffusers(sn).keys())
    back = wan_keys.convert_to_diffusers(sn)
    assert list(wlc.convert_to_original(back).keys()) == list(wan_keys.convert_to_original(back).keys()) == list(sn.keys())


def test_save_load_roundtrip_through_original_names(tmp_path):
    m = WanTransformer3DModel(WanConfig(**TOY), dtype=torch.float32)
    kw = dict(text_encoder=None, lora_dim=4, alpha=4, train_unet=True, train_text_encoder=False, transformer_only=True,
              is_transformer=True, target_lin_modules=["WanTransformer3DModel"])
    base = wan_keys.WanLoRABaseModel()
    n = LoRASpecialNetwork(unet=m, base_model=base, **kw)
    n.force_to("cpu", torch.float32); n._update_torch_multiplier(); n.apply_to(None, m, False, True)
    with torch.no_grad():
        for l in n.unet_loras:
            l.lora_up.weight.normal_(0, 0.1)
    f = str(tmp_path / "wan_lora.safetensors")
    n.save_weights(f, dtype=torch.float32, metadata={})
    from safetensors import safe_open
    with safe_open(f, "pt") as sf:
        assert all(k.startswith("diffusion_model.blocks.") for k in sf.keys())
    m2 = WanTransformer3DModel(WanConfig(**TOY), dtype=torch.float32)
    n2 = LoRASpecialNetwork(unet=m2, base_model=base, **kw)
    n2.force_to("cpu", torch.float32); n2._update_torch_multiplier(); n2.apply_to(None, m2, False, True)
    n2.load_weights(f)
    for a, b in zip(n.unet_loras, n2.unet_loras):
        assert torch.equal(a.lora_up.weight, b.lora_up.weight) and torch.equal(a.lora_down.weight, b.lora_down.weight)


# ------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("rope", [True, False])
def test_rms_rope_kernels_vs_torch(rope):
    from ai_toolkit_b200 import ops
    torch.manual_seed(0)
    B, L, H = 2, 50, 3
    D = H * 128
    cfg = wan_ref.WanConfig(num_attention_heads=H)
    buf = torch.randn(B * L, 2 * D, device=DEV).bfloat16()  # the tensor is a column slice of a wider buffer
    x = buf[:, D:]
    w = (1 + 0.1 * torch.randn(D, device=DEV)).bfloat16()
    norm = wan_ref.RMSNorm(D, 1e-6).to(DEV, torch.bfloat16)
    norm.weight.data.copy_(w)
    out = torch.empty(B, H, L, 128, device=DEV, dtype=torch.bfloat16)
    cos = sin = rot = None
    if rope:
        rot = wan_ref.rope_freqs(cfg, 2, 5, 5, DEV)  # 2 * 5 * 5 = 50 positions
        ang = torch.angle(rot[0, 0])
        cos, sin = ang.cos().repeat_interleave(2, 1).float().contiguous(), ang.sin().repeat_interleave(2, 1).float().contiguous()
    rstd = ops.rms_rope_fwd(x, w, cos, sin, out, B, L)
    xr = x.detach().clone().float().requires_grad_(True)
    y = norm(xr.view(B, L, D).to(torch.bfloat16)).unflatten(2, (H, -1)).transpose(1, 2)
    if rope:
        y = wan_ref.apply_rotary_emb(y, rot)
    assert _rel(out, y) < 4e-3
    # backward against autograd of the fp32 formula
    xf = x.detach().clone().float().requires_grad_(True)
    yf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()
    yf = yf.view(B, L, H, 128).transpose(1, 2)
    if rope:
        yf = torch.view_as_real(torch.view_as_complex(yf.double().unflatten(3, (-1, 2)).contiguous()) * rot).flatten(3, 4).float()
    dY = torch.randn(B, H, L, 128, device=DEV).bfloat16()
    yf.backward(dY.float())
    dx = torch.zeros(B * L, 2 * D, device=DEV, dtype=torch.bfloat16)
    ops.rms_rope_bwd(dY, x, w, cos, sin, rstd, dx[:, :D], B, L)
    assert _rel(dx[:, :D], xf.grad) < 6e-3 and float(dx[:, D:].abs().sum()) == 0.0
    # mode 0: plain re-layout and back
    ops.rms_rope_fwd(x, None, None, None, out, B, L, mode=0)
    assert torch.equal(out, x.view(B, L, H, 128).transpose(1, 2))
    back = torch.empty(B * L, D, device=DEV, dtype=torch.bfloat16)
    ops.rms_rope_bwd(out, None, None, None, None, None, back, B, L, mode=0)
    assert torch.equal(back, x)


def _setup_wan(B, Fr, Hh, Ww, Lt, rank, seed=0):
    torch.manual_seed(seed)
    ocfg = wan_ref.WanConfig(**TOY)
    omodel = wan_ref.init_synthetic_(wan_ref.WanTransformer3DModel(ocfg), seed=seed, std=0.05).requires_grad_(False)
    model = WanTransformer3DModel(WanConfig(**TOY), device=DEV)
    model.load_state_dict(omodel.state_dict(), strict=True)
    net = LoRASpecialNetwork(None, model, lora_dim=rank, alpha=rank, train_text_encoder=False, transformer_only=True,
                             is_transformer=True, target_lin_modules=["WanTransformer3DModel"], base_model=wan_keys.WanLoRABaseModel())
    net.force_to(DEV, torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, model, False, True)
    g = torch.Generator().manual_seed(seed + 1)
    onets = {}
    for name, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        om = copy.deepcopy(omodel).to(DEV, dt)
        on = lora_ref.LoRANetworkRef(om, lora_dim=rank, target_class="WanTransformer3DModel", block_substr="blocks")
        on.to(DEV, torch.float32)
        onets[name] = (om, on)
    with torch.no_grad():
        for i, lora in enumerate(net.get_all_modules()):
            up = torch.randn(lora.lora_up.weight.shape, generator=g) * 0.05
            lora.lora_up.weight.copy_(up)
            for om, on in onets.values():
                ol = on.loras[i]
                assert ol.lora_name == lora.lora_name
                ol.lora_down.weight.copy_(lora.lora_down.weight)
                ol.lora_up.weight.copy_(up)
    net.mark_params_changed()
    lat = torch.randn(B, 16, Fr, Hh, Ww, generator=g).bfloat16().to(DEV)
    noise = torch.randn(B, 16, Fr, Hh, Ww, generator=g).bfloat16().to(DEV)
    t = torch.tensor([500.0, 250.0, 750.0][:B], device=DEV)
    text = (torch.randn(B, Lt, TOY["text_dim"], generator=g) * 0.5).bfloat16().to(DEV)
    return model, net, onets, (lat, noise, t, text)


def _oracle_step_wan(om, on, batch, dtype):
    lat, noise, t, text = batch
    noisy = lora_ref.add_noise_flowmatch(lat, noise, t).to(torch.bfloat16).to(dtype)
    on.zero_grad(set_to_none=True)
    with on:
        pred = wan_ref.wan_predict(om, noisy, t, text.to(dtype))
        loss = lora_ref.flow_loss(pred, lat, noise)
        loss.backward()
    grads = torch.cat([p.grad.reshape(-1) for lora in on.loras for p in (lora.lora_down.weight, lora.lora_up.weight)])
    return loss.item(), pred.detach(), grads


@pytest.mark.gpu
@pytest.mark.parametrize("B,Fr,Hh,Ww,Lt,rank", [(1, 3, 8, 8, 24, 8), (2, 2, 8, 12, 40, 16), (1, 1, 16, 16, 8, 4)])
def test_wan_engine_step_matches_oracle(B, Fr, Hh, Ww, Lt, rank):
    """Forward (prediction), flow-matching loss and every LoRA gradient of the fused Wan engine against the oracle on
    identical weights / latents / timesteps / embeddings; tolerance as for the FLUX engine:
    max(1e-3, 1.5 x the bf16 eager oracle's own distance to the fp32 oracle)."""
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.train_step import WanLoRATrainStep
    model, net, onets, batch = _setup_wan(B, Fr, Hh, Ww, Lt, rank)
    lat, noise, t, text = batch
    loss32, pred32, g32 = _oracle_step_wan(*onets["fp32"], batch, torch.float32)
    loss16, pred16, g16 = _oracle_step_wan(*onets["bf16"], batch, torch.bfloat16)
    opt = B200AdamW(net, lr=0.0, weight_decay=0.0, max_grad_norm=0.0)
    step = WanLoRATrainStep(model, net, opt, batch_size=B, latent_shape=(16, Fr, Hh, Ww), text_len=Lt, use_cuda_graph=False)
    out = step.hook_train_loop(dict(latents=lat, noise=noise, timesteps=t, text_embeds=text))
    g = net.flat_grads[:g32.numel()]
    # the model's own forward (diffusers call signature) for the prediction
    noisy = lora_ref.add_noise_flowmatch(lat, noise, t).to(torch.bfloat16)
    with net, torch.no_grad():
        pred = model(noisy, t, text)[0]
    floor_pred, floor_g = _rel(pred16, pred32), _rel(g16, g32)
    floor_loss = abs(loss16 - loss32) / abs(loss32)
    e_pred, e_g, e_loss = _rel(pred, pred32), _rel(g, g32), abs(out["loss"] - loss32) / abs(loss32)
    print(f"[wan] pred rel {e_pred:.3e} (bf16 eager floor {floor_pred:.3e}); grads {e_g:.3e} (floor {floor_g:.3e}); loss "
          f"{e_loss:.3e} (floor {floor_loss:.3e})")
    assert pred.shape == pred32.shape and g32.norm() > 0
    assert e_pred < max(1e-3, 1.5 * floor_pred)
    assert e_g < max(1e-3, 1.5 * floor_g)
    assert e_loss < max(1e-3, 1.5 * floor_loss)
    off = 0
    for lora in net.get_all_modules():  # no single adapter may be off
        for w in (lora.lora_down.weight, lora.lora_up.weight):
            n = w.numel()
            assert _rel(g[off:off + n], g32[off:off + n]) < max(8e-3, 4 * floor_g), lora.lora_name
            off += n


@pytest.mark.gpu
def test_wan_loss_curve_with_cuda_graphs():
    """20 optimizer steps of WanLoRATrainStep (CUDA graphs, clip + AdamW + EMA) vs the eager bf16 oracle + torch AdamW."""
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.train_step import WanLoRATrainStep
    B, Fr, Hh, Ww, Lt = 2, 2, 8, 8, 24
    model, net, onets, batch = _setup_wan(B, Fr, Hh, Ww, Lt, 8, seed=4)
    lat, noise, t, text = batch
    om, on = onets["bf16"]
    oparams = [p for l in on.loras for p in (l.lora_down.weight, l.lora_up.weight)]
    oopt = torch.optim.AdamW(oparams, lr=2e-4, eps=1e-6, weight_decay=1e-2)
    opt = B200AdamW(net, lr=2e-4, eps=1e-6, weight_decay=1e-2, max_grad_norm=1.0, ema_decay=0.99)
    step = WanLoRATrainStep(model, net, opt, batch_size=B, latent_shape=(16, Fr, Hh, Ww), text_len=Lt, use_cuda_graph=True)
    bd = dict(latents=lat, noise=noise, timesteps=t, text_embeds=text)
    mine, ref = [], []
    for it in range(20):
        mine.append(step.hook_train_loop(bd)["loss"])
        oopt.zero_grad(set_to_none=True)
        noisy = lora_ref.add_noise_flowmatch(lat, noise, t).to(torch.bfloat16)
        with on:
            loss = lora_ref.flow_loss(wan_ref.wan_predict(om, noisy, t, text), lat, noise)
            loss.backward()
        torch.nn.utils.clip_grad_norm_(oparams, 1.0)
        oopt.step()
        ref.append(loss.item())
    rel = max(abs(a - b) / abs(b) for a, b in zip(mine, ref))
    print("wan loss curve max rel diff", rel)
    assert rel < 1e-3 and int(opt.state_buf[0].item()) == 20
