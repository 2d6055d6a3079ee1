"""The fused FLUX engine (forward, loss, backward, optimizer) against the oracle (eager restatement of the
diffusers blocks + the reference's LoRA forward) on identical weights / latents / timesteps / embeddings.

Metric: the oracle in fp32 is the reference value; the bf16 eager oracle's own distance to
it is the bf16 noise floor.  The B200 path must be within max(1e-3, 1.5 x that floor) relative error on the loss
and on the fp32 LoRA gradients (they are bf16-rounding limited), and bit-exact on index ops (pack / ids)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _setup(layers, single, heads, B, hl, wl, Lt, rank, seed=0, text_dim=64, pooled_dim=32, std=0.05, up_std=0.05,
           precisions=("bf16", "fp32")):
    from oracle import flux_ref, lora_ref
    from ai_toolkit_b200 import LoRASpecialNetwork
    from ai_toolkit_b200.flux import FluxConfig, FluxTransformer2DModel
    torch.manual_seed(seed)
    cfgd = dict(num_layers=layers, num_single_layers=single, num_attention_heads=heads, joint_attention_dim=text_dim,
                pooled_projection_dim=pooled_dim)
    ocfg = flux_ref.FluxConfig(**cfgd)
    omodel = flux_ref.init_synthetic_(flux_ref.FluxTransformer2DModel(ocfg), seed=seed, std=std)
    omodel.requires_grad_(False)
    model = FluxTransformer2DModel(FluxConfig(**cfgd), device=DEV)
    missing = model.load_state_dict(omodel.state_dict(), strict=True)
    net = LoRASpecialNetwork(text_encoder=None, unet=model, lora_dim=rank, alpha=rank, train_unet=True,
                             train_text_encoder=False, is_flux=True, transformer_only=True)
    net.force_to(DEV, torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, model, False, True)
    # oracle networks (bf16 eager and fp32 eager) with the same adapter values, lora_up made non-zero
    g = torch.Generator().manual_seed(seed + 1)
    onets = {}
    for name, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        if name not in precisions:
            continue
        om = copy.deepcopy(omodel).to(DEV, dt)
        on = lora_ref.LoRANetworkRef(om, lora_dim=rank)
        on.to(DEV, torch.float32)
        onets[name] = (om, on)
    with torch.no_grad():
        for i, lora in enumerate(net.get_all_modules()):
            up = torch.randn(lora.lora_up.weight.shape, generator=g) * up_std
            lora.lora_up.weight.copy_(up)
            for om, on in onets.values():
                ol = on.loras[i]
                assert ol.lora_name == lora.lora_name
                ol.lora_down.weight.copy_(lora.lora_down.weight)
                ol.lora_up.weight.copy_(up)
    net.mark_params_changed()
    lat = torch.randn(B, 16, hl, wl, generator=g).bfloat16().to(DEV)
    noise = torch.randn(B, 16, hl, wl, generator=g).bfloat16().to(DEV)
    t = torch.tensor([500.0, 250.0, 750.0][:B], device=DEV)  # t/1000 exact in bf16: the fp32 oracle sees the same timestep
    text = (torch.randn(B, Lt, text_dim, generator=g) * 0.5).bfloat16().to(DEV)
    pooled = torch.randn(B, pooled_dim, generator=g).bfloat16().to(DEV)
    return model, net, onets, (lat, noise, t, text, pooled)


def _oracle_step(om, on, batch, dtype):
    from oracle import flux_ref, lora_ref
    lat, noise, t, text, pooled = batch
    noisy = lora_ref.add_noise_flowmatch(lat, noise, t).to(torch.bfloat16).to(dtype)
    on.zero_grad(set_to_none=True)
    with on:
        pred = lora_ref.flux_predict(om, noisy, t, text.to(dtype), pooled.to(dtype), 1.0, flux_ref.pack_latents,
                                     flux_ref.unpack_latents, flux_ref.make_img_ids)
        loss = lora_ref.flow_loss(pred, lat, noise)
        loss.backward()
    grads = torch.cat([p.grad.reshape(-1) for lora in on.loras for p in (lora.lora_down.weight, lora.lora_up.weight)])
    return loss.item(), pred.detach(), grads


@pytest.mark.parametrize("layers,single,heads,B,hl,wl,Lt,rank", [(1, 1, 2, 1, 16, 16, 24, 4), (2, 2, 2, 2, 16, 24, 40, 16),
                                                                   (1, 1, 2, 1, 16, 16, 24, 32), (1, 1, 2, 3, 16, 16, 8, 64)])
def test_engine_step_matches_oracle(layers, single, heads, B, hl, wl, Lt, rank):
    from oracle import flux_ref
    from ai_toolkit_b200 import ops
    from ai_toolkit_b200.train_step import make_img_ids
    model, net, onets, batch = _setup(layers, single, heads, B, hl, wl, Lt, rank)
    lat, noise, t, text, pooled = batch
    loss32, pred32, g32 = _oracle_step(*onets["fp32"], batch, torch.float32)
    loss16, pred16, g16 = _oracle_step(*onets["bf16"], batch, torch.bfloat16)
    # index ops: bit-exact
    assert torch.equal(make_img_ids(hl, wl, DEV), flux_ref.make_img_ids(hl, wl, DEV))
    packed = ops.flow_add_noise(lat, noise, t, pack=True)
    eng = model.engine
    net.flat_grads.zero_()
    img_ids, txt_ids = make_img_ids(hl, wl, DEV), torch.zeros(Lt, 3, device=DEV)
    guidance = torch.ones(B, device=DEV)
    with net:
        pred = eng.forward(packed, t, text, pooled, guidance, txt_ids, img_ids, save=True, t_div=1000.0)
        tot, per, dpred = ops.flow_loss(pred.view(B, -1, 64), lat, noise, pack=True)
        eng.backward(dpred.view(-1, 64))
    torch.cuda.synchronize()
    pred_unp = flux_ref.unpack_latents(pred.view(B, -1, 64), hl, wl)
    floor_pred, floor_g = _rel(pred16, pred32), _rel(g16, g32)
    floor_loss = abs(loss16 - loss32) / abs(loss32)
    e_pred, e_g = _rel(pred_unp, pred32), _rel(net.flat_grads[:g32.numel()], g32)
    e_loss = abs(tot.item() - loss32) / abs(loss32)
    print(f"pred rel err {e_pred:.3e} (bf16 eager floor {floor_pred:.3e}); grads {e_g:.3e} (floor {floor_g:.3e}); "
          f"loss {e_loss:.3e} (floor {floor_loss:.3e}); vs bf16 eager: pred {_rel(pred_unp, pred16):.3e} "
          f"grads {_rel(net.flat_grads[:g32.numel()], g16):.3e} loss {abs(tot.item() - loss16) / abs(loss16):.3e}")
    assert e_pred < max(1e-3, 1.5 * floor_pred)
    assert e_g < max(1e-3, 1.5 * floor_g)
    assert e_loss < max(1e-3, 1.5 * floor_loss)


def test_engine_step_matches_oracle_at_flux_dims():
    """BASELINE.json configs[2] geometry on the blocks themselves: D = 3072, 24 heads x 128, MLP 12288, 4096 image + 512
    text tokens, T5 width 4096, pooled 768, r = 16 -- one double + one single block (the 19 + 38 of the full model repeat
    these two).  Engine (forward, loss, backward) vs the oracle on the GPU in fp32 AND in bf16:
      * loss (fp32 scalar)          <= 1e-3 relative to the fp32 oracle      (north-star criterion, asserted as is)
      * prediction, dA / dB (fp32)  reported against the fp32 oracle and against the bf16 eager oracle; asserted
                                    <= max(1e-3, 1.5 x the bf16 eager oracle's own distance to the fp32 oracle): bf16
                                    activations carry 2^-9 relative rounding per element, which no implementation that
                                    stores bf16 activations can undercut (the reference's own eager path sits AT this floor)
    The tensor-core accumulation itself is checked at <= 1e-3 against fp32 math on identical bf16 operands in
    tests/test_gpu_gemm.py (dgrad / wgrad at these shapes)."""
    from oracle import flux_ref
    from ai_toolkit_b200 import ops
    from ai_toolkit_b200.train_step import make_img_ids
    B, hl, wl, Lt, rank = 1, 128, 128, 512, 16
    model, net, onets, batch = _setup(1, 1, 24, B, hl, wl, Lt, rank, seed=3, text_dim=4096, pooled_dim=768, std=0.02,
                                      up_std=0.02)
    assert model.cfg.inner_dim == 3072 and len(net.get_all_modules()) == 14 + 6
    lat, noise, t, text, pooled = batch
    loss32, pred32, g32 = _oracle_step(*onets["fp32"], batch, torch.float32)
    loss16, pred16, g16 = _oracle_step(*onets["bf16"], batch, torch.bfloat16)
    packed = ops.flow_add_noise(lat, noise, t, pack=True)
    net.flat_grads.zero_()
    with net:
        pred = model.engine.forward(packed, t, text, pooled, torch.ones(B, device=DEV), torch.zeros(Lt, 3, device=DEV),
                                    make_img_ids(hl, wl, DEV), save=True, t_div=1000.0)
        tot, per, dpred = ops.flow_loss(pred.view(B, -1, 64), lat, noise, pack=True)
        model.engine.backward(dpred.view(-1, 64))
    torch.cuda.synchronize()
    pred_unp = flux_ref.unpack_latents(pred.view(B, -1, 64), hl, wl)
    g = net.flat_grads[:g32.numel()]
    floor_pred, floor_g = _rel(pred16, pred32), _rel(g16, g32)
    e_pred, e_g = _rel(pred_unp, pred32), _rel(g, g32)
    e_loss, floor_loss = abs(tot.item() - loss32) / abs(loss32), abs(loss16 - loss32) / abs(loss32)
    print(f"[FLUX dims] loss rel {e_loss:.3e} (bf16 eager oracle: {floor_loss:.3e}) | pred vs fp32 oracle {e_pred:.3e} "
          f"(bf16 eager: {floor_pred:.3e}), vs bf16 eager {_rel(pred_unp, pred16):.3e} | dA/dB vs fp32 oracle {e_g:.3e} "
          f"(bf16 eager: {floor_g:.3e}), vs bf16 eager {_rel(g, g16):.3e}")
    assert torch.isfinite(g).all() and g32.norm() > 0
    assert e_loss < 1e-3
    assert e_pred < max(1e-3, 1.5 * floor_pred)
    assert e_g < max(1e-3, 1.5 * floor_g)
    # per-adapter: no single module may be off (a wrong slice / stride would hide in the global norm)
    off = 0
    for lora in net.get_all_modules():
        for w in (lora.lora_down.weight, lora.lora_up.weight):
            n = w.numel()
            a, b = g[off:off + n], g32[off:off + n]
            assert _rel(a, b) < max(5e-3, 3 * floor_g), (lora.lora_name, _rel(a, b))
            off += n


def test_lora_module_autograd_dropin():
    """The per-module seam (LoRAModule.forward inside an eager model) against the oracle's adapter forward."""
    from oracle import lora_ref
    from ai_toolkit_b200 import LoRASpecialNetwork
    torch.manual_seed(0)

    class Toy(torch.nn.Module):  # class name the reference targets for generic transformers
        def __init__(self):
            super().__init__()
            self.transformer_blocks = torch.nn.ModuleList([torch.nn.Linear(256, 512), torch.nn.Linear(512, 256)])

        def forward(self, x):
            return self.transformer_blocks[1](torch.nn.functional.gelu(self.transformer_blocks[0](x), approximate="tanh"))

    Toy.__name__ = "FluxTransformer2DModel"
    m = Toy().to(DEV, torch.bfloat16).requires_grad_(False)
    m2 = copy.deepcopy(m)
    net = LoRASpecialNetwork(None, m, lora_dim=8, alpha=8, train_text_encoder=False, is_flux=True, transformer_only=True)
    net.force_to(DEV, torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, m, False, True)
    ref = lora_ref.LoRANetworkRef(m2, lora_dim=8).to(DEV)
    with torch.no_grad():
        for a, b in zip(net.get_all_modules(), ref.loras):
            a.lora_up.weight.normal_(0, 0.05)
            b.lora_down.weight.copy_(a.lora_down.weight)
            b.lora_up.weight.copy_(a.lora_up.weight)
    net.mark_params_changed()
    x = torch.randn(3, 200, 256, device=DEV).bfloat16()
    with net:
        y = m(x)
        (y.float() ** 2).mean().backward()
    with ref:
        y2 = m2(x)
        (y2.float() ** 2).mean().backward()
    assert _rel(y, y2) < 1e-2
    for a, b in zip(net.get_all_modules(), ref.loras):
        assert _rel(a.lora_down.weight.grad, b.lora_down.weight.grad) < 2e-2
        assert _rel(a.lora_up.weight.grad, b.lora_up.weight.grad) < 2e-2
    # inactive network = frozen model, untouched
    assert torch.equal(m(x), m2(x))


def test_engine_vs_reference_golden():
    """Committed golden vectors produced by the UNMODIFIED reference (CPU fp32, oracle/make_golden.py): the B200 path in
    bf16 must land within bf16 rounding of them (pred 2e-2 rel, loss 2e-3 rel, LoRA grads 3e-2 rel)."""
    import os
    from oracle import flux_ref, make_golden
    from ai_toolkit_b200 import LoRASpecialNetwork, ops
    from ai_toolkit_b200.flux import FluxConfig, FluxTransformer2DModel
    from ai_toolkit_b200.train_step import make_img_ids
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "lora_tiny.pt"), weights_only=False)
    cfg, omodel, batch = make_golden.build()
    model = FluxTransformer2DModel(FluxConfig(num_layers=1, num_single_layers=1, num_attention_heads=2, joint_attention_dim=64,
                                              pooled_projection_dim=32), device=DEV)
    model.load_state_dict(omodel.state_dict())
    net = LoRASpecialNetwork(None, model, lora_dim=4, alpha=4, train_text_encoder=False, is_flux=True, transformer_only=True)
    net.force_to(DEV, torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, model, False, True)
    net.load_state_dict(gold["init_state_dict"])
    net.mark_params_changed()
    lat, noise = batch["latents"].bfloat16().to(DEV), batch["noise"].bfloat16().to(DEV)
    t = batch["timesteps"].to(DEV)
    B, Lt = 2, 8
    packed = ops.flow_add_noise(lat, noise, t, pack=True)
    net.flat_grads.zero_()
    with net:
        pred = model.engine.forward(packed, t, batch["text"].bfloat16().to(DEV), batch["pooled"].bfloat16().to(DEV),
                                    torch.ones(B, device=DEV), torch.zeros(Lt, 3, device=DEV), make_img_ids(8, 8, DEV),
                                    save=True, t_div=1000.0)
        tot, _, dpred = ops.flow_loss(pred.view(B, -1, 64), lat, noise, pack=True)
        model.engine.backward(dpred.view(-1, 64))
    torch.cuda.synchronize()
    assert _rel(flux_ref.unpack_latents(pred.view(B, -1, 64), 8, 8), gold["pred0"].to(DEV)) < 2e-2
    assert abs(tot.item() - gold["losses"][0]) / gold["losses"][0] < 2e-3
    gref = torch.cat([gold["grads0"][n].reshape(-1) for n, _ in net.named_parameters()]).to(DEV)
    assert _rel(net.flat_grads[:gref.numel()], gref) < 3e-2


def test_loss_curve_matches_oracle_over_steps():
    """20 optimizer steps (clip 1.0 + AdamW eps 1e-6 + EMA) through FluxLoRATrainStep (CUDA graphs on) vs the eager bf16
    oracle with torch.optim.AdamW: loss per step within 1e-3 relative (north-star criterion, shortened from 100 steps to
    keep the GPU tier fast), parameters after the run within 2e-2 relative of the oracle's update."""
    from oracle import flux_ref, lora_ref
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.train_step import FluxLoRATrainStep
    model, net, onets, batch = _setup(1, 1, 2, 2, 16, 16, 24, 8, seed=5)
    lat, noise, t, text, pooled = batch
    om, on = onets["bf16"]
    oparams = [p for l in on.loras for p in (l.lora_down.weight, l.lora_up.weight)]
    p_init = torch.cat([p.detach().reshape(-1) for p in oparams]).clone()
    oopt = torch.optim.AdamW(oparams, lr=2e-4, eps=1e-6, weight_decay=1e-2)
    opt = B200AdamW(net, lr=2e-4, eps=1e-6, weight_decay=1e-2, max_grad_norm=1.0, ema_decay=0.99)
    step = FluxLoRATrainStep(model, net, opt, batch_size=2, latent_shape=(16, 16, 16), text_len=24, use_cuda_graph=True)
    step.text = torch.zeros_like(text)
    step.pooled = torch.zeros_like(pooled)
    bd = dict(latents=lat, noise=noise, timesteps=t, text_embeds=text, pooled_embeds=pooled)
    mine, ref = [], []
    for it in range(20):
        mine.append(step.hook_train_loop(bd)["loss"])
        oopt.zero_grad(set_to_none=True)
        noisy = lora_ref.add_noise_flowmatch(lat, noise, t).to(torch.bfloat16)
        with on:
            pred = lora_ref.flux_predict(om, noisy, t, text, pooled, 1.0, flux_ref.pack_latents, flux_ref.unpack_latents,
                                         flux_ref.make_img_ids)
            loss = lora_ref.flow_loss(pred, lat, noise)
            loss.backward()
        torch.nn.utils.clip_grad_norm_(oparams, 1.0)
        oopt.step()
        ref.append(loss.item())
    rel = max(abs(a - b) / abs(b) for a, b in zip(mine, ref))
    print("loss curve max rel diff", rel, mine[:3], ref[:3])
    assert rel < 1e-3
    p_ref = torch.cat([p.detach().reshape(-1) for p in oparams])
    p_mine = net.flat_params[:p_ref.numel()]
    assert _rel(p_mine - p_init, p_ref - p_init) < 5e-2
    assert int(opt.state_buf[0].item()) == 20


def test_gradient_accumulation_sums_micro_batches():
    """hook_train_loop([b0, b1]) == reference semantics: gradients summed over micro-batches, one optimizer step, mean
    loss (SDTrainer.py:2250-2268, :2312).  Checked against two separate engine backward passes."""
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.train_step import FluxLoRATrainStep
    model, net, onets, batch = _setup(1, 1, 2, 2, 16, 16, 24, 8, seed=9)
    lat, noise, t, text, pooled = batch
    opt = B200AdamW(net, lr=0.0, weight_decay=0.0, max_grad_norm=0.0)  # lr 0: keep the parameters, look at the gradients
    step = FluxLoRATrainStep(model, net, opt, batch_size=1, latent_shape=(16, 16, 16), text_len=24, use_cuda_graph=False)
    mb = [dict(latents=lat[i:i + 1], noise=noise[i:i + 1], timesteps=t[i:i + 1], text_embeds=text[i:i + 1],
               pooled_embeds=pooled[i:i + 1]) for i in range(2)]
    g = []
    losses = []
    for b in mb:
        losses.append(step.hook_train_loop(b)["loss"])
        g.append(net.flat_grads.clone())
    out = step.hook_train_loop(mb)
    assert abs(out["loss"] - sum(losses) / 2) < 1e-6 * abs(out["loss"]) + 1e-7
    assert _rel(net.flat_grads, g[0] + g[1]) < 1e-5
    assert int(opt.state_buf[0].item()) == 3


def test_loss_curve_100_steps_with_validation_protocol(tmp_path):
    """North-star criterion: 100 optimizer steps, loss per step within 1e-3 relative of the eager bf16 oracle +
    torch.optim.AdamW(eps 1e-6) + clip_grad_norm_(1.0), plus the reference's deterministic validation loss
    (BaseSDTrainProcess.validate :1681-1743: fixed CPU-generator noise seeds 42 + i, sigmas 1.0 / 0.75 / 0.5 / 0.25, network
    active, no grad) at steps 0, 50 and 100 within 1e-3; post-training parameters: the UPDATE (p - p_init) within 5e-2.
    A fresh batch every step (4 samples cycled with different noise / timesteps), CUDA graphs on, EMA on.  Then the state is
    saved from the device flat buffers, reloaded into a second network / optimizer, and one more step matches."""
    from oracle import flux_ref, lora_ref
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.train_step import FluxLoRATrainStep
    B, hl, wl, Lt = 2, 16, 16, 24
    model, net, onets, batch = _setup(1, 1, 2, B, hl, wl, Lt, 8, seed=13, precisions=("bf16",))
    _, _, _, text, pooled = batch
    om, on = onets["bf16"]
    oparams = [p for l in on.loras for p in (l.lora_down.weight, l.lora_up.weight)]
    p_init = torch.cat([p.detach().reshape(-1) for p in oparams]).clone()
    oopt = torch.optim.AdamW(oparams, lr=2e-4, eps=1e-6, weight_decay=1e-2)
    opt = B200AdamW(net, lr=2e-4, eps=1e-6, weight_decay=1e-2, max_grad_norm=1.0, ema_decay=0.99)
    step = FluxLoRATrainStep(model, net, opt, batch_size=B, latent_shape=(16, hl, wl), text_len=Lt, use_cuda_graph=True)
    g = torch.Generator().manual_seed(99)
    data = torch.randn(4, 16, hl, wl, generator=g).bfloat16()
    val_lat = [data[0:1].float(), data[3:4].float()]
    val_emb = [(text[0:1], pooled[0:1]), (text[1:2], pooled[1:2])]
    table = torch.linspace(1000, 1, 1000)

    def oracle_validate():
        losses = []
        sig = torch.tensor([1000.0, 750.0, 500.0, 250.0], device=DEV)
        with torch.no_grad(), on:
            for i, (lat, (te, pe)) in enumerate(zip(val_lat, val_emb)):
                noise = torch.randn(lat.shape, generator=torch.Generator(device="cpu").manual_seed(42 + i), dtype=torch.float32)
                lb = torch.cat([lat.to(DEV, torch.bfloat16)] * 4, 0)
                nb = torch.cat([noise.to(DEV, torch.bfloat16)] * 4, 0)
                noisy = lora_ref.add_noise_flowmatch(lb, nb, sig).to(torch.bfloat16)
                pred = lora_ref.flux_predict(om, noisy, sig, torch.cat([te] * 4, 0), torch.cat([pe] * 4, 0), 1.0,
                                             flux_ref.pack_latents, flux_ref.unpack_latents, flux_ref.make_img_ids)
                losses.append(torch.nn.functional.mse_loss(pred.float(), (nb - lb).float()))
        return float(torch.stack(losses).mean())

    mine, ref, vals = [], [], []
    for it in range(100):
        if it in (0, 50):
            vals.append((step.validate(val_lat, val_emb), oracle_validate()))
        sel = torch.tensor([(2 * it) % 4, (2 * it + 1) % 4])
        lat = data[sel].to(DEV)
        noise = torch.randn(B, 16, hl, wl, generator=g).bfloat16().to(DEV)
        t = table[torch.randint(0, 999, (B,), generator=g)].to(DEV)
        mine.append(step.hook_train_loop(dict(latents=lat, noise=noise, timesteps=t, text_embeds=text, pooled_embeds=pooled))["loss"])
        oopt.zero_grad(set_to_none=True)
        noisy = lora_ref.add_noise_flowmatch(lat, noise, t).to(torch.bfloat16)
        with on:
            pred = lora_ref.flux_predict(om, noisy, t, text, pooled, 1.0, flux_ref.pack_latents, flux_ref.unpack_latents,
                                         flux_ref.make_img_ids)
            loss = lora_ref.flow_loss(pred, lat, noise)
            loss.backward()
        torch.nn.utils.clip_grad_norm_(oparams, 1.0)
        oopt.step()
        ref.append(loss.item())
    vals.append((step.validate(val_lat, val_emb), oracle_validate()))
    rel = [abs(a - b) / abs(b) for a, b in zip(mine, ref)]
    vrel = [abs(a - b) / abs(b) for a, b in vals]
    p_ref = torch.cat([p.detach().reshape(-1) for p in oparams])
    upd = _rel(net.flat_params[:p_ref.numel()] - p_init, p_ref - p_init)
    print(f"100-step loss curve: max rel {max(rel):.3e} (mean {sum(rel) / len(rel):.3e}); validation loss rel at 0/50/100: "
          + " ".join(f"{v:.3e}" for v in vrel) + f"; val loss {vals[0][0]:.5f} -> {vals[-1][0]:.5f}; update rel err {upd:.3e}")
    assert max(rel) < 1e-3
    assert max(vrel) < 1e-3
    assert vals[-1][0] < vals[0][0]  # it trains
    assert upd < 5e-2 and int(opt.state_buf[0].item()) == 100
    # ---- save from the device flat buffers, resume into a fresh network / optimizer / step, continue identically
    root = str(tmp_path / "ckpt")
    step.save(root, "curve", step=100, epoch=1, dtype=torch.float32)
    model2, net2, _, _ = _setup(1, 1, 2, B, hl, wl, Lt, 8, seed=13, precisions=())
    opt2 = B200AdamW(net2, lr=2e-4, eps=1e-6, weight_decay=1e-2, max_grad_norm=1.0, ema_decay=0.99)
    step2 = FluxLoRATrainStep(model2, net2, opt2, batch_size=B, latent_shape=(16, hl, wl), text_len=Lt, use_cuda_graph=False)
    path, st, ep = step2.resume(root, "curve")
    assert (st, ep) == (100, 1) and path.endswith("curve_000000100.safetensors")
    assert torch.equal(net2.flat_params, net.flat_params) and int(opt2.state_buf[0]) == 100
    assert torch.equal(opt2.exp_avg, opt.exp_avg) and torch.equal(opt2.ema, net2.flat_params)
    bd = dict(latents=data[:2].to(DEV), noise=torch.randn(B, 16, hl, wl, generator=g).bfloat16().to(DEV),
              timesteps=torch.tensor([300.0, 700.0], device=DEV), text_embeds=text, pooled_embeds=pooled)
    l1 = step.hook_train_loop(bd)["loss"]
    l2 = step2.hook_train_loop(bd)["loss"]
    assert abs(l1 - l2) < 1e-6 * abs(l1) + 1e-8
    assert _rel(net2.flat_params, net.flat_params) < 1e-6


def test_prepare_batch_draws_the_reference_timesteps_and_weights():
    """a3 / a6 wired into the step: `prepare_batch` = set_train_timesteps + randint index + table[idx] + randn noise
    (BaseSDTrainProcess.py:1195-1421), bit-exact on the index ops given the generator; timestep weights reach the loss."""
    from ai_toolkit_b200 import timesteps as ts
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.train_step import FluxLoRATrainStep
    B, hl, wl, Lt = 2, 16, 16, 24
    model, net, onets, batch = _setup(1, 1, 2, B, hl, wl, Lt, 4, seed=3, precisions=())
    lat, _, _, text, pooled = batch
    opt = B200AdamW(net, lr=0.0, weight_decay=0.0, max_grad_norm=0.0)
    for ttype in ("sigmoid", "linear"):
        step = FluxLoRATrainStep(model, net, opt, batch_size=B, latent_shape=(16, hl, wl), text_len=Lt, use_cuda_graph=False,
                                 timestep_type=ttype, linear_timesteps=(ttype == "linear"), noise_multiplier=1.5,
                                 use_loss_options=True)
        g1, g2 = torch.Generator(device=DEV).manual_seed(5), torch.Generator(device=DEV).manual_seed(5)
        t, idx = step.prepare_batch(lat, text, pooled, generator=g1, loss_multiplier=[1.0, 0.5])
        table = ts.set_train_timesteps(1000, DEV, ttype, generator=g2)
        idx2 = torch.randint(0, 999, (B,), device=DEV, generator=g2).long()
        noise2 = torch.randn(lat.shape, device=DEV, dtype=lat.dtype, generator=g2) * 1.5
        assert torch.equal(idx, idx2) and torch.equal(t, table[idx2].float())
        assert torch.equal(step.timesteps, t) and torch.equal(step.noise, noise2.to(torch.bfloat16))
        want_w = torch.tensor([1.0, 0.5], device=DEV)
        if ttype == "linear":
            want_w = want_w * ts.weights_for_timesteps(table, t).to(DEV)
        assert torch.allclose(step.sample_weight, want_w)
        weighted = step.hook_train_loop(dict(latents=step.latents.clone(), noise=step.noise.clone(), timesteps=t, text_embeds=text,
                                             pooled_embeds=pooled))["loss"]
        per = step.loss_ws[:B].clone()
        step.sample_weight.fill_(1.0)
        step.run()
        plain = step.loss_ws[:B].clone()
        assert torch.allclose(per, plain * want_w, rtol=1e-5)
        assert weighted == pytest.approx(float((plain * want_w).mean()), rel=1e-5)


def _nccl_worker(rank, world, port, out_dir):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    global DEV
    DEV = f"cuda:{rank}"
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.train_step import FluxLoRATrainStep
    B, hl, wl, Lt = 1, 16, 16, 24
    model, net, onets, batch = _setup(1, 1, 2, 2, hl, wl, Lt, 8, seed=17, precisions=())
    lat, noise, t, text, pooled = batch
    with torch.no_grad():  # replicas start DIFFERENT: the step's constructor must broadcast rank 0's values
        net.flat_params.add_(0.01 * rank)
    net.mark_params_changed()
    opt = B200AdamW(net, lr=1e-3, eps=1e-6, weight_decay=1e-2, max_grad_norm=1.0, ema_decay=0.99)  # prescale default 1.0
    step = FluxLoRATrainStep(model, net, opt, batch_size=B, latent_shape=(16, hl, wl), text_len=Lt, use_cuda_graph=True)
    sl = slice(rank, rank + 1)
    bd = dict(latents=lat[sl], noise=noise[sl], timesteps=t[sl], text_embeds=text[sl], pooled_embeds=pooled[sl])
    losses = [step.hook_train_loop(bd)["loss"] for _ in range(4)]  # eager, eager, captured, replay
    torch.cuda.synchronize()
    torch.save(dict(params=net.flat_params.cpu(), grads=net.flat_grads.cpu(), losses=losses, norm=float(opt.grad_norm)),
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_ranks_nccl_equal_accumulation_over_two_samples(tmp_path):
    """SURVEY.md section 8e on real GPUs: 2 ranks x 1 sample over NCCL == 1 rank accumulating the same 2 samples with
    grad_prescale 1/2 -- parameters after 4 optimizer steps (clipping sees the GLOBAL averaged gradient), through CUDA
    graphs with the all-reduce between the two graph replays."""
    import torch.multiprocessing as mp
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.train_step import FluxLoRATrainStep
    mp.spawn(_nccl_worker, args=(2, 29617, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(str(tmp_path / f"r{k}.pt")) for k in range(2))
    assert torch.equal(r0["params"], r1["params"])  # replicas stay identical
    assert torch.equal(r0["grads"], r1["grads"])
    B, hl, wl, Lt = 1, 16, 16, 24
    model, net, onets, batch = _setup(1, 1, 2, 2, hl, wl, Lt, 8, seed=17, precisions=())
    lat, noise, t, text, pooled = batch
    opt = B200AdamW(net, lr=1e-3, eps=1e-6, weight_decay=1e-2, max_grad_norm=1.0, ema_decay=0.99, grad_prescale=0.5)
    step = FluxLoRATrainStep(model, net, opt, batch_size=B, latent_shape=(16, hl, wl), text_len=Lt, use_cuda_graph=False)
    mbs = [dict(latents=lat[i:i + 1], noise=noise[i:i + 1], timesteps=t[i:i + 1], text_embeds=text[i:i + 1],
                pooled_embeds=pooled[i:i + 1]) for i in range(2)]
    losses = [step.hook_train_loop(mbs)["loss"] for _ in range(4)]
    mean_ddp = [(a + b) / 2 for a, b in zip(r0["losses"], r1["losses"])]
    assert max(abs(a - b) / abs(b) for a, b in zip(mean_ddp, losses)) < 1e-4
    assert _rel(r0["params"].to(DEV), net.flat_params) < 1e-4  # fp32 atomics in the wgrad: not bit-identical
    assert abs(r0["norm"] - float(opt.grad_norm)) < 1e-3 * float(opt.grad_norm)


def test_prior_prediction_as_target():
    """`get_prior_prediction` (SDTrainer.py:1211-1339): the frozen model's prediction (network inactive, no grad) is the loss
    target (`calculate_loss`: `target = prior_pred`, :619-621).  The inactive engine must equal the oracle's frozen model, and
    the step's loss must equal mse(active prediction, prior prediction) with LoRA gradients flowing."""
    from oracle import flux_ref, lora_ref
    from ai_toolkit_b200 import ops
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.train_step import FluxLoRATrainStep, make_img_ids
    B, hl, wl, Lt = 2, 16, 16, 24
    model, net, onets, batch = _setup(1, 1, 2, B, hl, wl, Lt, 8, seed=31, precisions=("bf16",))
    lat, noise, t, text, pooled = batch
    om, on = onets["bf16"]
    opt = B200AdamW(net, lr=0.0, weight_decay=0.0, max_grad_norm=0.0)
    step = FluxLoRATrainStep(model, net, opt, batch_size=B, latent_shape=(16, hl, wl), text_len=Lt, use_cuda_graph=False,
                             prior_target=True)
    out = step.hook_train_loop(dict(latents=lat, noise=noise, timesteps=t, text_embeds=text, pooled_embeds=pooled))
    packed = ops.flow_add_noise(lat, noise, t, pack=True)
    args = (packed, t, text, pooled, torch.ones(B, device=DEV), torch.zeros(Lt, 3, device=DEV), make_img_ids(hl, wl, DEV))
    with torch.no_grad():
        prior = model.engine.forward(*args, save=False, t_div=1000.0)  # network inactive
        with net:
            active = model.engine.forward(*args, save=False, t_div=1000.0)
        noisy = lora_ref.add_noise_flowmatch(lat, noise, t).to(torch.bfloat16)
        frozen = lora_ref.flux_predict(om, noisy, t, text, pooled, 1.0, flux_ref.pack_latents, flux_ref.unpack_latents,
                                       flux_ref.make_img_ids)  # oracle network inactive = the frozen oracle model
    assert _rel(flux_ref.unpack_latents(prior.view(B, -1, 64), hl, wl), frozen) < 1.5e-2
    want = torch.nn.functional.mse_loss(active.float(), prior.float()).item()
    assert want > 0 and abs(out["loss"] - want) < 1e-5 * want + 1e-9
    assert float(net.flat_grads.abs().sum()) > 0


def test_preservation_micro_batch_adds_to_the_normal_step():
    """diff_output_preservation as the plugin runs it (SDTrainer.py:2182-2219): a normal micro-batch, then a prior-target
    micro-batch on the class embeddings with the multiplier, ONE optimizer step.  The accumulated LoRA gradient must be
    grad(normal) + multiplier * grad(preservation), where both are measured by separate single-pass steps."""
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.train_step import FluxLoRATrainStep
    B, hl, wl, Lt = 2, 16, 16, 24
    model, net, onets, batch = _setup(1, 1, 2, B, hl, wl, Lt, 8, seed=41, precisions=("bf16",))
    lat, noise, t, text, pooled = batch
    g = torch.Generator().manual_seed(5)
    text2 = (torch.randn(text.shape, generator=g) * 0.1).bfloat16().to(DEV)       # the class-prompt embeddings
    pooled2 = torch.randn(pooled.shape, generator=g).bfloat16().to(DEV)
    mult = 0.5
    opt = B200AdamW(net, lr=0.0, weight_decay=0.0, max_grad_norm=0.0)
    kw = dict(batch_size=B, latent_shape=(16, hl, wl), text_len=Lt, use_cuda_graph=False)
    main = FluxLoRATrainStep(model, net, opt, **kw)
    pres = FluxLoRATrainStep(model, net, opt, prior_target=True, loss_multiplier=mult, **kw)
    b_main = dict(latents=lat, noise=noise, timesteps=t, text_embeds=text, pooled_embeds=pooled)
    b_pres = dict(latents=lat, noise=noise, timesteps=t, text_embeds=text2, pooled_embeds=pooled2)
    main._load_dict(b_main)
    l_main = float(main.run(True, True).item())
    g_main = net.flat_grads.clone()
    pres._load_dict(b_pres)
    l_pres = float(pres.run(True, True).item())
    g_pres = net.flat_grads.clone()                  # already x multiplier (gscale)
    assert l_pres > 0 and float(g_pres.abs().sum()) > 0
    main._load_dict(b_main)
    la = float(main.run(first_micro_batch=True, last_micro_batch=False).item())
    pres._load_dict(b_pres)
    lb = float(pres.run(first_micro_batch=False, last_micro_batch=True).item())
    assert la == pytest.approx(l_main, rel=1e-5) and lb == pytest.approx(l_pres, rel=1e-5)
    assert _rel(net.flat_grads, g_main + g_pres) < 1e-3
    # the gradient of the preservation pass scales with the multiplier (the reported loss does not: the plugin multiplies it)
    pres1 = FluxLoRATrainStep(model, net, opt, prior_target=True, loss_multiplier=1.0, **kw)
    pres1._load_dict(b_pres)
    pres1.run(True, True)
    assert _rel(g_pres, mult * net.flat_grads) < 2e-2   # (dpred is rounded to bf16 after the scaling)
