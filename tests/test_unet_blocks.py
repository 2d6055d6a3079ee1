"""The UNet `Transformer2DModel` engine (SD1.5 / SDXL adapter-bearing blocks, BASELINE.json configs[0] / [1]).
CPU: container == oracle parameter names; kohya adapter names / saved keys under a UNet identical to the live reference.
GPU: the row kernels vs torch, and the engine (forward, dX, every LoRA gradient) vs the oracle in fp32 and bf16 for the
SD1.5 form (1x1-conv projections, head dims 40 / 80) and the SDXL form (Linear projections, head dim 64, depth 2)."""
import copy

import pytest
import torch
import torch.nn.functional as F

from ai_toolkit_b200 import LoRASpecialNetwork
from ai_toolkit_b200.unet_blocks import Transformer2DModel
from oracle import lora_ref, ref_import, unet_ref

DEV = "cuda:0"


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_container_matches_oracle_parameter_names():
    for linear in (False, True):
        o = unet_ref.Transformer2DModel(4, 40, 160, num_layers=2, cross_dim=96, use_linear_projection=linear)
        m = Transformer2DModel(4, 40, 160, num_layers=2, cross_dim=96, use_linear_projection=linear, dtype=torch.float32)
        so, sm = o.state_dict(), m.state_dict()
        assert list(so.keys()) == list(sm.keys())
        assert all(so[k].shape == sm[k].shape for k in so)
        m.load_state_dict(so, strict=True)
    Transformer2DModel(8, 160, 1280, device="meta")  # SD1.5's deepest level: CUDA-core attention kernel
    with pytest.raises(NotImplementedError):
        Transformer2DModel(8, 136, 1088)


def _unet_with(t2d_cls, **kw):
    cls = type("UNet2DConditionModel", (torch.nn.Module,), {})

    def init(self):
        torch.nn.Module.__init__(self)
        self.down_blocks = torch.nn.ModuleList([torch.nn.Module()])
        self.down_blocks[0].attentions = torch.nn.ModuleList([t2d_cls(4, 40, 160, num_layers=1, cross_dim=96, **kw)])
        self.mid_block = torch.nn.Module()
        self.mid_block.attentions = torch.nn.ModuleList([t2d_cls(4, 40, 160, num_layers=1, cross_dim=96, **kw)])

    cls.__init__ = init
    return cls()


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_kohya_names_under_a_unet_identical_to_live_reference():
    RefNet, _ = ref_import.reference_lora()
    u1 = _unet_with(unet_ref.Transformer2DModel)
    u2 = _unet_with(Transformer2DModel, dtype=torch.float32)
    kw = dict(text_encoder=None, lora_dim=4, alpha=2, train_unet=True, train_text_encoder=False)
    torch.manual_seed(2)
    r = RefNet(unet=u1, **kw)
    r.force_to("cpu", torch.float32); r._update_torch_multiplier(); r.apply_to(None, u1, False, True)
    torch.manual_seed(2)
    n = LoRASpecialNetwork(unet=u2, **kw)
    n.force_to("cpu", torch.float32); n._update_torch_multiplier(); n.apply_to(None, u2, False, True)
    names = [l.lora_name for l in n.unet_loras]
    assert names == [l.lora_name for l in r.unet_loras]
    assert "lora_unet_down_blocks_0_attentions_0_transformer_blocks_0_attn1_to_q" in names
    assert "lora_unet_mid_block_attentions_0_proj_in" in names and len(names) == 2 * (2 + 10)  # proj_in/out + 10 Linears per block
    sr, sn = r.get_state_dict(dtype=torch.float32), n.get_state_dict(dtype=torch.float32)
    assert list(sr.keys()) == list(sn.keys()) and all(torch.equal(sr[k], sn[k]) for k in sr)


def test_adopt_unet_transformers_shares_storage_and_keeps_names():
    """A loaded eager UNet (here the oracle's, with diffusers' attribute layout) gets its Transformer2DModels replaced by the
    engine's containers over the SAME tensors; state-dict keys and the kohya adapter names under it do not change."""
    from ai_toolkit_b200.unet_blocks import adopt_unet_transformers
    _, oc = _tiny_cfgs("sdxl")
    u = unet_ref.init_synthetic_(unet_ref.UNet2DConditionModel(oc), seed=1)
    kw = dict(text_encoder=None, lora_dim=4, alpha=2, train_unet=True, train_text_encoder=False)
    names_before = [l.lora_name for l in LoRASpecialNetwork(unet=u, **kw).unet_loras]
    before = {k: v.data_ptr() for k, v in u.state_dict().items()}
    assert adopt_unet_transformers(u) == 11 and adopt_unet_transformers(u) == 0
    after = {k: v.data_ptr() for k, v in u.state_dict().items()}
    assert list(before) == list(after) and all(before[k] == after[k] for k in before)
    assert sum(isinstance(m, Transformer2DModel) for m in u.modules()) == 11
    assert [l.lora_name for l in LoRASpecialNetwork(unet=u, **kw).unet_loras] == names_before
    assert not any(p.requires_grad for m in u.modules() if isinstance(m, Transformer2DModel) for p in m.parameters())


def test_adopt_unet_transformers_fails_loudly_on_what_the_engine_does_not_cover():
    from ai_toolkit_b200.unet_blocks import adopt_unet_transformers
    u = unet_ref.UNet2DConditionModel(unet_ref.UNetConfig(block_out_channels=(64, 136 * 2), attn_layers=(1, 1), heads=(1, 2),
                                                          cross_attention_dim=96, norm_groups=8))
    with pytest.raises(NotImplementedError):  # head dim 136: neither <= 128 nor one of the CUDA-core kernel's 160 / 192 / 256
        adopt_unet_transformers(u)
    _, oc = _tiny_cfgs("sdxl")
    u = unet_ref.UNet2DConditionModel(oc)
    u.mid_block.attentions[0].norm.eps = 1e-5
    with pytest.raises(NotImplementedError):  # the container's GroupNorm eps is diffusers' 1e-6
        adopt_unet_transformers(u)


def test_unet_step_save_resume_round_trip_with_kohya_keys(tmp_path):
    """`UNetLoRATrainStep.save / resume`: the reference's file layout with kohya keys for a UNet (host logic; runs on the CPU)."""
    from safetensors.torch import load_file

    from ai_toolkit_b200 import unet as host
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.unet_blocks import adopt_unet_transformers
    _, oc = _tiny_cfgs("sd15")

    def build(seed):
        u = unet_ref.init_synthetic_(unet_ref.UNet2DConditionModel(oc), seed=1)
        adopt_unet_transformers(u)
        torch.manual_seed(seed)
        net = LoRASpecialNetwork(None, u, lora_dim=4, alpha=2, train_text_encoder=False)
        net.force_to("cpu", torch.float32)
        net._update_torch_multiplier()
        net.apply_to(None, u, False, True)
        with torch.no_grad():
            for lora in net.unet_loras:
                lora.lora_up.weight.normal_(0, 0.02)
        opt = B200AdamW(net, lr=1e-4)
        return host.UNetLoRATrainStep(u, net, opt, prediction_type="epsilon"), net

    step, net = build(0)
    path = step.save(str(tmp_path), "unet_lora", step=7)
    keys = load_file(path)
    assert "lora_unet_down_blocks_1_attentions_0_proj_in.lora_down.weight" in keys
    assert "lora_unet_mid_block_attentions_0_transformer_blocks_0_attn2_to_k.alpha" in keys
    step2, net2 = build(1)
    assert not torch.equal(net2.flat_params, net.flat_params)
    out = step2.resume(str(tmp_path), "unet_lora")
    assert out[1] == 7
    # saved in fp16 (save.dtype default of the reference): equal after the same rounding
    assert torch.equal(net2.flat_params.half(), net.flat_params.half())


# ------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_row_kernels_vs_torch():
    from ai_toolkit_b200 import ops
    torch.manual_seed(0)
    # LayerNorm affine at a width that is not a multiple of 256
    M, D = 300, 640
    x = torch.randn(M, D, device=DEV).bfloat16()
    w = (1 + 0.1 * torch.randn(D, device=DEV)).bfloat16()
    b = (0.1 * torch.randn(D, device=DEV)).bfloat16()
    y, mean, rstd = ops.ln_affine_fwd(x, w, b, 1e-5)
    xr = x.float().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), w.float(), b.float(), 1e-5)
    assert _rel(y, yr) < 4e-3
    dy = torch.randn(M, D, device=DEV).bfloat16()
    dres = torch.randn(M, D, device=DEV).bfloat16()
    yr.backward(dy.float())
    assert _rel(ops.ln_affine_bwd(dy, x, mean, rstd, w, dres=dres), xr.grad + dres.float()) < 5e-3
    # GroupNorm (with and without SiLU)
    for silu in (False, True):
        B, C, H, W = 2, 320, 12, 10
        x = torch.randn(B, C, H, W, device=DEV).bfloat16()
        w = (1 + 0.1 * torch.randn(C, device=DEV)).bfloat16()
        b = (0.1 * torch.randn(C, device=DEV)).bfloat16()
        y, mean, rstd = ops.groupnorm_fwd(x, w, b, 32, 1e-6, silu=silu)
        xr = x.float().requires_grad_(True)
        yr = F.group_norm(xr, 32, w.float(), b.float(), 1e-6)
        yr = F.silu(yr) if silu else yr
        assert _rel(y, yr) < 5e-3
        dy = torch.randn_like(x)
        yr.backward(dy.float())
        assert _rel(ops.groupnorm_bwd(dy, x, w, b, mean, rstd, 32, silu=silu), xr.grad) < 8e-3
    # GEGLU
    M, Fh = 200, 1280
    proj = torch.randn(M, 2 * Fh, device=DEV).bfloat16()
    pr = proj.float().requires_grad_(True)
    h, g = pr.chunk(2, dim=-1)
    yr = h * F.gelu(g)
    y = ops.geglu_fwd(proj)
    assert _rel(y, yr) < 5e-3
    dy = torch.randn(M, Fh, device=DEV).bfloat16()
    yr.backward(dy.float())
    assert _rel(ops.geglu_bwd(dy, proj), pr.grad) < 6e-3
    # zero-padded head re-layout
    B, L, H, d = 2, 37, 5, 64
    x = torch.randn(B * L, H * d + 16, device=DEV).bfloat16()
    hm = torch.full((B, H, L, 128), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.heads_pad(x[:, :H * d], hm, B, L, d)
    want = x[:, :H * d].view(B, L, H, d).transpose(1, 2)
    assert torch.equal(hm[..., :d], want) and float(hm[..., d:].abs().sum()) == 0.0
    back = torch.zeros(B * L, H * d, device=DEV, dtype=torch.bfloat16)
    ops.heads_unpad(hm, back, B, L, d)
    assert torch.equal(back, x[:, :H * d])
    # three tensors of one (cross) attention in one launch: q has L tokens, k / v have Lk
    Lk = 11
    qkv = torch.randn(B * L, 3 * H * d, device=DEV).bfloat16()
    kv = torch.randn(B * Lk, 2 * H * d, device=DEV).bfloat16()
    Q = torch.full((B, H, L, 128), float("nan"), device=DEV, dtype=torch.bfloat16)
    K = torch.full((B, H, Lk, 128), float("nan"), device=DEV, dtype=torch.bfloat16)
    V = torch.full((B, H, Lk, 128), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.heads_pad_multi([(qkv[:, H * d:2 * H * d], Q), (kv[:, :H * d], K), (kv[:, H * d:], V)], B, d)
    for hm_, tm_, L_ in ((Q, qkv[:, H * d:2 * H * d], L), (K, kv[:, :H * d], Lk), (V, kv[:, H * d:], Lk)):
        assert torch.equal(hm_[..., :d], tm_.reshape(B, L_, H, d).transpose(1, 2)) and float(hm_[..., d:].abs().sum()) == 0.0
    dq = torch.zeros(B * L, H * d, device=DEV, dtype=torch.bfloat16)
    dkv = torch.zeros(B * Lk, 2 * H * d, device=DEV, dtype=torch.bfloat16)
    ops.heads_pad_multi([(dq, Q), (dkv[:, :H * d], K), (dkv[:, H * d:], V)], B, d, to_heads=False)
    assert torch.equal(dq, qkv[:, H * d:2 * H * d]) and torch.equal(dkv, kv)


def _setup(linear, heads, dim_head, layers, cross_dim, rank, alpha, seed=0):
    C = heads * dim_head
    torch.manual_seed(seed)
    o = unet_ref.init_synthetic_(unet_ref.Transformer2DModel(heads, dim_head, C, layers, cross_dim, linear), seed=seed, std=0.05)
    o.requires_grad_(False)
    root = type("UNet2DConditionModel", (torch.nn.Module,), {})()
    torch.nn.Module.__init__(root)
    root.t2d = Transformer2DModel(heads, dim_head, C, layers, cross_dim, linear, device=DEV)
    root.t2d.load_state_dict(o.state_dict(), strict=True)
    net = LoRASpecialNetwork(None, root, lora_dim=rank, alpha=alpha, train_text_encoder=False)
    net.force_to(DEV, torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, root, False, True)
    g = torch.Generator().manual_seed(seed + 1)
    refs = {}
    for name, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        refs[name] = copy.deepcopy(o).to(DEV, dt)
    with torch.no_grad():
        for lora in net.unet_loras:
            lora.lora_up.weight.copy_(torch.randn(lora.lora_up.weight.shape, generator=g) * 0.05)
    net.mark_params_changed()
    return root.t2d, net, refs, C


def _attach_eager_adapters(om, net):
    """Hooks on the oracle's layers applying `org + bf16(scale * up(down(x.float())))` with leaf copies of the adapter weights."""
    leaves = []
    name_to_mod = {("lora_unet_t2d_" + n.replace(".", "_")): m for n, m in om.named_modules()}
    for lora in net.unet_loras:
        mod = name_to_mod[lora.lora_name]
        A = lora.lora_down.weight.detach().clone().requires_grad_(True)
        Bw = lora.lora_up.weight.detach().clone().requires_grad_(True)
        leaves.append((A, Bw))

        def hook(m, inp, out, A=A, Bw=Bw, s=lora.scale):
            x = inp[0].float()
            lx = F.conv2d(F.conv2d(x, A), Bw) if isinstance(m, torch.nn.Conv2d) else F.linear(F.linear(x, A), Bw)
            return out + (lx * s).to(out.dtype)

        mod.register_forward_hook(hook)
    return leaves


@pytest.mark.gpu
@pytest.mark.parametrize("linear,heads,dim_head,layers,cross_dim,B,H,W,Lc,rank", [
    (False, 4, 40, 1, 96, 2, 16, 12, 77, 4),    # SD1.5 form: 1x1-conv projections, head dim 40 (320 / 8)
    (False, 2, 80, 1, 96, 1, 8, 8, 77, 8),      # SD1.5 second level: head dim 80
    (True, 5, 64, 2, 256, 2, 16, 16, 77, 8),    # SDXL form: Linear projections, head dim 64, depth 2
    (False, 2, 160, 1, 96, 2, 16, 16, 77, 4),   # SD1.5 deepest levels: head dim 160 -> CUDA-core attention kernel (256 tokens)
    (False, 2, 160, 1, 96, 1, 6, 6, 77, 8),     # ... 36 tokens: ragged tiles
])
def test_transformer2d_engine_matches_oracle(linear, heads, dim_head, layers, cross_dim, B, H, W, Lc, rank):
    model, net, refs, C = _setup(linear, heads, dim_head, layers, cross_dim, rank, rank / 2)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, C, H, W, generator=g).bfloat16().to(DEV)
    ctx = torch.randn(B, Lc, cross_dim, generator=g).bfloat16().to(DEV)
    dout = torch.randn(B, C, H, W, generator=g).bfloat16().to(DEV)
    res = {}
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        om = refs[name]
        leaves = _attach_eager_adapters(om, net)
        xr = x.detach().to(dt).clone().requires_grad_(True)
        y = om(xr, ctx.to(dt))
        y.backward(dout.to(dt))
        res[name] = (y.detach(), xr.grad, torch.cat([p.grad.reshape(-1) for ab in leaves for p in ab]))
    xm = x.detach().clone().requires_grad_(True)
    net.flat_grads.zero_()
    with net:
        y = model(xm, ctx)
        y.backward(dout)
    gm = net.flat_grads[:res["fp32"][2].numel()]
    fl_y, fl_dx, fl_g = (_rel(res["bf16"][i], res["fp32"][i]) for i in range(3))
    e_y, e_dx, e_g = _rel(y, res["fp32"][0]), _rel(xm.grad, res["fp32"][1]), _rel(gm, res["fp32"][2])
    print(f"[t2d linear={linear} d={dim_head}] y {e_y:.3e} (floor {fl_y:.3e}); dx {e_dx:.3e} (floor {fl_dx:.3e}); dA/dB {e_g:.3e} (floor {fl_g:.3e})")
    assert e_y < max(1e-3, 1.5 * fl_y) and e_dx < max(1e-3, 1.5 * fl_dx) and e_g < max(1e-3, 1.5 * fl_g)
    # inactive network: the frozen block
    with torch.no_grad():
        y0 = model(x, ctx)
        y0_ref = copy.deepcopy(unet_ref.Transformer2DModel(heads, dim_head, C, layers, cross_dim, linear)).to(DEV, torch.float32)
        y0_ref.load_state_dict({k: v.float() for k, v in model.state_dict().items()})
        assert _rel(y0, y0_ref(x.float(), ctx.float())) < 1.5e-2


# ------------------------------------------------------------------------------------- the UNet host (ai_toolkit_b200/unet.py)
def _tiny_cfgs(form="sdxl"):
    from ai_toolkit_b200 import unet as host
    if form == "sd15":  # 1x1-conv projections, no text-time embedding, head dims 64 and 160 (the CUDA-core attention kernel)
        kw = dict(block_out_channels=(64, 320), attn_layers=(1, 1), heads=(1, 2), cross_attention_dim=96, use_linear_projection=False)
    else:
        kw = dict(block_out_channels=(64, 128), attn_layers=(1, 2), heads=(1, 2), cross_attention_dim=96, use_linear_projection=True,
                  addition_embed=True, addition_time_embed_dim=16, projection_class_embeddings_input_dim=32 + 6 * 16)
    return host.UNetConfig(**kw), unet_ref.UNetConfig(**kw)


def test_unet_host_matches_oracle_parameter_names():
    from ai_toolkit_b200 import unet as host
    hc, oc = _tiny_cfgs()
    for a, b in ((hc, oc), (host.sd15_config(), unet_ref.sd15_config()), (host.sdxl_config(), unet_ref.sdxl_config())):
        with torch.device("meta"):
            so = unet_ref.UNet2DConditionModel(b).state_dict()
        sm = host.UNet2DConditionModel(a, device="meta").state_dict()
        assert list(so.keys()) == list(sm.keys())
        assert all(so[k].shape == sm[k].shape for k in so)
    n_sdxl = sum(v.numel() for v in sm.values())
    assert abs(n_sdxl - 2.567e9) < 0.01e9  # SDXL-base UNet: 2.57 B parameters
    n_sd15 = sum(v.numel() for v in host.UNet2DConditionModel(host.sd15_config(), device="meta").state_dict().values())
    assert abs(n_sd15 - 859.5e6) < 1e6  # SD1.5 UNet: 860 M parameters
    # the FLOP figure bench.py reports against (FlopCounterMode over the oracle forward + backward)
    from torch.utils.flop_counter import FlopCounterMode
    with torch.device("meta"):
        om = unet_ref.UNet2DConditionModel(unet_ref.sdxl_config()).requires_grad_(False)
        x = torch.randn(2, 4, 128, 128, requires_grad=True)
        with FlopCounterMode(display=False) as fc:
            y = om(x, torch.ones(2), torch.randn(2, 77, 2048),
                   added_cond_kwargs={"text_embeds": torch.randn(2, 1280), "time_ids": torch.randn(2, 6)})[0]
            f_fwd = fc.get_total_flops()
            y.backward(torch.randn_like(y))
        f_step = fc.get_total_flops()
    assert abs(f_step / 2 - host.SDXL_STEP_FLOPS_PER_SAMPLE) < 1e6
    _, lin, att = host.unet_flops(host.sdxl_config(), 2, 128, 128)
    assert abs((lin + att) - f_fwd) / f_fwd < 1e-4  # the closed form restates the counter's forward


@pytest.mark.gpu
@pytest.mark.parametrize("form", ["sdxl", "sd15"])
def test_unet_host_step_matches_oracle(form):
    """Tiny SDXL-form / SD1.5-form UNet: prediction, loss and every LoRA gradient of `UNetLoRATrainStep` vs the fp32 oracle UNet with eager
    adapters and the oracle's `calculate_loss`; then one optimizer step moves the flat parameters."""
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200 import unet as host
    hc, oc = _tiny_cfgs(form)
    o = unet_ref.init_synthetic_(unet_ref.UNet2DConditionModel(oc), seed=3, std=0.05)
    o.requires_grad_(False)
    m = host.UNet2DConditionModel(hc, device=DEV)
    m.load_state_dict(o.state_dict(), strict=True)
    rank = 8
    net = LoRASpecialNetwork(None, m, lora_dim=rank, alpha=4, train_text_encoder=False)
    net.force_to(DEV, torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, m, False, True)
    g = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for lora in net.unet_loras:
            lora.lora_up.weight.copy_(torch.randn(lora.lora_up.weight.shape, generator=g) * 0.05)
    net.mark_params_changed()
    # 11 Transformer2DModels (2 + 2 down, 1 mid, 3 + 3 up) x proj_in/out + their BasicTransformerBlocks x 10 Linears
    n_blocks = (2 * 1 + 2 * 2 + 1 * 2 + 3 * 2 + 3 * 1) if form == "sdxl" else 11
    assert len(net.unet_loras) == 2 * 11 + 10 * n_blocks
    B, H, W = 2, 16, 16
    lat = torch.randn(B, 4, H, W, generator=g).bfloat16().to(DEV)
    noise = torch.randn(B, 4, H, W, generator=g).bfloat16().to(DEV)
    text = torch.randn(B, 77, 96, generator=g).bfloat16().to(DEV)
    pooled = torch.randn(B, 32, generator=g).bfloat16().to(DEV)
    ts = torch.tensor([37, 801], device=DEV, dtype=torch.int64)
    opt = B200AdamW(net, lr=1e-3)
    step = host.UNetLoRATrainStep(m, net, opt, prediction_type="epsilon")
    p0 = net.flat_params.clone()

    res = {}
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        om = copy.deepcopy(o).to(DEV, dt)
        leaves = []
        name_to_mod = {("lora_unet_" + n.replace(".", "_")): mod for n, mod in om.named_modules()}
        for lora in net.unet_loras:
            mod = name_to_mod[lora.lora_name]
            A = lora.lora_down.weight.detach().clone().requires_grad_(True)
            Bw = lora.lora_up.weight.detach().clone().requires_grad_(True)
            leaves.append((A, Bw))

            def hook(mm, inp, out, A=A, Bw=Bw, s=lora.scale):
                x_in = inp[0].float()
                lx = F.conv2d(F.conv2d(x_in, A), Bw) if isinstance(mm, torch.nn.Conv2d) else F.linear(F.linear(x_in, A), Bw)
                return out + (lx * s).to(out.dtype)

            mod.register_forward_hook(hook)
        # DDPMScheduler.add_noise on the table of the reference's training scheduler; the noisy latents are bf16 (the train dtype)
        ac = step.table.alphas_cumprod.to(DEV)[ts].float()
        noisy = (ac.sqrt()[:, None, None, None] * lat.float() + (1 - ac).sqrt()[:, None, None, None] * noise.float())
        noisy = noisy.to(torch.bfloat16).to(dt)
        pred = om(noisy, ts.float(), text.to(dt), added_cond_kwargs={"text_embeds": pooled.to(dt), "time_ids": step.time_ids(B, H, W)}
                  if form == "sdxl" else None)[0]
        loss = ((pred.float() - noise.float()) ** 2).mean()
        loss.backward()
        res[name] = (loss.detach(), torch.cat([p.grad.reshape(-1) for ab in leaves for p in ab]))
    tot = step.run(lat, noise, ts, text, pooled if form == "sdxl" else None)
    gm = net.flat_grads[:res["fp32"][1].numel()]
    # (run() already stepped the optimizer; the gradients are still in the flat buffer until the next zero_grad)
    e_l = abs(tot.item() - res["fp32"][0].item()) / res["fp32"][0].item()
    fl_l = abs(res["bf16"][0].item() - res["fp32"][0].item()) / res["fp32"][0].item()
    e_g, fl_g = _rel(gm, res["fp32"][1]), _rel(res["bf16"][1], res["fp32"][1])
    print(f"[unet host {form}] loss {tot.item():.6f} vs {res['fp32'][0].item():.6f} rel {e_l:.3e} (floor {fl_l:.3e}); dA/dB {e_g:.3e} (floor {fl_g:.3e})")
    assert e_l < max(2e-3, 1.5 * fl_l) and e_g < max(2e-3, 1.5 * fl_g)
    assert not torch.equal(net.flat_params, p0)
    batch = dict(latents=lat, noise=noise, timesteps=ts.cpu(), text_embeds=text)
    if form == "sdxl":
        batch["pooled_embeds"] = pooled
    out = step.hook_train_loop(batch)
    assert out["loss"] > 0 and out["loss"] == out["loss"]


@pytest.mark.gpu
def test_unet_host_step_cuda_graph_equals_eager():
    """The whole-step CUDA graph (eager frozen body + engine blocks + autograd backward + optimizer) replays the eager step:
    same losses and same flat parameters after 5 steps on changing batches (fp32 atomics in the wgrads: tolerance, not bits)."""
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200 import unet as host
    hc, _ = _tiny_cfgs()
    runs = {}
    for graph in (False, True):
        m = host.UNet2DConditionModel(hc, device=DEV).init_synthetic_(seed=5, std=0.05)
        net = LoRASpecialNetwork(None, m, lora_dim=4, alpha=4, train_text_encoder=False)
        net.force_to(DEV, torch.float32)
        net._update_torch_multiplier()
        net.apply_to(None, m, False, True)
        g = torch.Generator().manual_seed(6)
        with torch.no_grad():
            for lora in net.unet_loras:
                lora.lora_down.weight.copy_(torch.randn(lora.lora_down.weight.shape, generator=g) * 0.05)
                lora.lora_up.weight.copy_(torch.randn(lora.lora_up.weight.shape, generator=g) * 0.05)
        net.mark_params_changed()
        opt = B200AdamW(net, lr=1e-3, max_grad_norm=1.0)
        step = host.UNetLoRATrainStep(m, net, opt, prediction_type="v_prediction", min_snr_gamma=5.0, use_cuda_graph=graph)
        losses = []
        for i in range(5):
            lat = torch.randn(2, 4, 16, 16, generator=g).bfloat16()
            noise = torch.randn(2, 4, 16, 16, generator=g).bfloat16()
            text = torch.randn(2, 77, 96, generator=g).bfloat16()
            pooled = torch.randn(2, 32, generator=g).bfloat16()
            ts = torch.randint(1, 999, (2,), generator=g)
            losses.append(step.hook_train_loop(dict(latents=lat, noise=noise, timesteps=ts, text_embeds=text, pooled_embeds=pooled))["loss"])
        assert (step._graph is not None) == graph
        runs[graph] = (losses, net.flat_params.clone())
    le, lg = runs[False][0], runs[True][0]
    print("[unet graph] eager", le, "graph", lg)
    assert all(abs(a - b) / abs(a) < 2e-3 for a, b in zip(le, lg))
    assert _rel(runs[True][1], runs[False][1]) < 2e-3


@pytest.mark.gpu
def test_adopted_eager_unet_trains_like_the_host_unet():
    """The plugin's route for SD1.5 / SDXL (`hook_after_model_load` -> `adopt_unet_transformers`): a loaded eager UNet whose
    Transformer2DModels were adopted, stepped by `UNetLoRATrainStep` (gradient accumulation over two micro-batches, then one
    whole step), against this package's host UNet with the same weights and adapters."""
    from ai_toolkit_b200 import unet as host
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.unet_blocks import adopt_unet_transformers
    hc, oc = _tiny_cfgs("sdxl")
    o = unet_ref.init_synthetic_(unet_ref.UNet2DConditionModel(oc), seed=3, std=0.05).requires_grad_(False)
    runs = {}
    for kind in ("host", "adopted"):
        if kind == "host":
            m = host.UNet2DConditionModel(hc, device=DEV)
            m.load_state_dict(o.state_dict(), strict=True)
        else:
            m = copy.deepcopy(o).to(DEV, torch.bfloat16)
            assert adopt_unet_transformers(m) == 11
        net = LoRASpecialNetwork(None, m, lora_dim=8, alpha=4, train_text_encoder=False)
        net.force_to(DEV, torch.float32)
        net._update_torch_multiplier()
        net.apply_to(None, m, False, True)
        g = torch.Generator().manual_seed(4)
        with torch.no_grad():
            for lora in net.unet_loras:
                lora.lora_down.weight.copy_(torch.randn(lora.lora_down.weight.shape, generator=g) * 0.05)
                lora.lora_up.weight.copy_(torch.randn(lora.lora_up.weight.shape, generator=g) * 0.05)
        net.mark_params_changed()
        opt = B200AdamW(net, lr=1e-3, max_grad_norm=1.0)
        step = host.UNetLoRATrainStep(m, net, opt, prediction_type="epsilon", use_cuda_graph=False)
        losses = []
        for i in range(3):
            lat = torch.randn(2, 4, 16, 16, generator=g).bfloat16()
            noise = torch.randn(2, 4, 16, 16, generator=g).bfloat16()
            text = torch.randn(2, 77, 96, generator=g).bfloat16()
            pooled = torch.randn(2, 32, generator=g).bfloat16()
            ts = torch.randint(1, 999, (2,), generator=g).float()  # the trainer passes float timesteps
            first, last = (i != 1), (i != 0)                      # micro-batches 0 + 1 accumulate, batch 2 is a whole step
            losses.append(float(step.run(lat, noise, ts, text, pooled, first_micro_batch=first, last_micro_batch=last).item()))
        runs[kind] = (losses, net.flat_params.clone())
    print("[adopted unet] host", runs["host"][0], "adopted", runs["adopted"][0])
    assert all(abs(a - b) / abs(a) < 2e-3 for a, b in zip(runs["host"][0], runs["adopted"][0]))
    assert _rel(runs["adopted"][1], runs["host"][1]) < 2e-3
