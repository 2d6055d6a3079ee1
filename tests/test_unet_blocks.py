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
    with pytest.raises(NotImplementedError):
        Transformer2DModel(8, 160, 1280)  # SD1.5's deepest level


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
