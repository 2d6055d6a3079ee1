"""csrc/batch_ops.cu against plain torch restatements of the reference arithmetic (SURVEY.md section 8 rows a5', a10, a12):
DDPM add_noise (bit-exact vs the bf16 tensor expression), the general training loss (flow / eps / v / given target,
sample weights = timestep weights x loss multiplier x SNR weights, mask), NCHW <-> rows, im2col / col2im, and the Conv2d
LoRA adapters (1x1 and k x k, strided) through `LoRAModule.forward` against the eager reference formula
(toolkit/network_mixins.py:304-342 with toolkit/lora_special.py:95-104 convs)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def test_ddpm_add_noise_bit_exact_vs_bf16_tensor_expression():
    from ai_toolkit_b200 import ops
    from ai_toolkit_b200.samplers import DDPMTable
    tab = DDPMTable(device=DEV)
    g = torch.Generator().manual_seed(0)
    for shape in ((2, 4, 128, 128), (3, 4, 64, 64), (1, 16, 6, 10)):
        x = torch.randn(shape, generator=g).bfloat16().to(DEV)
        n = torch.randn(shape, generator=g).bfloat16().to(DEV)
        t = tab.sample_timesteps(shape[0], generator=g).to(DEV)
        got = ops.ddpm_add_noise(x, n, t, tab.device_table)
        want = tab.add_noise(x, n, t)  # bf16 tensor arithmetic, exactly as DDPMScheduler.add_noise evaluates it
        assert want.dtype == torch.bfloat16
        assert torch.equal(got, want)
    # edge timesteps of the table
    t = torch.tensor([0, 999], device=DEV)
    x, n = torch.randn(2, 4, 8, 8, device=DEV).bfloat16(), torch.randn(2, 4, 8, 8, device=DEV).bfloat16()
    assert torch.equal(ops.ddpm_add_noise(x, n, t, tab.device_table), tab.add_noise(x, n, t))


def _ref_loss(pred, lat, noise, *, kind, tab=None, t=None, target=None, w=None, mask=None):
    """SDTrainer.calculate_loss default path restated with torch ops (SDTrainer.py:619-650, :916, :923-959, :987-1013)."""
    pred = pred.detach().clone().requires_grad_(True)
    if target is not None:
        tg = target
    elif kind == "flow":
        tg = (noise - lat).detach()
    elif kind == "eps":
        tg = noise
    else:
        tg = tab.get_velocity(lat, noise, t)
    loss = F.mse_loss(pred.float(), tg.float(), reduction="none")
    if w is not None and w[0] is not None:  # timestep weights, before the mean
        loss = loss * w[0].view(-1, 1, 1, 1)
    if mask is not None:
        loss = loss * mask
    loss = loss.mean([1, 2, 3])
    if w is not None and w[1] is not None:  # loss_multiplier and SNR weights, after the mean
        loss = loss * w[1]
    total = loss.mean()
    total.backward()
    return total.detach(), loss.detach(), pred.grad


@pytest.mark.parametrize("kind", ["flow", "eps", "v", "given"])
@pytest.mark.parametrize("with_mask", [False, True])
def test_train_loss_matches_calculate_loss(kind, with_mask):
    from ai_toolkit_b200 import ops
    from ai_toolkit_b200.samplers import DDPMTable
    torch.manual_seed(1)
    B, C, H, W = 3, 4, 32, 48
    lat = torch.randn(B, C, H, W, device=DEV).bfloat16()
    noise = torch.randn(B, C, H, W, device=DEV).bfloat16()
    pred = torch.randn(B, C, H, W, device=DEV).bfloat16()
    tab = DDPMTable(prediction_type="v_prediction" if kind == "v" else "epsilon", device=DEV)
    t = torch.tensor([5, 400, 990], device=DEV)
    tw = torch.tensor([0.5, 1.0, 2.0], device=DEV)
    after = torch.tensor([1.0, 0.25, 3.0], device=DEV) * tab.snr_weights(t, 5.0).to(DEV)
    mask = None
    if with_mask:
        mask = (torch.rand(B, 1, H, W, device=DEV) > 0.3).float() * 1.5
    kw = {}
    if kind == "eps" or kind == "v":
        cn, cl = tab.target_coefficients(t)
        kw = dict(coef_noise=cn.to(DEV), coef_latent=cl.to(DEV))
    target = None
    if kind == "given":
        target = torch.randn(B, C, H, W, device=DEV).bfloat16()
        kw = dict(target=target)
    tot, per, dpred = ops.train_loss(pred, lat, noise, sample_weight=(tw * after).contiguous(), mask=mask, pack=False,
                                     gscale=1.0, **kw)
    rt, rper, rg = _ref_loss(pred, lat, noise, kind=kind, tab=tab, t=t, target=target, w=(tw, after), mask=mask)
    assert abs(tot.item() - rt.item()) <= 2e-6 * abs(rt.item()) + 1e-7
    torch.testing.assert_close(per, rper, rtol=5e-6, atol=1e-7)
    assert _rel(dpred, rg) < 4e-3  # one bf16 rounding of the stored gradient
    # 5-D video latents fold to 4-D
    if kind == "flow" and not with_mask:
        l5, n5, p5 = (x.view(B, C, 4, 8, W) for x in (lat, noise, pred))
        tot5, _, _ = ops.train_loss(p5, l5, n5, sample_weight=(tw * after).contiguous())
        assert abs(tot5.item() - tot.item()) < 1e-7 + 1e-6 * abs(tot.item())


def test_flux_packed_train_loss_equals_flow_loss():
    from ai_toolkit_b200 import ops
    torch.manual_seed(2)
    B, C, H, W = 2, 16, 16, 24
    lat = torch.randn(B, C, H, W, device=DEV).bfloat16()
    noise = torch.randn(B, C, H, W, device=DEV).bfloat16()
    pred = torch.randn(B, (H // 2) * (W // 2), C * 4, device=DEV).bfloat16()
    a = ops.flow_loss(pred, lat, noise, pack=True, gscale=0.5)
    b = ops.train_loss(pred, lat, noise, pack=True, gscale=0.5)
    assert torch.equal(a[2], b[2])  # same arithmetic per element: the gradients are bit-identical
    torch.testing.assert_close(a[0], b[0], rtol=1e-6, atol=0)  # the sums are fp32 atomics (order-dependent last bits)
    torch.testing.assert_close(a[1], b[1], rtol=1e-6, atol=0)


@pytest.mark.parametrize("B,C,H,W", [(2, 320, 16, 24), (1, 8, 5, 7), (3, 40, 33, 31)])
def test_nchw_rows_roundtrip(B, C, H, W):
    from ai_toolkit_b200 import ops
    x = torch.randn(B, C, H, W, device=DEV).bfloat16()
    rows = ops.nchw_to_rows(x)
    assert torch.equal(rows, x.permute(0, 2, 3, 1).reshape(B * H * W, C))
    assert torch.equal(ops.rows_to_nchw(rows, B, C, H, W), x)


@pytest.mark.parametrize("k,s,p", [((3, 3), (1, 1), (1, 1)), ((3, 3), (2, 2), (1, 1)), ((1, 1), (2, 2), (0, 0)),
                                   ((5, 3), (1, 2), (2, 0))])
def test_im2col_col2im_vs_unfold_fold(k, s, p):
    from ai_toolkit_b200 import ops
    B, C, H, W = 2, 16, 13, 18
    x = torch.randn(B, C, H, W, device=DEV).bfloat16()
    cols = ops.im2col(x, k, s, p)
    want = F.unfold(x.float(), k, padding=p, stride=s)  # [B, C kh kw, L] with (c, ky, kx) ordering
    Ho, Wo = ops.conv_out_hw(H, W, k, s, p)
    want = want.transpose(1, 2).reshape(B * Ho * Wo, -1)
    assert torch.equal(cols[:, :want.shape[1]].float(), want)  # pure index op: bit-exact
    d = torch.randn_like(cols)
    dx = ops.col2im(d, (B, C, H, W), k, s, p)
    ref = F.fold(d[:, :want.shape[1]].float().view(B, Ho * Wo, -1).transpose(1, 2), (H, W), k, padding=p, stride=s)
    assert _rel(dx, ref) < 4e-3
    dx2 = ops.col2im(d, (B, C, H, W), k, s, p, out=dx.clone(), accumulate=True)
    assert _rel(dx2, ref + dx.float()) < 6e-3


class _Res(torch.nn.Module):
    def __init__(self, cin, cout, k, s, p, bias=True):
        super().__init__()
        self.conv = torch.nn.Conv2d(cin, cout, k, s, p, bias=bias)

    def forward(self, x):
        return self.conv(x)


@pytest.mark.parametrize("cin,cout,k,s,p,rank", [(32, 64, 1, 1, 0, 4), (64, 64, 3, 1, 1, 8), (64, 128, 3, 2, 1, 4)])
def test_conv_lora_module_matches_reference_formula(cin, cout, k, s, p, rank):
    """y = conv(x) + bf16(m s up(down(x.float())))  (network_mixins.py:304-342) and its gradients w.r.t. the adapter weights
    and the input, for 1x1 and k x k (strided) Conv2d adapters."""
    from ai_toolkit_b200 import LoRASpecialNetwork
    torch.manual_seed(3)
    _Res.__name__ = "ResnetBlock2D"
    root = type("UNet2DConditionModel", (torch.nn.Module,), {})()
    torch.nn.Module.__init__(root)
    root.block = _Res(cin, cout, k, s, p)
    root = root.to(DEV, torch.bfloat16).requires_grad_(False)
    net = LoRASpecialNetwork(None, root, lora_dim=rank, alpha=rank / 2, conv_lora_dim=rank, conv_alpha=rank / 2,
                             train_text_encoder=False, target_lin_modules=["ResnetBlock2D"], target_conv_modules=["ResnetBlock2D"])
    net.force_to(DEV, torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, root, False, True)
    assert len(net.unet_loras) == 1
    lora = net.unet_loras[0]
    with torch.no_grad():
        lora.lora_up.weight.normal_(0, 0.05)
    net.mark_params_changed()
    net.multiplier = [1.0, 0.5]
    B, H, W = 2, 20, 28
    x = torch.randn(B, cin, H, W, device=DEV).bfloat16().requires_grad_(True)
    net.flat_grads.zero_()
    with net:
        y = root.block(x)
        go = torch.randn_like(y)
        y.backward(go)
    dx = x.grad.clone()
    # eager reference
    conv = root.block.conv
    xr = x.detach().clone().requires_grad_(True)
    A = lora.lora_down.weight.detach().clone().requires_grad_(True)
    Bw = lora.lora_up.weight.detach().clone().requires_grad_(True)
    org = lora.org_forward(xr)
    lx = F.conv2d(F.conv2d(xr.float(), A, None, conv.stride, conv.padding), Bw) * lora.scale
    m = torch.tensor([1.0, 0.5], device=DEV).view(2, 1, 1, 1)
    yr = org + (lx * m).to(torch.bfloat16)
    yr.backward(go)
    assert y.shape == yr.shape
    assert _rel(y, yr) < 1e-2
    assert _rel(dx, xr.grad) < 1.5e-2
    assert _rel(lora.lora_down.weight.grad, A.grad) < 1.5e-2
    assert _rel(lora.lora_up.weight.grad, Bw.grad) < 1.5e-2
    # inactive network: the frozen conv, untouched
    assert torch.equal(root.block(x.detach()), conv(x.detach()))


def test_dropout_variants_on_the_module_seam():
    """dropout + rank_dropout in training mode: same generator consumption as `_call_forward` (F.dropout on the rank-side
    activation, then torch.rand((B, r))), so with the same CUDA seed the fused path reproduces the eager formula."""
    from ai_toolkit_b200 import LoRASpecialNetwork
    torch.manual_seed(4)

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.transformer_blocks = torch.nn.ModuleList([torch.nn.Linear(256, 384)])

    Toy.__name__ = "FluxTransformer2DModel"
    m = Toy().to(DEV, torch.bfloat16).requires_grad_(False)
    net = LoRASpecialNetwork(None, m, lora_dim=8, alpha=8, train_text_encoder=False, is_flux=True, transformer_only=True,
                             dropout=0.25, rank_dropout=0.5)
    net.force_to(DEV, torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, m, False, True)
    lora = net.unet_loras[0]
    with torch.no_grad():
        lora.lora_up.weight.normal_(0, 0.05)
    net.mark_params_changed()
    net.train()
    B, L = 2, 160
    x = torch.randn(B, L, 256, device=DEV).bfloat16()
    lin = m.transformer_blocks[0]
    torch.cuda.manual_seed(11)
    net.flat_grads.zero_()
    with net:
        y = lin(x)
        (y.float() ** 2).mean().backward()
    # eager formula with the same generator state
    torch.cuda.manual_seed(11)
    A = lora.lora_down.weight.detach().clone().requires_grad_(True)
    Bw = lora.lora_up.weight.detach().clone().requires_grad_(True)
    lx = F.linear(x.float(), A)
    lx = F.dropout(lx.view(B * L, 8), p=0.25).view(B, L, 8)
    mask = torch.rand((B, 8), device=DEV) > 0.5
    lx = lx * mask.unsqueeze(1)
    yr = lora.org_forward(x) + (F.linear(lx, Bw) * (lora.scale * 2.0)).to(torch.bfloat16)
    (yr.float() ** 2).mean().backward()
    assert _rel(y, yr) < 1e-2
    assert _rel(lora.lora_down.weight.grad, A.grad) < 2e-2 and _rel(lora.lora_up.weight.grad, Bw.grad) < 2e-2
    net.eval()
    with net, torch.no_grad():
        y_eval = lin(x)
    ye = lora.org_forward(x) + (F.linear(F.linear(x.float(), A), Bw) * lora.scale).to(torch.bfloat16)
    assert _rel(y_eval, ye) < 1e-2


def test_per_sample_multipliers_through_the_fused_engine():
    """network.multiplier = [m_0, m_1] (SDTrainer.py:1558): every adapter, including the AdaLN projections of the
    conditioning vector, scales sample b by m_b (network_mixins.py:311-322) -- engine vs oracle."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_gpu_flux_engine import _oracle_step, _setup
    from ai_toolkit_b200 import ops
    from ai_toolkit_b200.train_step import make_img_ids
    B, hl, wl, Lt = 2, 16, 16, 24
    model, net, onets, batch = _setup(1, 1, 2, B, hl, wl, Lt, 8, seed=21)
    lat, noise, t, text, pooled = batch
    mult = [1.0, -0.5]
    net.multiplier = mult
    for om, on in onets.values():
        on.torch_multiplier = torch.tensor(mult)  # what `_update_torch_multiplier` builds from the list (:791-845)
    loss32, pred32, g32 = _oracle_step(*onets["fp32"], batch, torch.float32)
    loss16, pred16, g16 = _oracle_step(*onets["bf16"], batch, torch.bfloat16)
    packed = ops.flow_add_noise(lat, noise, t, pack=True)
    net.flat_grads.zero_()
    with net:
        pred = model.engine.forward(packed, t, text, pooled, torch.ones(B, device=DEV), torch.zeros(Lt, 3, device=DEV),
                                    make_img_ids(hl, wl, DEV), save=True, t_div=1000.0)
        tot, _, dpred = ops.flow_loss(pred.view(B, -1, 64), lat, noise, pack=True)
        model.engine.backward(dpred.view(-1, 64))
    g = net.flat_grads[:g32.numel()]
    floor = _rel(g16, g32)
    assert abs(tot.item() - loss32) / abs(loss32) < max(1e-3, 1.5 * abs(loss16 - loss32) / abs(loss32))
    assert _rel(g, g32) < max(1e-3, 1.5 * floor)
    # and it is NOT what a scalar multiplier gives
    net.multiplier = 1.0
    net.flat_grads.zero_()
    with net:
        pred1 = model.engine.forward(packed, t, text, pooled, torch.ones(B, device=DEV), torch.zeros(Lt, 3, device=DEV),
                                     make_img_ids(hl, wl, DEV), save=False, t_div=1000.0)
    assert _rel(pred1, pred) > 1e-3


class Transformer2DModel(torch.nn.Module):  # kohya_lora.py:750: the Linear / 1x1-conv targets inside a UNet
    def __init__(self, ch, ctx_dim):
        super().__init__()
        self.proj_in = torch.nn.Conv2d(ch, ch, 1)
        self.to_q = torch.nn.Linear(ch, ch, bias=False)
        self.to_k = torch.nn.Linear(ctx_dim, ch, bias=False)
        self.to_v = torch.nn.Linear(ctx_dim, ch, bias=False)
        self.to_out = torch.nn.Linear(ch, ch)
        self.proj_out = torch.nn.Conv2d(ch, ch, 1)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        h = self.proj_in(x).flatten(2).transpose(1, 2)  # [B, HW, C]
        q, k, v = self.to_q(h), self.to_k(ctx), self.to_v(ctx)
        a = torch.nn.functional.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1)).squeeze(1)
        h = h + self.to_out(a)
        return x + self.proj_out(h.transpose(1, 2).reshape(B, C, H, W))


class ResnetBlock2D(torch.nn.Module):  # kohya_lora.py:751: the 3x3-conv targets
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1 = torch.nn.GroupNorm(8, cin)
        self.conv1 = torch.nn.Conv2d(cin, cout, 3, padding=1)
        self.conv_shortcut = torch.nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        return self.conv_shortcut(x) + self.conv1(torch.nn.functional.silu(self.norm1(x)))


def _toy_unet(ch=64, ctx_dim=96):
    cls = type("UNet2DConditionModel", (torch.nn.Module,), {})

    def init(self):
        torch.nn.Module.__init__(self)
        self.conv_in = torch.nn.Conv2d(4, ch, 3, padding=1)
        self.res = ResnetBlock2D(ch, ch)
        self.attn = Transformer2DModel(ch, ctx_dim)
        self.down = ResnetBlock2D(ch, ch)
        self.conv_out = torch.nn.Conv2d(ch, 4, 3, padding=1)

    def fwd(self, x, ctx):
        return self.conv_out(self.down(self.attn(self.res(self.conv_in(x)), ctx)))

    cls.__init__, cls.forward = init, fwd
    return cls()


@pytest.mark.parametrize("pred_type", ["epsilon", "v_prediction"])
def test_unet_style_eps_training_step_plumbing(pred_type):
    """BASELINE.json configs[0] / [1] PLUMBING (not a UNet engine): a UNet-shaped toy model whose Linear, 1x1-conv and 3x3-conv
    layers carry kohya-format adapters (lora_dim 4, conv_lora_dim 4, alpha 2 -> scale 0.5), trained for 3 steps with the
    eps / v path of the trainer: DDPM add_noise with integer timesteps (kernel), prediction through `LoRAModule.forward`
    (fused GEMM for every adapted layer; the frozen non-adapted layers are torch), the fused eps / v MSE with min-SNR-gamma
    weights (kernel), clip + AdamW (kernel) -- against the same three steps in eager PyTorch with the reference formulas."""
    import copy
    from ai_toolkit_b200 import LoRASpecialNetwork, calc_loss, ops
    from ai_toolkit_b200.optimizer import B200AdamW
    from ai_toolkit_b200.samplers import DDPMTable
    torch.manual_seed(7)
    unet = _toy_unet().to(DEV, torch.bfloat16).requires_grad_(False)
    ref_unet = copy.deepcopy(unet)
    net = LoRASpecialNetwork(None, unet, lora_dim=4, alpha=2, conv_lora_dim=4, conv_alpha=2, train_text_encoder=False,
                             target_lin_modules=["Transformer2DModel"], target_conv_modules=["ResnetBlock2D"])
    net.force_to(DEV, torch.float32)
    net._update_torch_multiplier()
    net.apply_to(None, unet, False, True)
    names = [l.lora_name for l in net.unet_loras]
    assert "lora_unet_attn_proj_in" in names and "lora_unet_attn_to_q" in names and "lora_unet_res_conv1" in names
    assert all(abs(l.scale - 0.5) < 1e-9 for l in net.unet_loras)
    with torch.no_grad():
        for l in net.unet_loras:
            l.lora_up.weight.normal_(0, 0.05)
    net.mark_params_changed()
    # eager twin: explicit A / B leaves applied with the reference formula through forward hooks on the twin's layers
    twins = []
    for l in net.unet_loras:
        path = l.lora_name[len("lora_unet_"):]
        mod = ref_unet
        for part in (path.split("_", 1) if not path.startswith("attn_") else ["attn", path[5:]]):
            mod = getattr(mod, part)
        A = l.lora_down.weight.detach().clone().requires_grad_(True)
        Bw = l.lora_up.weight.detach().clone().requires_grad_(True)
        twins.append((mod, A, Bw, l))

        def hook(m, inp, out, A=A, Bw=Bw, l=l):
            x = inp[0].float()
            if isinstance(m, torch.nn.Conv2d):
                lx = F.conv2d(F.conv2d(x, A, None, m.stride, m.padding), Bw)
            else:
                lx = F.linear(F.linear(x, A), Bw)
            return out + (lx * l.scale).to(out.dtype)

        mod.register_forward_hook(hook)
    rparams = [p for _, A, Bw, _ in twins for p in (A, Bw)]
    ropt = torch.optim.AdamW(rparams, lr=1e-3, eps=1e-6, weight_decay=1e-2)
    opt = B200AdamW(net, lr=1e-3, eps=1e-6, weight_decay=1e-2, max_grad_norm=1.0)
    tab = DDPMTable(prediction_type=pred_type, device=DEV)
    g = torch.Generator().manual_seed(3)
    B = 2
    lat = (torch.randn(B, 4, 32, 32, generator=g) * 0.18215 * 5).bfloat16().to(DEV)
    ctx = torch.randn(B, 77, 96, generator=g).bfloat16().to(DEV)
    mine, ref = [], []
    for it in range(3):
        noise = torch.randn(B, 4, 32, 32, generator=g).bfloat16().to(DEV)
        t = tab.sample_timesteps(B, generator=g).to(DEV)
        v = calc_loss.loss_vectors(t, is_flow_matching=False, prediction_type=pred_type, ddpm_table=tab, min_snr_gamma=5.0, device=DEV)
        # --- B200 path
        opt.zero_grad()
        noisy = ops.ddpm_add_noise(lat, noise, t, tab.device_table)
        with net:
            pred = unet(noisy, ctx)
            tot, _, dpred = ops.train_loss(pred.contiguous(), lat, noise, pack=False, **v)
            pred.backward(dpred)
        opt.step()
        mine.append(tot.item())
        # --- eager reference
        ropt.zero_grad(set_to_none=True)
        rn = tab.add_noise(lat, noise, t)
        assert torch.equal(rn, noisy)
        rp = ref_unet(rn, ctx)
        target = noise if pred_type == "epsilon" else tab.get_velocity(lat, noise, t)
        loss = F.mse_loss(rp.float(), target.float(), reduction="none").mean([1, 2, 3])
        loss = (loss * tab.snr_weights(t, 5.0).to(DEV)).mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(rparams, 1.0)
        ropt.step()
        ref.append(loss.item())
    assert max(abs(a - b) / abs(b) for a, b in zip(mine, ref)) < 5e-3, (mine, ref)
    off = 0
    for _, A, Bw, l in twins:
        for mine_p, ref_p in ((l.lora_down.weight, A), (l.lora_up.weight, Bw)):
            n = ref_p.numel()
            upd_m = net.flat_params[off:off + n] - 0  # noqa: F841
            off += n
            assert _rel(mine_p, ref_p) < 2e-2, l.lora_name
    sd = net.get_state_dict(dtype=torch.float16)
    assert "lora_unet_res_conv1.lora_down.weight" in sd and sd["lora_unet_res_conv1.lora_down.weight"].shape == (4, 64, 3, 3)
    assert "lora_unet_attn_to_q.alpha" in sd
