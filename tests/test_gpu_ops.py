"""Row / loss / optimizer kernels vs plain PyTorch fp32 references of the same op (through the C ABI)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("M,D,rps", [(64, 256, 32), (1000, 3072, 500), (4608, 3072, 4608)])
def test_ln_modulate_fwd_bwd(M, D, rps):
    from ai_toolkit_b200 import ops
    torch.manual_seed(0)
    S = (M + rps - 1) // rps
    x = torch.randn(M, D, device=DEV).bfloat16()
    mod = (torch.randn(S, 6 * D, device=DEV) * 0.3).bfloat16()
    shift, scale = mod[:, :D], mod[:, D:2 * D]
    out, mean, rstd = ops.ln_modulate_fwd(x, shift, scale, rps)
    # eager bf16 model arithmetic
    idx = torch.arange(M, device=DEV) // rps
    ref = F.layer_norm(x, (D,), eps=1e-6) * (1 + scale[idx]) + shift[idx]
    assert _rel(out, ref) < 4e-3
    # backward vs fp32 autograd
    dy = torch.randn(M, D, device=DEV).bfloat16()
    dres = torch.randn(M, D, device=DEV).bfloat16()
    xf = x.float().requires_grad_(True)
    scf = scale.float().requires_grad_(True)
    shf = shift.float().requires_grad_(True)
    y = F.layer_norm(xf, (D,), eps=1e-6) * (1 + scf[idx]) + shf[idx]
    y.backward(dy.float())
    dx = ops.ln_modulate_bwd(dy, x, mean, rstd, scale, rps, dres=dres)
    assert _rel(dx, xf.grad + dres.float()) < 6e-3
    dmod = torch.zeros(S, 6 * D, device=DEV)
    ops.col_reduce(dy, rps, b=x, mean=mean, rstd=rstd, sum_a=dmod[:, :D], sum_ab=dmod[:, D:2 * D])
    assert _rel(dmod[:, :D], shf.grad) < 2e-3
    assert _rel(dmod[:, D:2 * D], scf.grad) < 6e-3


def test_col_reduce_gate():
    from ai_toolkit_b200 import ops
    M, D, rps = 1024, 3072, 512
    dh = torch.randn(M, D, device=DEV).bfloat16()
    y = torch.randn(M, D, device=DEV).bfloat16()
    gate = torch.randn(2, 3 * D, device=DEV).bfloat16()[:, D:2 * D]
    dgate = torch.zeros(2, D, device=DEV)
    dy = torch.empty_like(dh)
    ops.col_reduce(dh, rps, b=y, g=gate, mul_out=dy, sum_ab=dgate)
    idx = torch.arange(M, device=DEV) // rps
    assert _rel(dy, dh.float() * gate.float()[idx]) < 4e-3
    want = torch.stack([(dh.float() * y.float())[:512].sum(0), (dh.float() * y.float())[512:].sum(0)])
    assert _rel(dgate, want) < 1e-4


@pytest.mark.parametrize("M,rps", [(996, 498), (1030, 515), (77, 77)])
def test_col_reduce_ragged_rows(M, rps):
    """The 16-byte kernel loads 4 rows at a time with a clamped tail: sample boundaries / tails that are not multiples of 4."""
    from ai_toolkit_b200 import ops
    torch.manual_seed(1)
    D = 3072
    S = (M + rps - 1) // rps
    a = torch.randn(M, D, device=DEV).bfloat16()
    b = torch.randn(M, D, device=DEV).bfloat16()
    mean = torch.randn(M, device=DEV) * 0.1
    rstd = torch.rand(M, device=DEV) + 0.5
    idx = torch.arange(M, device=DEV) // rps
    sums = torch.zeros(S, 2 * D, device=DEV)
    ops.col_reduce(a, rps, b=b, mean=mean, rstd=rstd, sum_a=sums[:, :D], sum_ab=sums[:, D:])
    xh = (b.float() - mean[:, None]) * rstd[:, None]
    want_a = torch.zeros(S, D, device=DEV).index_add_(0, idx, a.float())
    want_ab = torch.zeros(S, D, device=DEV).index_add_(0, idx, a.float() * xh)
    assert _rel(sums[:, :D], want_a) < 1e-5 and _rel(sums[:, D:], want_ab) < 1e-5
    gate = torch.randn(S, D, device=DEV).bfloat16()
    out = torch.full((M + 1, D), 7.0, device=DEV).bfloat16()
    dg = torch.zeros(S, D, device=DEV)
    ops.col_reduce(a, rps, b=b, g=gate, mul_out=out[:M], sum_ab=dg)
    assert torch.equal(out[:M], (a.float() * gate.float()[idx]).bfloat16()) and bool((out[M] == 7.0).all())
    assert _rel(dg, torch.zeros(S, D, device=DEV).index_add_(0, idx, a.float() * b.float())) < 1e-5


def _rope_ref(x, cos, sin):
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(-2)
    return x.float() * cos + rot.float() * sin


def test_qk_norm_rope_fwd_bwd():
    from ai_toolkit_b200 import ops
    torch.manual_seed(0)
    B, Lseg, off, Ltot, H = 2, 96, 32, 128, 3
    D = H * 128
    qkv = torch.randn(B * Lseg, 3 * D, device=DEV).bfloat16()
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    wq = (1 + 0.1 * torch.randn(128, device=DEV)).bfloat16()
    wk = (1 + 0.1 * torch.randn(128, device=DEV)).bfloat16()
    ang = torch.rand(Ltot, 64, device=DEV, dtype=torch.float64) * 6.28
    cos = ang.cos().repeat_interleave(2, 1).float().contiguous()
    sin = ang.sin().repeat_interleave(2, 1).float().contiguous()
    Q = torch.zeros(B, H, Ltot, 128, device=DEV, dtype=torch.bfloat16)
    K, V = torch.zeros_like(Q), torch.zeros_like(Q)
    ops.qk_norm_rope_fwd(q, k, v, wq, wk, cos, sin, Q, K, V, B, Lseg, off)

    def ref(xin, w):
        x = xin.reshape(B, Lseg, H, 128).transpose(1, 2)  # [B,H,L,128]
        var = x.float().pow(2).mean(-1, keepdim=True)
        xn = (x * torch.rsqrt(var + 1e-6)).to(torch.bfloat16) * w
        return _rope_ref(xn, cos[off:off + Lseg], sin[off:off + Lseg])

    assert _rel(Q[:, :, off:off + Lseg], ref(q, wq)) < 4e-3
    assert _rel(K[:, :, off:off + Lseg], ref(k, wk)) < 4e-3
    assert torch.equal(V[:, :, off:off + Lseg], v.reshape(B, Lseg, H, 128).transpose(1, 2))
    assert Q[:, :, :off].abs().sum() == 0
    # backward vs fp32 autograd of the smooth function
    dQ = torch.randn_like(Q)
    dK = torch.randn_like(Q)
    dV = torch.randn_like(Q)
    dqkv = torch.zeros_like(qkv)
    ops.qk_norm_rope_bwd(dQ, dK, dV, q, k, wq, wk, cos, sin, dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], B, Lseg, off)
    for xin, w, dout, got in ((q, wq, dQ, dqkv[:, :D]), (k, wk, dK, dqkv[:, D:2 * D])):
        xf = xin.float().requires_grad_(True)
        x = xf.reshape(B, Lseg, H, 128).transpose(1, 2)
        xn = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + 1e-6) * w.float()
        y = _rope_ref(xn, cos[off:off + Lseg], sin[off:off + Lseg])
        y.backward(dout[:, :, off:off + Lseg].float())
        assert _rel(got, xf.grad) < 6e-3
    assert torch.equal(dqkv[:, 2 * D:].reshape(B, Lseg, H, 128).transpose(1, 2), dV[:, :, off:off + Lseg])


def test_silu_timestep_add():
    from ai_toolkit_b200 import ops
    x = torch.randn(3, 3072, device=DEV).bfloat16()
    assert _rel(ops.silu(x), F.silu(x)) < 3e-3
    t01 = torch.tensor([0.5005, 0.001, 0.999], device=DEV)
    got = ops.timestep_embed(t01)
    t = (t01.bfloat16() * 1000).float()
    half = 128
    f = torch.exp(-math.log(10000.0) * torch.arange(half, device=DEV, dtype=torch.float32) / half)
    ref = torch.cat([torch.cos(t[:, None] * f), torch.sin(t[:, None] * f)], -1).bfloat16()
    assert (got.float() - ref.float()).abs().max() < 2e-2
    a, b, c = (torch.randn(1000, device=DEV).bfloat16() for _ in range(3))
    assert torch.equal(ops.add_bf16(a, b, c), (a + b) + c)


@pytest.mark.parametrize("Bm,N,K,r", [(1, 18432, 3072, 16), (2, 9216, 3072, 4), (3, 512, 256, 0), (4, 1000, 264, 64)])
def test_lora_gemv(Bm, N, K, r):
    from ai_toolkit_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(Bm, K, device=DEV).bfloat16()
    W = (torch.randn(N, K, device=DEV) * 0.02).bfloat16()
    bias = (torch.randn(N, device=DEV) * 0.02).bfloat16()
    A = torch.randn(r, K, device=DEV) * 0.05 if r else None
    Bw = torch.randn(N, r, device=DEV) * 0.05 if r else None
    c = 0.7
    y, z = ops.lora_gemv_fwd(x, W, bias, A, Bw, c)
    base = (x.float() @ W.float().t() + bias.float()).bfloat16()
    if r:
        lora = (c * (x.float() @ A.t()) @ Bw.t()).bfloat16()
        ref = base + lora
        assert _rel(z, c * x.float() @ A.t()) < 1e-5
    else:
        ref = base
    assert _rel(y, ref) < 4e-3
    if r:
        dy = torch.randn(Bm, N, device=DEV)
        dA0, dB0 = torch.randn_like(A), torch.randn_like(Bw)
        dA, dBw = dA0.clone(), dB0.clone()
        ops.lora_gemv_bwd(dy, x, z, A, Bw, c, dA, dBw)
        t = c * dy @ Bw
        assert _rel(dA - dA0, t.t() @ x.float()) < 1e-4
        assert _rel(dBw - dB0, dy.t() @ (c * x.float() @ A.t())) < 1e-4


@pytest.mark.parametrize("B,C,H,W", [(1, 16, 128, 128), (3, 16, 32, 48)])
def test_flow_noise_and_loss(B, C, H, W):
    from ai_toolkit_b200 import ops
    torch.manual_seed(0)
    x0 = torch.randn(B, C, H, W, device=DEV).bfloat16()
    noise = torch.randn(B, C, H, W, device=DEV).bfloat16()
    t = torch.tensor([250.0, 999.0, 1.0][:B], device=DEV)

    def pack(z):
        return z.view(B, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // 2) * (W // 2), C * 4)

    noisy = ops.flow_add_noise(x0, noise, t, pack=True)
    t01 = (t / 1000).view(B, 1, 1, 1)
    ref = ((1 - t01) * x0 + t01 * noise).bfloat16()
    assert _rel(noisy, pack(ref)) < 2e-3  # fp32 fma vs mul+add may differ in the last bf16 bit
    pred = torch.randn(B, (H // 2) * (W // 2), C * 4, device=DEV).bfloat16()
    tot, per, dpred = ops.flow_loss(pred, x0, noise, pack=True)
    target = pack((noise - x0))
    pf = pred.float().requires_grad_(True)
    loss_ps = F.mse_loss(pf, target.float(), reduction="none").mean([1, 2])
    loss = loss_ps.mean()
    loss.backward()
    assert abs(tot.item() - loss.item()) / loss.item() < 1e-5
    assert _rel(per, loss_ps) < 1e-5
    assert _rel(dpred, pf.grad) < 4e-3


def test_clip_adamw_matches_torch():
    from ai_toolkit_b200 import ops
    torch.manual_seed(0)
    n = 100003
    p0 = torch.randn(n, device=DEV) * 0.1
    p = p0.clone()
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-3, eps=1e-6, weight_decay=1e-2)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    ema, ema_ref = p0.clone(), p0.clone()
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-6, 1e-2, 1.0, 0.99, 1.0], device=DEV)
    state = torch.zeros(8, device=DEV, dtype=torch.int64)
    sumsq = torch.zeros(1, device=DEV, dtype=torch.float64)
    norm = torch.zeros(1, device=DEV)
    for step in range(1, 6):
        g = torch.randn(n, device=DEV) * (0.02 if step % 2 else 0.001)
        ref_p.grad = g.clone()
        tn = torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        d = 0.99  # constant: the trainer's EMA has no warm-up (use_num_updates=False, BaseSDTrainProcess.py:798-803)
        ema_ref.sub_((1 - d) * (ema_ref - ref_p.data))
        ops.grad_sumsq(g, sumsq)
        ops.clip_adamw(p, g, m, v, sumsq, hyper, state, ema=ema, norm_out=norm)
        assert abs(norm.item() - tn.item()) / tn.item() < 1e-5
        assert _rel(p, ref_p.data) < 1e-6
        assert _rel(ema, ema_ref) < 1e-6
    assert state[0].item() == 5


def test_repack():
    import ctypes
    from ai_toolkit_b200 import cabi, ops
    flat = torch.randn(16 * 328 + 200 * 16, device=DEV)
    pack = torch.zeros(64 * 328 + 200 * 64, device=DEV, dtype=torch.bfloat16)
    tab = (cabi.RepackEntry * 2)()
    tab[0] = cabi.RepackEntry(0, 0, 16, 328, 328, 0)
    tab[1] = cabi.RepackEntry(16 * 328, 64 * 328, 200, 16, 64, 0)
    tab_dev = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8).to(DEV)
    ops.repack_lora(flat, pack, tab_dev, 2)
    A = pack[:64 * 328].view(64, 328)
    Bp = pack[64 * 328:].view(200, 64)
    assert torch.equal(A[:16], flat[:16 * 328].view(16, 328).bfloat16()) and A[16:].abs().sum() == 0
    assert torch.equal(Bp[:, :16], flat[16 * 328:].view(200, 16).bfloat16()) and Bp[:, 16:].abs().sum() == 0
