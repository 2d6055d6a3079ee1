"""tcgen05 flash attention (b200_attn_fwd / b200_attn_bwd) vs an fp32 PyTorch reference of the same op.
Tolerances: outputs are bf16 (one rounding, ulp 2^-8) of fp32-accumulated products of bf16 P: 1e-2 relative
Frobenius error on O and on dQ/dK/dV (the eager bf16 SDPA reference sits at the same distance from fp32)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _ref(Q, K, V, dO):
    q, k, v = (t.float().requires_grad_(True) for t in (Q, K, V))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(Q.shape[-1])
    lse = torch.logsumexp(s, -1)
    o = torch.softmax(s, -1) @ v
    o.backward(dO.float())
    return o.detach(), lse.detach(), q.grad, k.grad, v.grad


@pytest.mark.parametrize("B,H,L,split,scale_q", [(1, 2, 88, 24, 1.0), (2, 3, 320, 64, 1.0), (1, 2, 128, 0, 3.0),
                                                  (1, 1, 1000, 0, 6.0), (1, 4, 4608, 512, 1.0)])
def test_attention_fwd_bwd(B, H, L, split, scale_q):
    from ai_toolkit_b200 import attention
    torch.manual_seed(L)
    Q = (torch.randn(B, H, L, 128, device=DEV) * scale_q).bfloat16()
    K = torch.randn(B, H, L, 128, device=DEV).bfloat16()
    V = torch.randn(B, H, L, 128, device=DEV).bfloat16()
    D = H * 128
    pad = 64  # outputs live inside wider buffers (the single-stream concat buffer)
    o0 = torch.full((B * split, D), float("nan"), device=DEV, dtype=torch.bfloat16) if split else None
    o1 = torch.full((B * (L - split), D + pad), float("nan"), device=DEV, dtype=torch.bfloat16)
    lse = attention.fwd(Q, K, V, o0, o1[:, :D], split)
    torch.cuda.synchronize()
    dO = torch.randn(B, L, D, device=DEV).bfloat16()
    o_ref, lse_ref, dq_ref, dk_ref, dv_ref = _ref(Q, K, V, dO.view(B, L, H, 128).transpose(1, 2))
    o_tok = o_ref.transpose(1, 2).reshape(B, L, D)
    got = torch.cat(([o0.view(B, split, D)] if split else []) + [o1[:, :D].reshape(B, L - split, D)], 1)
    assert not torch.isnan(got.float()).any()
    assert torch.isnan(o1[:, D:].float()).all()  # nothing written outside the head columns
    assert _rel(got, o_tok) < 1e-2
    assert (lse - lse_ref).abs().max().item() < 2e-2
    do0 = dO[:, :split].reshape(B * split, D).contiguous() if split else None
    do1 = dO[:, split:].reshape(B * (L - split), D).contiguous()
    dQ, dK, dV = attention.bwd(Q, K, V, o0, o1[:, :D], do0, do1, lse, split)
    torch.cuda.synchronize()
    for name, g, r in (("dQ", dQ, dq_ref), ("dK", dK, dk_ref), ("dV", dV, dv_ref)):
        assert not torch.isnan(g.float()).any(), name
        assert _rel(g, r) < 1.5e-2, (name, _rel(g, r))


def test_attention_flux_shape_timing():
    """FLUX.1-dev joint attention (24 heads, L = 4608): report achieved TFLOP/s (informational)."""
    from ai_toolkit_b200 import attention
    B, H, L, split = 1, 24, 4608, 512
    Q, K, V = (torch.randn(B, H, L, 128, device=DEV).bfloat16() for _ in range(3))
    D = H * 128
    o0 = torch.empty(B * split, D, device=DEV, dtype=torch.bfloat16)
    o1 = torch.empty(B * (L - split), D, device=DEV, dtype=torch.bfloat16)
    do0, do1 = torch.randn_like(o0), torch.randn_like(o1)
    lse = attention.fwd(Q, K, V, o0, o1, split)
    attention.bwd(Q, K, V, o0, o1, do0, do1, lse, split)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(5):
        lse = attention.fwd(Q, K, V, o0, o1, split)
    ev[1].record()
    for _ in range(5):
        attention.bwd(Q, K, V, o0, o1, do0, do1, lse, split)
    ev[2].record()
    torch.cuda.synchronize()
    f = 4.0 * B * H * L * L * 128
    tf, tb = ev[0].elapsed_time(ev[1]) / 5, ev[1].elapsed_time(ev[2]) / 5
    print(f"attention fwd {tf*1e3:.0f} us = {f/tf/1e9:.0f} TFLOP/s; bwd {tb*1e3:.0f} us = {2.5*f/tb/1e9:.0f} TFLOP/s (algorithmic)")
    ref = torch.nn.functional.scaled_dot_product_attention(Q, K, V).transpose(1, 2).reshape(B, L, D)
    got = torch.cat([o0.view(B, split, D), o1.view(B, L - split, D)], 1)
    assert _rel(got, ref) < 1e-2


def test_attention_fwd_variant2_two_streams():
    """The alternative forward kernel (two independent online-softmax streams, B200_ATTN_FWD=2) stays correct; it is read
    once per process, so run it in a child process."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, math, torch; sys.path.insert(0, '.');\n"
        "from ai_toolkit_b200 import attention\n"
        "B,H,L,split=1,3,1000,0\n"
        "Q,K,V=(torch.randn(B,H,L,128,device='cuda').bfloat16() for _ in range(3))\n"
        "o1=torch.empty(B*L,H*128,device='cuda',dtype=torch.bfloat16)\n"
        "lse=attention.fwd(Q,K,V,None,o1,0)\n"
        "ref=torch.nn.functional.scaled_dot_product_attention(Q.float(),K.float(),V.float()).transpose(1,2).reshape(B*L,H*128)\n"
        "err=((o1.float()-ref).norm()/ref.norm()).item(); print('ERR',err); assert err<1e-2\n")
    env = dict(os.environ, B200_ATTN_FWD="2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr


@pytest.mark.parametrize("B,H,Lq,Lk", [(1, 2, 300, 77), (2, 3, 1000, 512), (1, 2, 64, 200), (1, 12, 3328, 512)])
def test_cross_attention_fwd_bwd(B, H, Lq, Lk):
    """Lq != Lk (b200_attn_fwd_x / b200_attn_bwd_x): Wan2.1's text cross-attention (toolkit/models/wan21/wan_attn.py:70-76),
    ragged on both sides, against the fp32 reference of the same op."""
    from ai_toolkit_b200 import attention
    torch.manual_seed(Lq + Lk)
    Q = torch.randn(B, H, Lq, 128, device=DEV).bfloat16()
    K = torch.randn(B, H, Lk, 128, device=DEV).bfloat16()
    V = torch.randn(B, H, Lk, 128, device=DEV).bfloat16()
    D = H * 128
    o1 = torch.full((B * Lq, D), float("nan"), device=DEV, dtype=torch.bfloat16)
    lse = attention.fwd(Q, K, V, None, o1, 0)
    dO = torch.randn(B, Lq, D, device=DEV).bfloat16()
    o_ref, lse_ref, dq_ref, dk_ref, dv_ref = _ref(Q, K, V, dO.view(B, Lq, H, 128).transpose(1, 2))
    assert not torch.isnan(o1.float()).any()
    assert _rel(o1.view(B, Lq, D), o_ref.transpose(1, 2).reshape(B, Lq, D)) < 1e-2
    assert (lse - lse_ref).abs().max().item() < 2e-2
    dQ, dK, dV = attention.bwd(Q, K, V, None, o1, None, dO.reshape(B * Lq, D).contiguous(), lse, 0)
    torch.cuda.synchronize()
    assert dQ.shape == Q.shape and dK.shape == K.shape and dV.shape == V.shape
    for name, g, r in (("dQ", dQ, dq_ref), ("dK", dK, dk_ref), ("dV", dV, dv_ref)):
        assert not torch.isnan(g.float()).any(), name
        assert _rel(g, r) < 1.5e-2, (name, _rel(g, r))


@pytest.mark.parametrize("L,Lk", [(640, 640), (300, 900)])
def test_attention_scores_growing_along_the_sequence(L, Lk):
    """The default forward keeps an OPTIMISTIC running maximum (the exact row maximum is only computed on the first block or
    when a row's partial sum exceeds 1e9).  Scores that grow by hundreds of log2 units from one KV block to the next force that
    fallback on every block, and rows whose large scores come FIRST exercise the opposite case (tiny later terms)."""
    from ai_toolkit_b200 import attention
    torch.manual_seed(5)
    B, H = 1, 2
    Q = (torch.randn(B, H, L, 128, device=DEV) * 3.0).bfloat16()
    K = torch.randn(B, H, Lk, 128, device=DEV)
    ramp = torch.linspace(0.2, 12.0, Lk, device=DEV).view(1, 1, Lk, 1)
    K[:, 0] = K[:, 0] * ramp[:, 0]            # head 0: scores grow along the keys
    K[:, 1] = K[:, 1] * ramp.flip(2)[:, 0]    # head 1: the largest scores come first
    K = K.bfloat16()
    V = torch.randn(B, H, Lk, 128, device=DEV).bfloat16()
    o1 = torch.full((B * L, H * 128), float("nan"), device=DEV, dtype=torch.bfloat16)
    lse = attention.fwd(Q, K, V, None, o1, 0)
    q, k, v = Q.float(), K.float(), V.float()
    s = (q @ k.transpose(-1, -2)) / math.sqrt(128)
    assert float(s.max()) > 100  # natural-log units: far beyond what exp2 could hold without the running maximum
    o_ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, L, H * 128)
    assert torch.isfinite(o1.float()).all() and torch.isfinite(lse).all()
    assert _rel(o1.view(B, L, -1), o_ref) < 1e-2
    assert (lse - torch.logsumexp(s, -1)).abs().max().item() < 5e-2


@pytest.mark.parametrize("B,H,Lq,Lk,d", [(2, 5, 1024, 1024, 64), (1, 10, 4096, 4096, 64), (2, 3, 600, 77, 64), (1, 2, 300, 300, 40)])
def test_attention_zero_padded_half_heads_skip_the_padding(B, H, Lq, Lk, d):
    """head_live = 64 (`b200_attn_fwd_xd / _bwd_xd`): heads of <= 64 channels zero-padded to 128 (SDXL 64, SD1.5 40).  The kernels
    skip the zero half of every contraction / output; the kept half is the same sequence of MMAs, so the results must equal the
    head_live = 128 run on the same padded tensors BIT FOR BIT, the padded output columns are exact zeros over NaN-filled buffers,
    and both match the fp32 reference."""
    from ai_toolkit_b200 import attention
    torch.manual_seed(Lq + Lk + d)
    def pad(x):
        out = torch.zeros(*x.shape[:-1], 128, device=DEV, dtype=torch.bfloat16)
        out[..., :d] = x
        return out
    Q = pad(torch.randn(B, H, Lq, d, device=DEV).bfloat16())
    K = pad(torch.randn(B, H, Lk, d, device=DEV).bfloat16())
    V = pad(torch.randn(B, H, Lk, d, device=DEV).bfloat16())
    D = H * 128
    scale = d ** -0.5
    dO = pad(torch.randn(B, Lq, H, d, device=DEV).bfloat16()).reshape(B * Lq, D)
    res = {}
    for live in (128, 64):
        o1 = torch.full((B * Lq, D), float("nan"), device=DEV, dtype=torch.bfloat16)
        lse = attention.fwd(Q, K, V, None, o1, 0, scale=scale, head_live=live)
        dQ, dK, dV = attention.bwd(Q, K, V, None, o1, None, dO, lse, 0, scale=scale, head_live=live)
        torch.cuda.synchronize()
        res[live] = (o1, lse, dQ, dK, dV)
    o64 = res[64][0].view(B * Lq, H, 128)
    assert bool((o64[..., 64:] == 0).all()) and torch.isfinite(o64.float()).all()
    assert torch.equal(res[64][0], res[128][0]) and torch.equal(res[64][1], res[128][1])
    for i, name in ((2, "dQ"), (3, "dK"), (4, "dV")):
        g64, g128 = res[64][i], res[128][i]
        assert bool((g64[..., 64:] == 0).all()), name
        assert torch.equal(g64[..., :64], g128[..., :64]), name
    # against the fp32 reference of the un-padded op
    q, k, v = (t[..., :d].float() for t in (Q, K, V))
    q.requires_grad_(True); k.requires_grad_(True); v.requires_grad_(True)
    o_ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, scale=scale)
    o_ref.backward(dO.view(B, Lq, H, 128)[..., :d].transpose(1, 2).float())
    assert _rel(o64[..., :d].reshape(B, Lq, H, d).transpose(1, 2), o_ref) < 1e-2
    for g, r, name in ((res[64][2], q.grad, "dQ"), (res[64][3], k.grad, "dK"), (res[64][4], v.grad, "dV")):
        assert _rel(g[..., :d], r) < 1.5e-2, name


@pytest.mark.parametrize("B,H,Lq,Lk,d", [(2, 8, 256, 256, 160), (1, 8, 64, 64, 160), (2, 2, 1000, 77, 160), (1, 3, 37, 300, 192),
                                          (1, 2, 100, 100, 256)])
def test_small_attention_head_dims_above_128(B, H, Lq, Lk, d):
    """`b200_attn_small_fwd / _bwd` (CUDA-core kernel for SD1.5's 160-wide heads): token-major strided views in, vs the fp32
    reference of softmax(q k^T / sqrt(d)) v and its gradients; ragged tiles on both sides; cross attention."""
    from ai_toolkit_b200 import attention
    torch.manual_seed(Lq + Lk + d)
    inner = H * d
    qkv = torch.randn(B * Lq, 3 * inner + 8, device=DEV).bfloat16()      # q is a column slice of a wider buffer
    kv = torch.randn(B * Lk, 2 * inner, device=DEV).bfloat16()
    q, k, v = qkv[:, 8:8 + inner], kv[:, :inner], kv[:, inner:]
    o, lse = attention.small_fwd(q, k, v, B, H, Lq, Lk, d)
    dO = torch.randn(B * Lq, inner, device=DEV).bfloat16()
    dq = torch.full((B * Lq, inner), float("nan"), device=DEV, dtype=torch.bfloat16)
    dkv = torch.full((B * Lk, 2 * inner), float("nan"), device=DEV, dtype=torch.bfloat16)
    attention.small_bwd(q, k, v, o, dO, lse, dq, dkv[:, :inner], dkv[:, inner:], B, H, Lq, Lk, d)
    torch.cuda.synchronize()

    def heads(t, L):
        return t.float().reshape(B, L, H, d).transpose(1, 2).detach().clone().requires_grad_(True)
    qr, kr, vr = heads(q, Lq), heads(k, Lk), heads(v, Lk)
    s = (qr @ kr.transpose(-1, -2)) * d ** -0.5
    o_ref = torch.softmax(s, -1) @ vr
    o_ref.backward(heads(dO, Lq).detach())
    back = lambda t, L: t.transpose(1, 2).reshape(B * L, inner)  # noqa: E731
    assert _rel(o, back(o_ref, Lq)) < 6e-3
    assert (lse - torch.logsumexp(s, -1)).abs().max().item() < 2e-3
    assert torch.isfinite(dq.float()).all() and torch.isfinite(dkv.float()).all()
    assert _rel(dq, back(qr.grad, Lq)) < 1e-2 and _rel(dkv[:, :inner], back(kr.grad, Lk)) < 1e-2
    assert _rel(dkv[:, inner:], back(vr.grad, Lk)) < 1e-2
    with pytest.raises(Exception):
        attention.small_fwd(q[:, :H * 128], k[:, :H * 128], v[:, :H * 128], B, H, Lq, Lk, 128)
