"""Tensor-level wrappers over the C ABI (``include/b200_lora.h``).

PyTorch tensors are used as device-memory handles only; every function here launches hand-written
sm_100a kernels through ``ai_toolkit_b200.cabi`` on torch's current stream and raises ``B200Error``
when the library or the device is missing.  There is no eager / CPU fallback.
"""
from __future__ import annotations

from ctypes import c_void_p

import torch

from . import cabi
from .cabi import ACT_GELU_TANH, ACT_NONE, gemm_bf16  # noqa: F401

RANK_PAD = 64  # rank-side operands are padded to one 64-wide k-block


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


def _ld(t):
    return 0 if t is None else int(t.stride(-2))


def _dev(t):
    return t.device.index


def ln_modulate_fwd(x, shift, scale, rows_per_sample, out=None, save_stats=True, eps=1e-6):
    """x [M, D] bf16; shift/scale [S, D] bf16 views (row stride = stride(0)) or None."""
    M, D = x.shape
    out = torch.empty((M, D), device=x.device, dtype=torch.bfloat16) if out is None else out
    mean = torch.empty(M, device=x.device, dtype=torch.float32) if save_stats else None
    rstd = torch.empty(M, device=x.device, dtype=torch.float32) if save_stats else None
    ldmod = int(scale.stride(0)) if scale is not None else (int(shift.stride(0)) if shift is not None else 8)
    cabi.call("b200_ln_modulate_fwd", _p(x), _ld(x), _p(shift), _p(scale), ldmod, int(rows_per_sample), _p(out), _ld(out),
              _p(mean), _p(rstd), M, D, float(eps), device=_dev(x))
    return out, mean, rstd


def ln_modulate_bwd(dy, x, mean, rstd, scale, rows_per_sample, dres=None, out=None):
    M, D = x.shape
    out = torch.empty((M, D), device=x.device, dtype=torch.bfloat16) if out is None else out
    ldmod = int(scale.stride(0)) if scale is not None else 8
    cabi.call("b200_ln_modulate_bwd", _p(dy), _ld(dy), _p(x), _ld(x), _p(mean), _p(rstd), _p(scale), ldmod,
              int(rows_per_sample), _p(dres), _ld(dres), _p(out), _ld(out), M, D, device=_dev(x))
    return out


def col_reduce(a, rows_per_sample, b=None, mean=None, rstd=None, g=None, mul_out=None, sum_a=None, sum_ab=None):
    """sum_a / sum_ab are fp32 views [S, D] (row stride = stride(0)) that are ACCUMULATED into."""
    M, D = a.shape
    ref = sum_a if sum_a is not None else sum_ab
    ldsum = int(ref.stride(0)) if ref is not None else D
    if sum_a is not None and sum_ab is not None:
        assert sum_a.stride(0) == sum_ab.stride(0)
    ldg = int(g.stride(0)) if g is not None else 0
    cabi.call("b200_col_reduce", _p(a), _ld(a), _p(b), _ld(b), _p(mean), _p(rstd), _p(g), ldg, _p(mul_out), _ld(mul_out),
              _p(sum_a), _p(sum_ab), ldsum, int(rows_per_sample), M, D, device=_dev(a))


def qk_norm_rope_fwd(q, k, v, wq, wk, cos, sin, Q, K, V, B, Lseg, seq_off, eps=1e-6):
    """q/k/v: [B*Lseg, H*128] bf16 views sharing one row stride; Q/K/V: [B, H, Ltot, 128] bf16."""
    H, Ltot = Q.shape[1], Q.shape[2]
    assert q.stride(0) == k.stride(0) == v.stride(0)
    cabi.call("b200_qk_norm_rope_fwd", _p(q), _p(k), _p(v), int(q.stride(0)), _p(wq), _p(wk), _p(cos), _p(sin), _p(Q),
              _p(K), _p(V), int(B), int(Lseg), int(seq_off), int(Ltot), int(H), 128, float(eps), device=_dev(q))


def qk_norm_rope_bwd(dQ, dK, dV, q, k, wq, wk, cos, sin, dq, dk, dv, B, Lseg, seq_off, eps=1e-6):
    H, Ltot = dQ.shape[1], dQ.shape[2]
    assert q.stride(0) == k.stride(0) and dq.stride(0) == dk.stride(0) == dv.stride(0)
    cabi.call("b200_qk_norm_rope_bwd", _p(dQ), _p(dK), _p(dV), _p(q), _p(k), int(q.stride(0)), _p(wq), _p(wk), _p(cos),
              _p(sin), _p(dq), _p(dk), _p(dv), int(dq.stride(0)), int(B), int(Lseg), int(seq_off), int(Ltot), int(H), 128,
              float(eps), device=_dev(q))


def rms_rope_fwd(x, weight, cos, sin, out, B, Lseg, seq_off=0, eps=1e-6, mode=1):
    """x [B*Lseg, >= H*128] bf16 view -> out [B, H, Ltot, 128] bf16 (RMSNorm across heads + optional RoPE, or mode 0: plain
    re-layout); returns rstd [B*Lseg] fp32 (None for mode 0)."""
    H, Ltot = out.shape[1], out.shape[2]
    rstd = torch.empty(B * Lseg, device=x.device, dtype=torch.float32) if mode == 1 else None
    cabi.call("b200_rms_rope_fwd", _p(x), int(x.stride(0)), _p(weight), _p(cos), _p(sin), _p(out), _p(rstd), int(B), int(Lseg),
              int(seq_off), int(Ltot), int(H), float(eps), int(mode), device=_dev(x))
    return rstd


def rms_rope_bwd(dY, x, weight, cos, sin, rstd, dx, B, Lseg, seq_off=0, mode=1):
    """dY [B, H, Ltot, 128] -> dx [B*Lseg, >= H*128] view (token-major)."""
    H, Ltot = dY.shape[1], dY.shape[2]
    cabi.call("b200_rms_rope_bwd", _p(dY), _p(x), 0 if x is None else int(x.stride(0)), _p(weight), _p(cos), _p(sin), _p(rstd),
              _p(dx), int(dx.stride(0)), int(B), int(Lseg), int(seq_off), int(Ltot), int(H), int(mode), device=_dev(dY))
    return dx


def ln_affine_fwd(x, weight, bias, eps=1e-5, out=None):
    """LayerNorm(x) * weight + bias over the last dim of x [M, D] bf16 (any D % 8 == 0) -> (y, mean, rstd)."""
    M, D = x.shape
    out = torch.empty((M, D), device=x.device, dtype=torch.bfloat16) if out is None else out
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    cabi.call("b200_ln_affine_fwd", _p(x), _ld(x), _p(weight), _p(bias), _p(out), _ld(out), _p(mean), _p(rstd), int(M), int(D),
              float(eps), device=_dev(x))
    return out, mean, rstd


def ln_affine_bwd(dy, x, mean, rstd, weight, dres=None, out=None):
    M, D = x.shape
    out = torch.empty((M, D), device=x.device, dtype=torch.bfloat16) if out is None else out
    cabi.call("b200_ln_affine_bwd", _p(dy), _ld(dy), _p(x), _ld(x), _p(mean), _p(rstd), _p(weight), _p(dres), _ld(dres), _p(out),
              _ld(out), int(M), int(D), device=_dev(x))
    return out


def groupnorm_fwd(x, weight, bias, groups=32, eps=1e-5, silu=False):
    """x [B, C, H, W] bf16 contiguous -> (y like x, mean [B*G], rstd [B*G])."""
    B, C = x.shape[0], x.shape[1]
    HW = x.numel() // (B * C)
    out = torch.empty_like(x)
    mean = torch.empty(B * groups, device=x.device, dtype=torch.float32)
    rstd = torch.empty(B * groups, device=x.device, dtype=torch.float32)
    cabi.call("b200_groupnorm_fwd", _p(x), _p(weight), _p(bias), _p(out), _p(mean), _p(rstd), int(B), int(C), int(HW),
              int(groups), float(eps), int(bool(silu)), device=_dev(x))
    return out, mean, rstd


def groupnorm_bwd(dy, x, weight, bias, mean, rstd, groups=32, silu=False):
    B, C = x.shape[0], x.shape[1]
    HW = x.numel() // (B * C)
    dx = torch.empty_like(x)
    cabi.call("b200_groupnorm_bwd", _p(dy), _p(x), _p(weight), _p(bias), _p(mean), _p(rstd), _p(dx), int(B), int(C), int(HW),
              int(groups), int(bool(silu)), device=_dev(x))
    return dx


def geglu_fwd(proj, out=None):
    """proj [M, 2F] = (hidden | gate) -> hidden * gelu(gate) [M, F]."""
    M, F2 = proj.shape
    F = F2 // 2
    out = torch.empty((M, F), device=proj.device, dtype=torch.bfloat16) if out is None else out
    cabi.call("b200_geglu_fwd", _p(proj), _ld(proj), _p(out), _ld(out), int(M), int(F), device=_dev(proj))
    return out


def geglu_bwd(dy, proj, out=None):
    M, F2 = proj.shape
    out = torch.empty_like(proj) if out is None else out
    cabi.call("b200_geglu_bwd", _p(dy), _ld(dy), _p(proj), _ld(proj), _p(out), _ld(out), int(M), int(F2 // 2), device=_dev(proj))
    return out


def heads_pad(x, out, B, L, head_dim):
    """x [B*L, >= H*head_dim] view -> out [B, H, L, 128] (zero-padded heads)."""
    cabi.call("b200_heads_pad", _p(x), _p(out), int(x.stride(0)), int(B), int(L), int(out.shape[1]), int(head_dim), 1,
              device=_dev(x))
    return out


def heads_pad_multi(pairs, B, head_dim, to_heads=True):
    """Up to three (token-major view [B*L_i, >= H*head_dim], head-major [B, H, L_i, 128]) pairs of ONE attention in one launch
    (q / k / v, or with to_heads=False the gradients dQ / dK / dV back into token-major views)."""
    assert 1 <= len(pairs) <= 3
    args = []
    H = None
    for tm, hm in pairs:
        H = int(hm.shape[1])
        src, dst = (tm, hm) if to_heads else (hm, tm)
        args += [_p(src), _p(dst), int(tm.stride(0)), int(hm.shape[2])]
    for _ in range(3 - len(pairs)):
        args += [None, None, 0, 0]
    cabi.call("b200_heads_pad3", *args, len(pairs), int(B), H, int(head_dim), int(bool(to_heads)), device=_dev(pairs[0][0]))


def heads_unpad(hm, out, B, L, head_dim):
    """head-major [B, H, L, 128] -> out [B*L, >= H*head_dim] view (first head_dim columns of every head)."""
    cabi.call("b200_heads_pad", _p(hm), _p(out), int(out.stride(0)), int(B), int(L), int(hm.shape[1]), int(head_dim), 0,
              device=_dev(hm))
    return out


def silu(x, out=None):
    out = torch.empty_like(x) if out is None else out
    cabi.call("b200_silu", _p(x), _p(out), x.numel(), device=_dev(x))
    return out


def timestep_embed(t, dim=256, div=1.0, mult=1000.0, max_period=10000.0):
    """t fp32 [B]; the embedded value is bf16(bf16(t / div) * mult)."""
    B = t.shape[0]
    out = torch.empty((B, dim), device=t.device, dtype=torch.bfloat16)
    cabi.call("b200_timestep_embed", _p(t), _p(out), B, dim, float(max_period), float(div), float(mult), device=_dev(t))
    return out


def add_bf16(a, b, c=None, out=None):
    out = torch.empty_like(a) if out is None else out
    cabi.call("b200_add_bf16", _p(a), _p(b), _p(c), _p(out), a.numel(), device=_dev(a))
    return out


def lora_gemv_fwd(x, W, bias, A=None, Bw=None, c=1.0, out=None, row_c=None):
    """x [Bm, K] bf16, W [N, K] bf16, A [r, K] fp32, Bw [N, r] fp32 -> (y [Bm, N] bf16, z [Bm, r] fp32 | None).
    row_c: optional fp32 [Bm] per-row coefficient (per-sample multipliers), multiplied with c."""
    Bm, K = x.shape
    N = W.shape[0]
    r = 0 if A is None else int(A.shape[0])
    y = torch.empty((Bm, N), device=x.device, dtype=torch.bfloat16) if out is None else out
    z = torch.empty((Bm, r), device=x.device, dtype=torch.float32) if r > 0 else None
    cabi.call("b200_lora_gemv_fwd_rows", _p(x), _ld(x), _p(W), _ld(W), _p(bias), _p(A), _p(Bw), r, float(c), _p(row_c), _p(y),
              _ld(y), _p(z), Bm, N, K, device=_dev(x))
    return y, z


def lora_gemv_bwd(dy, x, z, A, Bw, c, dA, dBw, t_ws=None, row_c=None):
    """dy fp32 [Bm, N] (row stride = stride(0)); accumulates into dA [r, K], dBw [N, r] (fp32)."""
    Bm, K = x.shape
    N, r = Bw.shape
    t_ws = torch.empty(Bm * r, device=x.device, dtype=torch.float32) if t_ws is None else t_ws
    cabi.call("b200_lora_gemv_bwd_rows", _p(dy), int(dy.stride(0)), _p(x), _ld(x), _p(z), _p(A), _p(Bw), r, float(c),
              _p(row_c), _p(dA), _p(dBw), _p(t_ws), Bm, N, K, device=_dev(x))


def flow_add_noise(latents, noise, t, pack=True, out=None):
    B, C, H, W = latents.shape
    if out is None:
        shape = (B, (H // 2) * (W // 2), C * 4) if pack else (B, C, H, W)
        out = torch.empty(shape, device=latents.device, dtype=torch.bfloat16)
    cabi.call("b200_flow_add_noise", _p(latents), _p(noise), _p(t), _p(out), B, C, H, W, int(pack), device=_dev(latents))
    return out


def flow_loss(pred, latents, noise, pack=True, gscale=1.0, dpred=None, want_grad=True, loss_ws=None):
    """-> (loss_total [1] fp32, loss_per_sample [B] fp32, dpred like pred)."""
    B, C, H, W = latents.shape
    if want_grad and dpred is None:
        dpred = torch.empty_like(pred)
    if loss_ws is None:
        loss_ws = torch.empty(B + 1, device=pred.device, dtype=torch.float32)
    per, tot = loss_ws[:B], loss_ws[B:B + 1]
    cabi.call("b200_flow_loss", _p(pred), _p(latents), _p(noise), _p(dpred if want_grad else None), _p(per), _p(tot), B, C,
              H, W, int(pack), float(gscale), device=_dev(pred))
    return tot, per, dpred


def train_loss(pred, latents, noise, *, target=None, coef_noise=None, coef_latent=None, sample_weight=None, mask=None,
               pack=False, gscale=1.0, dpred=None, want_grad=True, loss_ws=None, target_in_pred_layout=False):
    """The default 'mse' path of `SDTrainer.calculate_loss` (SDTrainer.py:522-1052) in one launch; see include/b200_lora.h.
    latents gives the geometry [B, C, H, W] (5-D video latents [B, C, T, H, W] are folded to [B, C, T*H, W]);
    -> (loss_total [1] fp32, loss_per_sample [B] fp32, dpred like pred)."""
    shp = latents.shape if latents is not None else target.shape
    if len(shp) == 5:
        B, C, H, W = shp[0], shp[1], shp[2] * shp[3], shp[4]
    else:
        B, C, H, W = shp
    if want_grad and dpred is None:
        dpred = torch.empty_like(pred)
    if loss_ws is None:
        loss_ws = torch.empty(B + 1, device=pred.device, dtype=torch.float32)
    per, tot = loss_ws[:B], loss_ws[B:B + 1]
    mc = 0
    if mask is not None:
        assert mask.dtype == torch.float32 and mask.is_contiguous()
        mc = int(mask.shape[1])
    cabi.call("b200_train_loss", _p(pred), _p(latents), _p(noise), _p(target), _p(coef_noise), _p(coef_latent),
              _p(sample_weight), _p(mask), mc, _p(dpred if want_grad else None), _p(per), _p(tot), int(B), int(C), int(H),
              int(W), int(bool(pack)) | (2 if target_in_pred_layout else 0), float(gscale), device=_dev(pred))
    return tot, per, dpred


def ddpm_add_noise(latents, noise, timesteps_i64, alphas_cumprod, out=None):
    """DDPMScheduler.add_noise (toolkit/sampler.py:31-50 config) on bf16 latents with integer timesteps."""
    B = latents.shape[0]
    per = latents.numel() // B
    out = torch.empty_like(latents) if out is None else out
    assert timesteps_i64.dtype == torch.int64 and alphas_cumprod.dtype == torch.float32
    cabi.call("b200_ddpm_add_noise", _p(latents), _p(noise), _p(timesteps_i64), _p(alphas_cumprod),
              int(alphas_cumprod.numel()), _p(out), int(B), int(per), device=_dev(latents))
    return out


def nchw_to_rows(x, out=None):
    """[B, C, H, W] bf16 contiguous -> rows [B*H*W, C] (channels-last rows: the operand layout of the fused GEMM)."""
    B, C, H, W = x.shape
    out = torch.empty((B * H * W, C), device=x.device, dtype=torch.bfloat16) if out is None else out
    cabi.call("b200_nchw_rows", _p(x), _p(out), int(B), int(C), int(H * W), int(out.stride(0)), 1, device=_dev(x))
    return out


def rows_to_nchw(rows, B, C, H, W, out=None):
    out = torch.empty((B, C, H, W), device=rows.device, dtype=torch.bfloat16) if out is None else out
    cabi.call("b200_nchw_rows", _p(rows), _p(out), int(B), int(C), int(H * W), int(rows.stride(0)), 0, device=_dev(rows))
    return out


def conv_out_hw(H, W, k, s, p):
    return (H + 2 * p[0] - k[0]) // s[0] + 1, (W + 2 * p[1] - k[1]) // s[1] + 1


def im2col(x, k, s, p, out=None):
    """[B, C, H, W] bf16 -> rows [B*Ho*Wo, ld] with ld = C*kh*kw rounded up to 8 (zero columns), order (c, ky, kx)."""
    B, C, H, W = x.shape
    Ho, Wo = conv_out_hw(H, W, k, s, p)
    ld = (C * k[0] * k[1] + 7) // 8 * 8
    out = torch.empty((B * Ho * Wo, ld), device=x.device, dtype=torch.bfloat16) if out is None else out
    cabi.call("b200_im2col", _p(x), _p(out), int(B), int(C), int(H), int(W), int(k[0]), int(k[1]), int(s[0]), int(s[1]),
              int(p[0]), int(p[1]), int(out.stride(0)), device=_dev(x))
    return out


def col2im(dcols, shape, k, s, p, out=None, accumulate=False):
    B, C, H, W = shape
    out = torch.empty((B, C, H, W), device=dcols.device, dtype=torch.bfloat16) if out is None else out
    cabi.call("b200_col2im", _p(dcols), _p(out), int(B), int(C), int(H), int(W), int(k[0]), int(k[1]), int(s[0]), int(s[1]),
              int(p[0]), int(p[1]), int(dcols.stride(0)), int(bool(accumulate)), device=_dev(dcols))
    return out


def mask_rows(z, cols, row_mask=None, col_mask=None, rows_per_sample=0):
    """z[:, :cols] *= row_mask [rows, cols] * col_mask [samples, cols] (fp32, either may be None), in place."""
    cabi.call("b200_mask_rows", _p(z), int(z.stride(0)), _p(row_mask), 0 if row_mask is None else int(row_mask.stride(0)),
              _p(col_mask), int(rows_per_sample), int(z.shape[0]), int(cols), device=_dev(z))
    return z


def grad_sumsq(g, out_f64):
    cabi.call("b200_grad_sumsq", _p(g), g.numel(), _p(out_f64), device=_dev(g))


def clip_adamw(p, g, m, v, sumsq_f64, hyper, state, ema=None, norm_out=None):
    cabi.call("b200_clip_adamw", _p(p), _p(g), _p(m), _p(v), _p(ema), _p(sumsq_f64), _p(hyper), _p(state), p.numel(),
              _p(norm_out), device=_dev(p))


def repack_lora(flat, pack, table_dev, n_entries):
    cabi.call("b200_repack_lora", _p(flat), _p(pack), _p(table_dev), int(n_entries), device=_dev(flat))
