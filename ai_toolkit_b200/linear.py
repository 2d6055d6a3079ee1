"""Forward / backward of ONE LoRA-wrapped Linear as launches of the C-ABI kernels.

Shared by the per-module autograd path (`autograd.lora_linear`, the drop-in for
`ToolkitModuleMixin.forward`, toolkit/network_mixins.py:274-348) and by the fused FLUX engine
(`flux_engine`).  With c = multiplier * scale (per sample when the multiplier is a list):

  forward    Zc = bf16(c * X A^T)                        rank-side GEMM, N = 64 (rank padded)
             Y  = epilogue(X W^T + Zc B^T + bias)        ONE tcgen05 tile: base + adapter
  backward   T  = bf16(c * dY B)                         rank-side GEMM (B consumed MN-major)
             dX = epilogue(dY W + T A)                   W and A consumed MN-major: no transposed copies
             dB += dY^T Zc ,  dA += T^T X                token-contraction GEMMs, fp32 atomics into the flat
                                                         gradient buffer; dW [out, in] is never formed
"""
from __future__ import annotations

import torch

from . import cabi, ops
from .cabi import gemm_bf16

RANK_PAD = 64
GEMV_MAX_ROWS = 8


def rank_side(x2, w, out, r_live, *, trans_b=False, alpha=1.0, row_alpha=None, rows_per_sample=0):
    """Zc = bf16(c * X A_pack^T) or (trans_b) T = bf16(c * dY B_pack) on the cluster split-K tcgen05 skinny GEMM.
    (Round 2 measured a CUDA-core kernel for live ranks <= 16 against it -- a warp per 4 rows, weight slice in shared memory,
    packed fp32 FMAs, 1 and 3 chunks of X in flight: 20.5 / 24.6 us vs 15.3 us at M 4608, K 3072, r 16 and 2-3x slower at
    K >= 12288; FLUX step 204.6 / 207.1 vs 195.4 / 191.3 ms -- issue-bound, removed; profiles/r2_time_rank_v*.md.)"""
    return gemm_bf16(x2, w, out, trans_b=trans_b, alpha=alpha, row_alpha=row_alpha, rows_per_sample=rows_per_sample)


def wgrad_splits(tokens: int, features: int, sm_count: int = 148) -> int:
    """Split the token contraction so that (features/128) * splits CTAs fill the GPU."""
    tiles = max(1, (features + 127) // 128)
    kb = max(1, (tokens + 63) // 64)
    s = max(1, min(kb, sm_count // tiles))  # whole waves: tiles * splits <= SMs
    return s


def lora_wgrad(lora, dy, zc, x2, t):
    """dB += dY^T Zc, dA += T^T X for one adapter in ONE launch (zc / t may be column slices of a fused group's Zc / T)."""
    from ctypes import c_void_p

    tokens = dy.shape[0]
    r = lora.lora_dim
    zcols = int(zc.shape[1])
    feats = lora.out_dim + lora.in_dim
    cabi.call("b200_lora_wgrad", c_void_p(dy.data_ptr()), int(dy.stride(0)), c_void_p(zc.data_ptr()), int(zc.stride(0)),
              c_void_p(x2.data_ptr()), int(x2.stride(0)), c_void_p(t.data_ptr()), int(t.stride(0)),
              c_void_p(lora.lora_up.weight.grad.data_ptr()), c_void_p(lora.lora_down.weight.grad.data_ptr()), tokens,
              lora.out_dim, lora.in_dim, r, zcols, 1.0, wgrad_splits(tokens, feats), device=dy.device.index)


def lora_coeff(lora, n_rows: int):
    """-> (alpha scalar, row_alpha tensor | None, rows_per_sample) for this call (no device sync for scalars)."""
    net = lora.network_ref()
    m = net._multiplier
    if isinstance(m, (int, float)):
        return float(m) * lora.scale, None, 0
    if isinstance(m, list) and len(m) == 1 and isinstance(m[0], (int, float)):
        return float(m[0]) * lora.scale, None, 0
    tm = net.torch_multiplier
    if tm.numel() == 1:
        return float(tm.item()) * lora.scale, None, 0
    nb = tm.numel()
    if n_rows % nb != 0:
        raise ValueError(f"{lora.lora_name}: {n_rows} rows cannot be split over {nb} per-sample multipliers")
    return lora.scale, tm.to(torch.float32).contiguous(), n_rows // nb


def linear_fwd(lin, x2, out, *, lora=None, zc_out=None, zc_hook=None, **epi):
    """out = epilogue(x2 @ W^T [+ adapter] + bias).  Returns Zc (saved for backward) or None."""
    W = lin.weight if lin.weight.dim() == 2 else lin.weight.view(lin.weight.shape[0], -1)
    if lora is None:
        gemm_bf16(x2, W, out, bias=lin.bias, **epi)
        return None
    alpha, row_alpha, rps = lora_coeff(lora, x2.shape[0])
    zc = zc_out if zc_out is not None else torch.empty((x2.shape[0], RANK_PAD), device=x2.device, dtype=torch.bfloat16)
    rank_side(x2, lora.a_pack, zc, lora.lora_dim, alpha=alpha, row_alpha=row_alpha, rows_per_sample=rps)
    if zc_hook is not None:  # dropout / rank-dropout masks on the rank-side activation (network_mixins.py:211-226)
        zc_hook(zc)
    gemm_bf16(x2, W, out, a1=zc, b1=lora.b_pack, bias=lin.bias, **epi)
    return zc


def linear_bwd(lin, dy, x2, zc, dx_out, *, lora=None, n_slices=None, t_hook=None, **epi):
    """dX (into dx_out, may be None) and, with an adapter, accumulation of dA / dB into their `.grad` views.

    n_slices: optional list of (col0, col1, epi_kwargs) to produce dX in column ranges with different epilogues
    (the single-stream block applies gelu' only to the MLP part of the concatenated operand)."""
    W = lin.weight if lin.weight.dim() == 2 else lin.weight.view(lin.weight.shape[0], -1)
    t = None
    if lora is not None:
        alpha, row_alpha, rps = lora_coeff(lora, dy.shape[0])
        t = torch.empty((dy.shape[0], RANK_PAD), device=dy.device, dtype=torch.bfloat16)
        rank_side(dy, lora.b_pack, t, lora.lora_dim, trans_b=True, alpha=alpha, row_alpha=row_alpha, rows_per_sample=rps)
        if t_hook is not None:
            t_hook(t)
    if dx_out is not None:
        slices = n_slices if n_slices is not None else [(0, W.shape[1], epi)]
        for c0, c1, e in slices:
            a1 = t if lora is not None else None
            b1 = lora.a_pack[:, c0:c1] if lora is not None else None
            gemm_bf16(dy, W[:, c0:c1], dx_out[:, c0:c1], a1=a1, b1=b1, trans_b=True, N=c1 - c0, **e)
    if lora is not None:
        lora_wgrad(lora, dy, zc, x2, t)
    return t


def live_lora(lin):
    """The active adapter of a Linear, or None (inactive network, merged in, multiplier 0, or not wrapped)."""
    ref = getattr(lin, "_b200_lora", None)
    lora = ref() if ref is not None else None
    if lora is None or not lora.is_live():
        return None
    return lora


# ------------------------------------------------------------------------------------------------------
# adapters that share one input (to_q / to_k / to_v ...): one rank-side GEMM + one fused GEMM for the group
# ------------------------------------------------------------------------------------------------------
def fuse_linear_weights(lins):
    """Re-point the weights / biases of Linear layers that share an input at row blocks of ONE contiguous
    [sum out, in] matrix (no extra copy afterwards; state_dict / load_state_dict keep working on the views)."""
    first = lins[0]
    cached = getattr(first, "_b200_fused_w", None)
    if cached is not None and cached[0].data_ptr() == first.weight.data_ptr():
        return cached
    with torch.no_grad():
        W = torch.cat([l.weight.detach() for l in lins], 0).contiguous()
        bias = torch.cat([l.bias.detach() for l in lins], 0).contiguous() if first.bias is not None else None
        row = 0
        for l in lins:
            n = l.weight.shape[0]
            l.weight.data = W[row:row + n]
            if bias is not None:
                l.bias.data = bias[row:row + n]
            row += n
    first._b200_fused_w = (W, bias)
    return W, bias


def group_fwd(group, lins, x2, out, **epi):
    """out[M, sum out] = x2 @ [W_0; W_1; ...]^T + Zc @ B_fused^T + bias, Zc = bf16(c * x2 @ A_fused^T)."""
    W, bias = fuse_linear_weights(lins)
    alpha, row_alpha, rps = lora_coeff(group.loras[0], x2.shape[0])
    zc = torch.empty((x2.shape[0], RANK_PAD), device=x2.device, dtype=torch.bfloat16)
    rank_side(x2, group.a_fused, zc, group.r * len(group.loras), alpha=alpha, row_alpha=row_alpha, rows_per_sample=rps)
    gemm_bf16(x2, W, out, a1=zc, b1=group.b_fused, bias=bias, **epi)
    return zc


def group_bwd(group, lins, dy, x2, zc, dx_out, **epi):
    """dX for the whole group in one GEMM (K = sum out), dA / dB per adapter from column slices of T / Zc."""
    W, _ = fuse_linear_weights(lins)
    alpha, row_alpha, rps = lora_coeff(group.loras[0], dy.shape[0])
    t = torch.empty((dy.shape[0], RANK_PAD), device=dy.device, dtype=torch.bfloat16)
    rank_side(dy, group.b_fused, t, group.r * len(group.loras), trans_b=True, alpha=alpha, row_alpha=row_alpha, rows_per_sample=rps)
    if dx_out is not None:
        gemm_bf16(dy, W, dx_out, a1=t, b1=group.a_fused, trans_b=True, **epi)
    r = group.r
    row = 0
    for j, lora in enumerate(group.loras):
        n = lora.out_dim
        lora_wgrad(lora, dy[:, row:row + n], zc[:, j * r:], x2, t[:, j * r:])
        row += n


def shared_input_fwd(lins, n, out):
    """Projections of ONE input (to_q / to_k / to_v, or the to_k / to_v of a cross attention) -> column blocks of `out`:
    one rank-side product + one fused GEMM when their adapters form a registered group, else one pair per Linear.
    Returns the token to hand to `shared_input_bwd`."""
    loras = [live_lora(l) for l in lins]
    grp = loras[0].network_ref().fused_group(loras) if loras[0] is not None else None
    if grp is not None:
        return ("g", group_fwd(grp, lins, n, out))
    col, zs = 0, []
    for lin, lo in zip(lins, loras):
        D = lin.weight.shape[0]
        zs.append(linear_fwd(lin, n, out[:, col:col + D], lora=lo))
        col += D
    return ("s", zs)


def shared_input_bwd(lins, dout, n, saved, dn, accumulate=False):
    """Backward of `shared_input_fwd`: dn (+)= sum_j dout_j W_j (+ adapters), dA / dB of every adapter; dn may be None."""
    loras = [live_lora(l) for l in lins]
    kind, z = saved
    if kind == "g":
        grp = loras[0].network_ref().fused_group(loras)
        group_bwd(grp, lins, dout, n, z, dn, **({"res": dn} if (accumulate and dn is not None) else {}))
        return
    col = 0
    for j, (lin, lo) in enumerate(zip(lins, loras)):
        D = lin.weight.shape[0]
        acc = dn is not None and (accumulate or j > 0)
        linear_bwd(lin, dout[:, col:col + D], n, z[j], dn, lora=lo, **({"res": dn} if acc else {}))
        col += D
