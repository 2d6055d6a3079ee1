"""A few launches of the hot kernels at FLUX shapes for `ncu --set full` (tools/gpu_profile_trip.sh)."""
import sys

import torch

sys.path.insert(0, ".")
from ai_toolkit_b200 import attention, cabi  # noqa: E402

dev = torch.device("cuda:0")
which = sys.argv[1]
if which == "gemm":
    M, N, K = 4608, 12288, 3072
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    zc = (torch.randn(M, 64, device=dev) * 0.1).bfloat16()
    bp = (torch.randn(N, 64, device=dev) * 0.02).bfloat16()
    bias = torch.zeros(N, device=dev, dtype=torch.bfloat16)
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, bias=bias, act=1, aux_out=pre)
    # dgrad (W consumed MN-major) at the same layer
    dy = torch.randn(M, N, device=dev).bfloat16()
    t = (torch.randn(M, 64, device=dev) * 0.1).bfloat16()
    ap = (torch.randn(64, K, device=dev) * 0.02).bfloat16()
    dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        cabi.gemm_bf16(dy, w, dx, a1=t, b1=ap, trans_b=True)
elif which == "r2new":
    # kernels added / changed late in round 2, one warm + one measured launch each (ncu -c picks them up in order):
    from ai_toolkit_b200 import ops
    # (1) 128 x 160 tile GEMM at SDXL's [2048, 1280, 1280] with the LoRA segment, bias and residual; (2) its 128 x 192 dgrad
    M, N, K = 2048, 1280, 1280
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    zc = (torch.randn(M, 64, device=dev) * 0.1).bfloat16()
    bp = (torch.randn(N, 64, device=dev) * 0.02).bfloat16()
    bias = torch.zeros(N, device=dev, dtype=torch.bfloat16)
    res = torch.randn(M, N, device=dev).bfloat16()
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, bias=bias, res=res)
    dy = torch.randn(M, N, device=dev).bfloat16()
    ap = (torch.randn(64, K, device=dev) * 0.02).bfloat16()
    dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        cabi.gemm_bf16(dy, w, dx, a1=zc, b1=ap, trans_b=True)
    # (3) ln_modulate_bwd (rows staged in shared memory) and (4) col_reduce at the FLUX stream shape
    Mf, D = 4608, 3072
    xf = torch.randn(Mf, D, device=dev).bfloat16()
    dyf = torch.randn(Mf, D, device=dev).bfloat16()
    dres = torch.randn(Mf, D, device=dev).bfloat16()
    mod = (torch.randn(1, 6 * D, device=dev) * 0.3).bfloat16()
    _, mean, rstd = ops.ln_modulate_fwd(xf, mod[:, :D], mod[:, D:2 * D], Mf)
    dmod = torch.zeros(1, 6 * D, device=dev)
    for _ in range(2):
        ops.ln_modulate_bwd(dyf, xf, mean, rstd, mod[:, D:2 * D], Mf, dres=dres)
    for _ in range(2):
        ops.col_reduce(dyf, Mf, b=xf, mean=mean, rstd=rstd, sum_a=dmod[:, :D], sum_ab=dmod[:, D:2 * D])
    # (5) attention at SDXL's 64 x 64 level with head_live = 64 (forward + the two backward kernels)
    B, H, L = 2, 10, 4096
    def pad(t):
        o = torch.zeros(*t.shape[:-1], 128, device=dev, dtype=torch.bfloat16)
        o[..., :64] = t
        return o
    Q, K_, V = (pad(torch.randn(B, H, L, 64, device=dev).bfloat16()) for _ in range(3))
    o1 = torch.empty(B * L, H * 128, device=dev, dtype=torch.bfloat16)
    dO = pad(torch.randn(B, L, H, 64, device=dev).bfloat16()).reshape(B * L, H * 128)
    for _ in range(2):
        lse = attention.fwd(Q, K_, V, None, o1, 0, scale=0.125, head_live=64)
    for _ in range(2):
        attention.bwd(Q, K_, V, None, o1, None, dO, lse, 0, scale=0.125, head_live=64)
    # (6) CUDA-core attention, head dim 160 (SD1.5 at 512^2: 8 heads x 256 tokens)
    q, k, v = (torch.randn(256, 1280, device=dev).bfloat16() for _ in range(3))
    for _ in range(2):
        o_s, lse_s = attention.small_fwd(q, k, v, 1, 8, 256, 256, 160)
    dq, dk, dv = (torch.empty_like(q) for _ in range(3))
    for _ in range(2):
        attention.small_bwd(q, k, v, o_s, torch.randn_like(q), lse_s, dq, dk, dv, 1, 8, 256, 256, 160)
else:
    B, H, L, split = 1, 24, 4608, 512
    Q, K_, V = (torch.randn(B, H, L, 128, device=dev).bfloat16() for _ in range(3))
    o0 = torch.empty(B * split, H * 128, device=dev, dtype=torch.bfloat16)
    o1 = torch.empty(B * (L - split), H * 128, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        lse = attention.fwd(Q, K_, V, o0, o1, split)
    do0, do1 = torch.randn_like(o0), torch.randn_like(o1)
    for _ in range(2):
        attention.bwd(Q, K_, V, o0, o1, do0, do1, lse, split)
torch.cuda.synchronize()
