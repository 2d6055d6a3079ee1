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
