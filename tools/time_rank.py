"""Rank-side products, CUDA-core kernel vs the tensor-core skinny GEMM, at the FLUX / SDXL shapes (L2 flushed and back-to-back)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_toolkit_b200 import cabi, ops  # noqa: E402

dev = "cuda:0"
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)


def timeit(fn, n=20, cold=True):
    for _ in range(3):
        fn()
    if not cold:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3 / 50
    t = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        t.append(a.elapsed_time(b) * 1e3)
    t.sort()
    return t[len(t) // 2]


print("| M | K | r | op | simt cold us | tensor cold us | simt warm us | tensor warm us | X MB |")
print("|---|---|---|---|---|---|---|---|---|")
for M, K, r, trans in [(4608, 3072, 16, False), (4608, 12288, 16, False), (4608, 15360, 16, False), (4608, 3072, 16, True),
                       (4608, 12288, 16, True), (16384, 3072, 16, False), (16384, 3072, 4, False), (2048, 1280, 8, False),
                       (2048, 1280, 8, True), (2048, 10240, 8, True), (2048, 5120, 8, False), (8192, 640, 8, False), (8192, 5120, 8, True),
                       (154, 2048, 8, False), (154, 1280, 8, True)]:
    x = torch.randn(M, K, device=dev).bfloat16()
    w = torch.zeros((K, 64) if trans else (64, K), device=dev)
    if trans:
        w[:, :r] = torch.randn(K, r, device=dev) * 0.05
    else:
        w[:r] = torch.randn(r, K, device=dev) * 0.05
    w = w.bfloat16()
    out = torch.empty(M, 64, device=dev, dtype=torch.bfloat16)
    f1 = lambda: ops.rank_gemm(x, w, out, r, trans_w=trans)  # noqa: E731
    f2 = lambda: cabi.gemm_bf16(x, w, out, trans_b=trans)  # noqa: E731
    print(f"| {M} | {K} | {r} | {'T=dY.B' if trans else 'Z=X.A^T'} | {timeit(f1):.1f} | {timeit(f2):.1f} | {timeit(f1, cold=False):.1f} | "
          f"{timeit(f2, cold=False):.1f} | {M * K * 2 / 1e6:.1f} |")
