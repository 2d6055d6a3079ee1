"""Attention at SDXL's shapes (head dim 64 zero-padded to 128): head_live 128 vs 64, forward and backward, CUDA events."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_toolkit_b200 import attention  # noqa: E402

dev = "cuda:0"
print("| B | H | L | Lk | live | fwd us | bwd us |\n|---|---|---|---|---|---|---|")
for B, H, L, Lk in [(2, 10, 4096, 4096), (2, 20, 1024, 1024), (2, 10, 4096, 77), (2, 20, 1024, 77)]:
    def pad(x):
        o = torch.zeros(*x.shape[:-1], 128, device=dev, dtype=torch.bfloat16)
        o[..., :64] = x
        return o
    Q, K, V = pad(torch.randn(B, H, L, 64, device=dev).bfloat16()), pad(torch.randn(B, H, Lk, 64, device=dev).bfloat16()), pad(torch.randn(B, H, Lk, 64, device=dev).bfloat16())
    dO = pad(torch.randn(B, L, H, 64, device=dev).bfloat16()).reshape(B * L, H * 128)
    o1 = torch.empty(B * L, H * 128, device=dev, dtype=torch.bfloat16)
    for live in (128, 64):
        def f():
            return attention.fwd(Q, K, V, None, o1, 0, scale=0.125, head_live=live)
        lse = f()
        def b():
            return attention.bwd(Q, K, V, None, o1, None, dO, lse, 0, scale=0.125, head_live=live)
        ts = []
        for fn in (f, b):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 20)
        print(f"| {B} | {H} | {L} | {Lk} | {live} | {ts[0]:.1f} | {ts[1]:.1f} |")
