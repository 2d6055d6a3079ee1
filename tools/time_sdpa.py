"""torch SDPA (the reference's attention on B200: F.scaled_dot_product_attention, toolkit/models/wan21/wan_attn.py:70-76)
forward and backward at the FLUX shape, per backend, same-box bar for the hand-written kernels."""
import sys

import torch
from torch.nn.attention import SDPBackend, sdpa_kernel

dev = "cuda:0"
B, H, L = 1, 24, int(sys.argv[1]) if len(sys.argv) > 1 else 4608
fl = 4 * B * H * L * L * 128
for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION), ("default", None)):
    try:
        q, k, v = (torch.randn(B, H, L, 128, device=dev, dtype=torch.bfloat16, requires_grad=True) for _ in range(3))
        do = torch.randn(B, H, L, 128, device=dev, dtype=torch.bfloat16)

        def f():
            if be is None:
                return torch.nn.functional.scaled_dot_product_attention(q, k, v)
            with sdpa_kernel(be):
                return torch.nn.functional.scaled_dot_product_attention(q, k, v)

        for _ in range(3):
            f().backward(do)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        with torch.no_grad():
            e0.record()
            for _ in range(n):
                f()
            e1.record()
        torch.cuda.synchronize()
        tf = e0.elapsed_time(e1) / n
        e0.record()
        for _ in range(n):
            f().backward(do)
        e1.record()
        torch.cuda.synchronize()
        tfb = e0.elapsed_time(e1) / n
        tb = tfb - tf
        print(f"[sdpa {name}] L={L} fwd {tf * 1e3:.1f} us ({fl / tf / 1e9:.0f} TF/s)  bwd {tb * 1e3:.1f} us ({2.5 * fl / tb / 1e9:.0f} TF/s)",
              flush=True)
    except Exception as e:  # backend not available for this shape / build
        print(f"[sdpa {name}] unavailable: {type(e).__name__}: {str(e)[:120]}", flush=True)
