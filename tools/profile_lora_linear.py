"""A few launches of the fused LoRA-Linear (+ its dgrad) at the rank-sweep shapes of BASELINE.json configs[4]
(bs 4: M = 16384 image rows / 18432 joint rows) for `ncu --set full`:  python tools/profile_lora_linear.py RANK SHAPE
SHAPE: mlp_up (16384 x 12288 x 3072, + bias + GELU + saved pre-activation), attn_out (16384 x 3072 x 3072, gate + residual),
single_out (18432 x 3072 x 15360, gate + residual)."""
import sys

import torch

sys.path.insert(0, ".")
from ai_toolkit_b200 import cabi  # noqa: E402

dev = torch.device("cuda:0")
r = int(sys.argv[1])
shape = sys.argv[2] if len(sys.argv) > 2 else "mlp_up"
M, N, K = {"mlp_up": (16384, 12288, 3072), "attn_out": (16384, 3072, 3072), "single_out": (18432, 3072, 15360)}[shape]
x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
zc = torch.zeros(M, 64, device=dev, dtype=torch.bfloat16)
zc[:, :r] = (torch.randn(M, r, device=dev) * 0.1).bfloat16()
bp = torch.zeros(N, 64, device=dev, dtype=torch.bfloat16)
bp[:, :r] = (torch.randn(N, r, device=dev) * 0.02).bfloat16()
bias = torch.zeros(N, device=dev, dtype=torch.bfloat16)
y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
if shape == "mlp_up":
    pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    kw = dict(act=cabi.ACT_GELU_TANH, aux_out=pre)
else:
    kw = dict(gate=(torch.randn(4, N, device=dev) * 0.1).bfloat16(), rows_per_sample=M // 4, res=torch.randn(M, N, device=dev).bfloat16())
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, bias=bias, **kw)
ev0.record()
for _ in range(10):
    cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, bias=bias, **kw)
ev1.record()
torch.cuda.synchronize()
us = ev0.elapsed_time(ev1) * 100
fl = 2.0 * M * N * K + 2.0 * M * r * N
print(f"fused LoRA-Linear {shape} M={M} N={N} K={K} r={r}: {us:.1f} us/launch (hot loop) = {fl / us / 1e6:.0f} TFLOP/s")
