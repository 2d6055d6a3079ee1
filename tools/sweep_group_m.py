"""Sweep the GEMM rasterisation group height (B200_GEMM_GROUP_M) for the main FLUX GEMM shapes, sustained loop."""
import os, sys, time, torch
sys.path.insert(0, ".")
from ai_toolkit_b200 import cabi
dev = torch.device("cuda:0")
def run(fn, flops, secs=0.6):
    for _ in range(3): fn()
    torch.cuda.synchronize(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    while time.time() - t0 < secs:
        for _ in range(10): fn()
        n += 10
    e1.record(); torch.cuda.synchronize()
    return flops / (e0.elapsed_time(e1) / n) / 1e9
shapes = [("fwd  M4608 N12288 K3072", 4608, 12288, 3072, False), ("fwd  M4608 N21504 K3072", 4608, 21504, 3072, False),
          ("fwd  M4096 N3072 K12288", 4096, 3072, 12288, False), ("fwd  M4608 N3072 K15360", 4608, 3072, 15360, False),
          ("dgrad M4608 N3072 K12288", 4608, 3072, 12288, True), ("dgrad M4608 N3072 K21504", 4608, 3072, 21504, True),
          ("dgrad M4608 N15360 K3072", 4608, 15360, 3072, True), ("dgrad M4096 N12288 K3072", 4096, 12288, 3072, True)]
for name, M, N, K, tb in shapes:
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    b = ((torch.randn(K, N, device=dev) if tb else torch.randn(N, K, device=dev)) * 0.02).bfloat16()
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    res = []
    for gm in (0, 2, 3, 4, 6, 9, 18):
        if gm: os.environ["B200_GEMM_GROUP_M"] = str(gm)
        else: os.environ.pop("B200_GEMM_GROUP_M", None)
        res.append((gm, run(lambda: cabi.gemm_bf16(a, b, y, trans_b=tb), 2.0 * M * N * K)))
    os.environ.pop("B200_GEMM_GROUP_M", None)
    print(name, " ".join(f"gm{g}:{t:.0f}" for g, t in res), flush=True)
