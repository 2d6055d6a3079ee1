import sys, torch
sys.path.insert(0, ".")
from ai_toolkit_b200 import cabi
dev = "cuda:0"
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for (M, K, tb) in ((4608, 3072, False), (4608, 3072, True), (4608, 9216, True), (4608, 21504, True), (4096, 12288, True), (512, 9216, True), (4608, 15360, False), (4096, 12288, False)):
    a = torch.randn(M, K, device=dev).bfloat16()
    b = (torch.randn(K, 64, device=dev) if tb else torch.randn(64, K, device=dev)).bfloat16()
    o = torch.empty(M, 64, device=dev, dtype=torch.bfloat16)
    res = []
    for cfg in (4, 5):
        for _ in range(3): cabi.gemm_bf16(a, b, o, trans_b=tb, config=cfg)
        e0.record()
        for _ in range(50): cabi.gemm_bf16(a, b, o, trans_b=tb, config=cfg)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 20)
    print(f"M={M} K={K} trans_b={tb}: persistent {res[0]:.1f} us, cluster {res[1]:.1f} us")
