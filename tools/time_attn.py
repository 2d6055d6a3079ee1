import sys, torch
sys.path.insert(0, ".")
from ai_toolkit_b200 import attention, cabi
dev = "cuda:0"
B, H, L, split = 1, 24, 4608, 512
Q, K, V = (torch.randn(B, H, L, 128, device=dev).bfloat16() for _ in range(3))
o0 = torch.empty(B * split, H * 128, device=dev, dtype=torch.bfloat16)
o1 = torch.empty(B * (L - split), H * 128, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    lse = attention.fwd(Q, K, V, o0, o1, split)
ref = torch.nn.functional.scaled_dot_product_attention(Q, K, V).transpose(1, 2).reshape(B, L, H * 128)
got = torch.cat([o0.view(B, split, -1), o1.view(B, L - split, -1)], 1)
err = ((got.float() - ref.float()).norm() / ref.float().norm()).item()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    attention.fwd(Q, K, V, o0, o1, split)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 200
print(f"{cabi.LIB_PATH.split('/')[-1]}: fwd {ms*1e3:.1f} us = {4*B*H*L*L*128/ms/1e9:.0f} TFLOP/s (sustained loop), rel err vs SDPA {err:.2e}")
# skinny Z GEMM timing (hot L2)
x = torch.randn(4608, 3072, device=dev).bfloat16(); ap = torch.randn(64, 3072, device=dev).bfloat16(); z = torch.empty(4608, 64, device=dev, dtype=torch.bfloat16)
for cfg in (0, 4):
    for _ in range(3): cabi.gemm_bf16(x, ap, z, config=cfg)
    e0.record()
    for _ in range(100): cabi.gemm_bf16(x, ap, z, config=cfg)
    e1.record(); torch.cuda.synchronize()
    print(f"  Z gemm 4608x64x3072 config {cfg}: {e0.elapsed_time(e1)*10:.1f} us")
