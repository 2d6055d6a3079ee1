"""Parity + timing of ONE attention variant pair (selected with B200_ATTN_FWD / B200_ATTN_BWD in the environment, read
once per process by the library).  Prints one line; used by tools/r2_attn_trip.sh to A/B the round-2 candidates inside a
single gpurun call (box-to-box spread is +-2 %, so only same-box numbers are comparable).

Parity: forward vs torch SDPA (fp32 math on the same bf16 inputs), backward vs autograd of that reference; tolerances are
the ones of tests/test_gpu_attention.py (rel. Frobenius error <= 1e-2 for bf16 outputs)."""
import os
import sys

import torch

sys.path.insert(0, ".")
from ai_toolkit_b200 import attention  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
B, H, split = 1, 24, 512
Ls = [int(a) for a in sys.argv[1:]] or [4608]
tag = f"fwd={os.environ.get('B200_ATTN_FWD', '1')} bwd={os.environ.get('B200_ATTN_BWD', '1')}"
for L in Ls:
    sp = min(split, L // 2)
    Q, K, V = (torch.randn(B, H, L, 128, device=dev).bfloat16() for _ in range(3))
    o0 = torch.empty(B * sp, H * 128, device=dev, dtype=torch.bfloat16)
    o1 = torch.empty(B * (L - sp), H * 128, device=dev, dtype=torch.bfloat16)
    do0, do1 = torch.randn_like(o0), torch.randn_like(o1)
    lse = attention.fwd(Q, K, V, o0, o1, sp)
    dQ, dK, dV = attention.bwd(Q, K, V, o0, o1, do0, do1, lse, sp)
    torch.cuda.synchronize()
    # reference in fp32
    q, k, v = (t.float().requires_grad_(True) for t in (Q, K, V))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v)           # [B, H, L, 128]
    ref_tok = ref.transpose(1, 2).reshape(B, L, H * 128)
    got = torch.cat([o0.view(B, sp, -1), o1.view(B, L - sp, -1)], 1)
    dO = torch.cat([do0.view(B, sp, -1), do1.view(B, L - sp, -1)], 1).float()
    ref_tok.backward(dO)
    rel = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()  # noqa: E731
    errs = dict(o=rel(got, ref_tok), dq=rel(dQ, q.grad), dk=rel(dK, k.grad), dv=rel(dV, v.grad))
    ok = all(e <= 1e-2 for e in errs.values()) and all(torch.isfinite(t.float()).all().item() for t in (got, dQ, dK, dV))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 100 if L >= 2048 else 300
    for _ in range(5):
        attention.fwd(Q, K, V, o0, o1, sp)
    e0.record()
    for _ in range(n):
        attention.fwd(Q, K, V, o0, o1, sp)
    e1.record()
    torch.cuda.synchronize()
    t_f = e0.elapsed_time(e1) / n
    for _ in range(5):
        attention.bwd(Q, K, V, o0, o1, do0, do1, lse, sp)
    e0.record()
    for _ in range(n):
        attention.bwd(Q, K, V, o0, o1, do0, do1, lse, sp)
    e1.record()
    torch.cuda.synchronize()
    t_b = e0.elapsed_time(e1) / n
    fl = 4 * B * H * L * L * 128
    print(f"[{tag}] L={L} parity={'OK' if ok else 'FAIL'} " + " ".join(f"{k}={v:.2e}" for k, v in errs.items())
          + f" | fwd {t_f * 1e3:.1f} us ({fl / t_f / 1e9:.0f} TF/s)  bwd {t_b * 1e3:.1f} us ({2.5 * fl / t_b / 1e9:.0f} TF/s)",
          flush=True)
