"""Sustained (power-capped) throughput of the hot kernels: run each back-to-back for ~2 s, report TFLOP/s and the
SM clock / power seen by nvidia-smi meanwhile; cuBLAS bf16 at the same shape for reference."""
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from ai_toolkit_b200 import attention, cabi  # noqa: E402

dev = torch.device("cuda:0")


class Smi(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop = [], False

    def run(self):
        while not self.stop:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True).stdout.strip().split(",")
            try:
                self.rows.append((float(o[0]), float(o[1])))
            except Exception:
                pass
            time.sleep(0.1)


def run(name, fn, flops, secs=2.0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    smi = Smi()
    smi.start()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize() if n % 200 == 0 else None
    e1.record()
    torch.cuda.synchronize()
    smi.stop = True
    ms = e0.elapsed_time(e1) / n
    rows = smi.rows[len(smi.rows) // 3:]
    clk = sorted(r[0] for r in rows)[len(rows) // 2] if rows else 0
    pw = max(r[1] for r in rows) if rows else 0
    print(f"{name:34s} {ms*1e3:8.1f} us  {flops/ms/1e9:7.0f} TFLOP/s  sm {clk:.0f} MHz  {pw:.0f} W", flush=True)


M, N, K = 4608, 12288, 3072
x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
zc = (torch.randn(M, 64, device=dev) * 0.1).bfloat16()
bp = (torch.randn(N, 64, device=dev) * 0.02).bfloat16()
bias = torch.zeros(N, device=dev, dtype=torch.bfloat16)
y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
run("cuBLAS bf16 4608x12288x3072", lambda: torch.matmul(x, w.t(), out=y), 2.0 * M * N * K)
run("fused LoRA-Linear (bias)", lambda: cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, bias=bias), 2.0 * M * N * (K + 64))
run("fused LoRA-Linear (bias,gelu,aux)", lambda: cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, bias=bias, act=1, aux_out=pre), 2.0 * M * N * (K + 64))
dy = torch.randn(M, N, device=dev).bfloat16()
t = (torch.randn(M, 64, device=dev) * 0.1).bfloat16()
ap = (torch.randn(64, K, device=dev) * 0.02).bfloat16()
dx = torch.empty(M, K, device=dev, dtype=torch.bfloat16)
run("cuBLAS bf16 dgrad 4608x3072x12288", lambda: torch.matmul(dy, w, out=dx), 2.0 * M * N * K)
run("dgrad (W MN-major) + LoRA", lambda: cabi.gemm_bf16(dy, w, dx, a1=t, b1=ap, trans_b=True), 2.0 * M * K * (N + 64))
B, H, L, split = 1, 24, 4608, 512
Q, K_, V = (torch.randn(B, H, L, 128, device=dev).bfloat16() for _ in range(3))
o0 = torch.empty(B * split, H * 128, device=dev, dtype=torch.bfloat16)
o1 = torch.empty(B * (L - split), H * 128, device=dev, dtype=torch.bfloat16)
lse = attention.fwd(Q, K_, V, o0, o1, split)
do0, do1 = torch.randn_like(o0), torch.randn_like(o1)
f = 4.0 * B * H * L * L * 128
run("attention fwd", lambda: attention.fwd(Q, K_, V, o0, o1, split), f)
run("attention bwd (algorithmic 2.5x)", lambda: attention.bwd(Q, K_, V, o0, o1, do0, do1, lse, split), 2.5 * f)
run("torch SDPA fwd (library)", lambda: torch.nn.functional.scaled_dot_product_attention(Q, K_, V), f)
