import sys, time, torch
sys.path.insert(0, ".")
from ai_toolkit_b200 import cabi
dev = torch.device("cuda:0")
M, N, K = 4608, 12288, 3072
x = (torch.randn(M, K, device=dev) * 0.5).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
zc = (torch.randn(M, 64, device=dev) * 0.1).bfloat16(); bp = (torch.randn(N, 64, device=dev) * 0.02).bfloat16()
bias = torch.zeros(N, device=dev, dtype=torch.bfloat16)
y = torch.empty(M, N, device=dev, dtype=torch.bfloat16); pre = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
res = torch.randn(M, N, device=dev).bfloat16(); gate = torch.randn(1, N, device=dev).bfloat16()
def run(name, fn, secs=1.5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    while time.time() - t0 < secs:
        for _ in range(20): fn()
        n += 20
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:36s} {ms*1e3:7.1f} us {2.0*M*N*(K+64)/ms/1e9:6.0f} TF/s", flush=True)
run("plain (no bias)", lambda: cabi.gemm_bf16(x, w, y, a1=zc, b1=bp))
run("bias", lambda: cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, bias=bias))
run("bias+gelu", lambda: cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, bias=bias, act=1))
run("bias+aux_out", lambda: cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, bias=bias, aux_out=pre))
run("bias+gelu+aux_out", lambda: cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, bias=bias, act=1, aux_out=pre))
run("bias+gate+res+aux_out", lambda: cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, bias=bias, gate=gate, rows_per_sample=M, res=res, aux_out=pre))
run("aux_in (gelu')", lambda: cabi.gemm_bf16(x, w, y, a1=zc, b1=bp, aux_in=pre))
