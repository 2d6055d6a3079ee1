"""GPU diagnostic for the tcgen05 GEMM (run on the B200 box through gpurun; one config per process
so that a trap in one configuration does not hide the others).

usage: python tools/gpu_check_gemm.py <config:int> [quick]
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from ai_toolkit_b200 import cabi  # noqa: E402


def ref_gemm(a0, b0, a1=None, b1=None, bias=None, act=0, gate=None, rps=0, res=None, aux_in=None):
    acc = a0.float() @ b0.float().t()
    if a1 is not None:
        acc = acc + a1.float() @ b1.float().t()
    if bias is not None:
        acc = acc + bias.float()
    y = acc.bfloat16()
    pre = y
    if act == 1:
        y = torch.nn.functional.gelu(y.float(), approximate="tanh").bfloat16()
    if aux_in is not None:
        x = aux_in.float().requires_grad_(True)
        gx = torch.autograd.grad(torch.nn.functional.gelu(x, approximate="tanh").sum(), x)[0]
        y = (y.float() * gx).bfloat16()
    if gate is not None:
        g = gate.float().repeat_interleave(rps, dim=0)[: y.shape[0]]
        y = (y.float() * g).bfloat16()
    if res is not None:
        y = (y.float() + res.float()).bfloat16()
    return y, pre


def err(a, b):
    a = a.float()
    b = b.float()
    d = (a - b).abs()
    return d.max().item(), (d.norm() / (b.norm() + 1e-30)).item()


def main():
    config = int(sys.argv[1])
    quick = len(sys.argv) > 2
    torch.manual_seed(0)
    dev = torch.device("cuda:0")
    print("device", torch.cuda.get_device_name(0), "config", config, flush=True)
    cases = [
        # M, N, K0, K1, flags
        (128, 256, 64, 0, ""),
        (128, 256, 256, 0, ""),
        (256, 256, 128, 0, ""),
        (256, 512, 3072, 0, ""),
        (384, 768, 512, 64, "bias"),
        (200, 264, 328, 64, "bias"),          # ragged M/N/K
        (512, 3072, 3072, 64, "bias,gelu,auxout"),
        (512, 3072, 3072, 64, "bias,gate,res"),
        (512, 3072, 3072, 0, "auxin"),
        (4608, 3072, 3072, 64, "bias"),
        (4608, 12288, 3072, 64, "bias,gelu,auxout"),
        (4608, 3072, 15360, 64, "bias,gate,res"),
    ]
    if config == cabi.GEMM_1CTA_N64:
        cases = [(128, 64, 64, 0, ""), (256, 64, 512, 0, ""), (200, 48, 328, 0, ""), (4608, 64, 3072, 0, ""),
                 (4608, 64, 3072, 0, "f32split4"), (4608, 64, 15360, 0, "f32split8"), (512, 16, 3072, 0, "f32split3")]
    if quick:
        cases = cases[:4]
    ok_all = True
    for (M, N, K0, K1, flags) in cases:
        a0 = (torch.randn(M, K0, device=dev) * 0.5).bfloat16()
        b0 = (torch.randn(N, K0, device=dev) * 0.05).bfloat16()
        a1 = (torch.randn(M, K1, device=dev) * 0.5).bfloat16() if K1 else None
        b1 = (torch.randn(N, K1, device=dev) * 0.05).bfloat16() if K1 else None
        bias = torch.randn(N, device=dev).bfloat16() if "bias" in flags else None
        act = 1 if "gelu" in flags else 0
        rps = 0
        gate = res = aux_in = aux_out = None
        if "gate" in flags:
            rps = max(1, M // 2)
            gate = torch.randn((M + rps - 1) // rps, N, device=dev).bfloat16()
        if "res" in flags:
            res = torch.randn(M, N, device=dev).bfloat16()
        if "auxin" in flags:
            aux_in = torch.randn(M, N, device=dev).bfloat16()
        if "auxout" in flags:
            aux_out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        splits = 1
        out_f32 = False
        if "f32split" in flags:
            splits = int(flags.split("f32split")[1])
            out_f32 = True
            out = torch.full((splits, M, N), float("nan"), device=dev, dtype=torch.float32)
        else:
            out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
        cabi.gemm_bf16(a0, b0, out, a1=a1, b1=b1, bias=bias, res=res, gate=gate, rows_per_sample=rps, aux_in=aux_in,
                       aux_out=aux_out, act=act, out_f32=out_f32, splits=splits, config=config)
        torch.cuda.synchronize()
        if out_f32:
            got = out.sum(0)
            want = a0.float() @ b0.float().t()
            mx, rel = err(got, want)
            tol = 2e-5
        else:
            want, pre = ref_gemm(a0, b0, a1, b1, bias, act, gate, rps, res, aux_in)
            mx, rel = err(out, want)
            tol = 6e-3
            if aux_out is not None:
                mx2, rel2 = err(aux_out, pre)
                print(f"   aux_out max {mx2:.4g} rel {rel2:.3g}")
                ok_all &= rel2 < tol
        # timing
        torch.cuda.synchronize()
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        iters = 10
        for _ in range(3):
            cabi.gemm_bf16(a0, b0, out, a1=a1, b1=b1, bias=bias, res=res if res is not out else res, gate=gate,
                           rows_per_sample=rps, aux_in=aux_in, aux_out=aux_out, act=act, out_f32=out_f32,
                           splits=splits, config=config)
        ev0.record()
        for _ in range(iters):
            cabi.gemm_bf16(a0, b0, out, a1=a1, b1=b1, bias=bias, res=res, gate=gate, rows_per_sample=rps,
                           aux_in=aux_in, aux_out=aux_out, act=act, out_f32=out_f32, splits=splits, config=config)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / iters
        tf = 2.0 * M * N * (K0 + K1) / ms / 1e9
        # cuBLAS reference timing
        for _ in range(3):
            torch.matmul(a0, b0.t())
        ev0.record()
        for _ in range(iters):
            torch.matmul(a0, b0.t())
        ev1.record()
        torch.cuda.synchronize()
        ms_ref = ev0.elapsed_time(ev1) / iters
        good = rel < tol and not (mx != mx)
        ok_all &= good
        print(f"{'OK ' if good else 'BAD'} M={M} N={N} K0={K0} K1={K1} [{flags}] max {mx:.4g} rel {rel:.3g} | "
              f"{ms*1e3:.1f} us {tf:.1f} TF/s (cublas plain {ms_ref*1e3:.1f} us)", flush=True)
    print("ALL_OK" if ok_all else "SOME_BAD", flush=True)
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
