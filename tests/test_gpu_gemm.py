"""tcgen05 GEMM (b200_gemm_bf16) vs an fp32 torch reference of the same op, through the C ABI.
Tolerance: bf16 outputs within 6e-3 relative Frobenius error (one bf16 rounding of an fp32 accumulation,
bf16 ulp = 2^-8); fp32 outputs within 2e-5."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _ref_epilogue(acc, bias=None, act=0, gate=None, rps=0, res=None, aux_in=None):
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


CASES = [
    # M, N, K0, K1, flags, config
    (128, 256, 64, 0, "", 1), (256, 512, 3072, 0, "", 1), (200, 264, 328, 64, "bias", 1),
    (512, 3072, 3072, 64, "bias,gelu,auxout", 2), (512, 3072, 3072, 64, "bias,gate,res", 2),
    (512, 3072, 3072, 0, "auxin", 2), (1160, 1544, 1032, 64, "bias", 2), (384, 768, 512, 64, "bias", 3),
    (200, 48, 328, 0, "", 4), (4608, 64, 3072, 0, "alpha", 4),
    (4608, 3072, 3072, 64, "bias", 2), (4608, 3072, 15360, 64, "bias,gate,res", 2),
    # 128 x 160 / 128 x 192 tiles (round 2: SDXL's 2048-token levels) explicitly, ragged edges, and through AUTO (config 0)
    (2048, 1280, 1280, 64, "bias,res", 6), (2048, 1280, 5120, 64, "bias,res", 7), (200, 328, 264, 64, "bias", 6),
    (1160, 1544, 1032, 64, "bias,gelu,auxout", 7), (2048, 1280, 1280, 64, "bias,res", 0), (2048, 10240, 1280, 64, "bias", 0),
    (154, 1280, 2048, 64, "", 0),
]


@pytest.mark.parametrize("M,N,K0,K1,flags,config", CASES)
def test_forward_nt(M, N, K0, K1, flags, config):
    from ai_toolkit_b200 import cabi
    torch.manual_seed(M + N + K0)
    dev = torch.device("cuda:0")
    a0 = (torch.randn(M, K0, device=dev) * 0.5).bfloat16()
    b0 = (torch.randn(N, K0, device=dev) * 0.05).bfloat16()
    a1 = (torch.randn(M, K1, device=dev) * 0.5).bfloat16() if K1 else None
    b1 = (torch.randn(N, K1, device=dev) * 0.05).bfloat16() if K1 else None
    bias = torch.randn(N, device=dev).bfloat16() if "bias" in flags else None
    act = 1 if "gelu" in flags else 0
    rps, gate, res, aux_in, aux_out = 0, None, None, None, None
    alpha = 0.37 if "alpha" in flags else 1.0
    if "gate" in flags:
        rps = max(1, M // 2)
        gate = torch.randn((M + rps - 1) // rps, N, device=dev).bfloat16()
    if "res" in flags:
        res = torch.randn(M, N, device=dev).bfloat16()
    if "auxin" in flags:
        aux_in = torch.randn(M, N, device=dev).bfloat16()
    if "auxout" in flags:
        aux_out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    cabi.gemm_bf16(a0, b0, out, a1=a1, b1=b1, bias=bias, res=res, gate=gate, rows_per_sample=rps, aux_in=aux_in,
                   aux_out=aux_out, act=act, alpha=alpha, config=config)
    torch.cuda.synchronize()
    acc = a0.float() @ b0.float().t()
    if a1 is not None:
        acc = acc + a1.float() @ b1.float().t()
    want, pre = _ref_epilogue(acc * alpha, bias, act, gate, rps, res, aux_in)
    assert not torch.isnan(out.float()).any()
    assert _err(out, want) < 6e-3
    if aux_out is not None:
        assert _err(aux_out, pre) < 6e-3


@pytest.mark.parametrize("M,N,K0,K1,config", [(256, 256, 128, 0, 1), (200, 328, 264, 64, 1), (512, 3072, 3072, 64, 2),
                                              (4608, 3072, 12288, 64, 2), (4608, 64, 3072, 0, 4), (384, 768, 512, 64, 3),
                                              (2048, 1280, 1280, 64, 7), (200, 328, 264, 64, 7), (2048, 1280, 10240, 64, 0),
                                              (2048, 5120, 1280, 64, 0)])
def test_dgrad_trans_b(M, N, K0, K1, config):
    """dX = dY W: B operands stored [K, N] (N contiguous) and consumed MN-major, no transposed copy."""
    from ai_toolkit_b200 import cabi
    torch.manual_seed(1 + M + N)
    dev = torch.device("cuda:0")
    a0 = (torch.randn(M, K0, device=dev) * 0.5).bfloat16()
    b0 = (torch.randn(K0, N, device=dev) * 0.05).bfloat16()
    a1 = (torch.randn(M, K1, device=dev) * 0.5).bfloat16() if K1 else None
    b1 = (torch.randn(K1, N, device=dev) * 0.05).bfloat16() if K1 else None
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16)
    cabi.gemm_bf16(a0, b0, out, a1=a1, b1=b1, trans_b=True, config=config)
    torch.cuda.synchronize()
    acc = a0.float() @ b0.float()
    if a1 is not None:
        acc = acc + a1.float() @ b1.float()
    assert _err(out, acc.bfloat16()) < 6e-3


@pytest.mark.parametrize("P,T,r,splits,trans", [(128, 64, 16, 1, False), (3072, 4608, 16, 6, False),
                                                (3072, 4608, 16, 6, True), (12288, 512, 4, 2, True),
                                                (328, 200, 64, 3, False), (3072, 1, 16, 1, False)])
def test_wgrad_trans_ab_atomic(P, T, r, splits, trans):
    """out[P, r] += alpha * L[T, P]^T R[T, 64]: both operands MN-major, fp32 atomic accumulate, optional
    transposed store — the dB = dY^T Z and dA^T = X^T T contractions."""
    from ai_toolkit_b200 import cabi
    torch.manual_seed(P + T)
    dev = torch.device("cuda:0")
    L = (torch.randn(T, P, device=dev) * 0.5).bfloat16()
    R = torch.zeros(T, 64, device=dev, dtype=torch.bfloat16)
    R[:, :r] = (torch.randn(T, r, device=dev) * 0.5).bfloat16()
    init = torch.randn(P, r, device=dev)
    out = init.t().contiguous() if trans else init.clone()
    cabi.gemm_bf16(L, R, out, trans_a=True, trans_b=True, alpha=0.5, f32_mode=2, f32_trans=trans, n_store=r, splits=splits,
                   config=4)
    torch.cuda.synchronize()
    want = init + 0.5 * (L.float().t() @ R.float()[:, :r])
    got = out.t() if trans else out
    assert _err(got, want) < 2e-5


def test_split_k_partials():
    from ai_toolkit_b200 import cabi
    dev = torch.device("cuda:0")
    M, N, K = 4608, 64, 15360
    a0 = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    b0 = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    out = torch.full((8, M, N), float("nan"), device=dev)
    cabi.gemm_bf16(a0, b0, out, f32_mode=1, splits=8, config=4)
    torch.cuda.synchronize()
    assert _err(out.sum(0), a0.float() @ b0.float().t()) < 2e-5


def test_bad_arguments_fail_loudly():
    from ai_toolkit_b200 import cabi
    dev = torch.device("cuda:0")
    a = torch.zeros(128, 60, device=dev, dtype=torch.bfloat16)  # K not a multiple of 8
    b = torch.zeros(128, 60, device=dev, dtype=torch.bfloat16)
    out = torch.zeros(128, 128, device=dev, dtype=torch.bfloat16)
    with pytest.raises(cabi.B200Error):
        cabi.gemm_bf16(a, b, out)


@pytest.mark.parametrize("M,N,K,trans_b", [(4608, 64, 3072, False), (4608, 64, 3072, True), (512, 64, 3072, False),
                                           (4096, 64, 15360, False), (200, 48, 328, True), (4608, 64, 12288, True),
                                           (130, 16, 64, False)])
def test_skinny_cluster_splitk(M, N, K, trans_b):
    """Rank-side GEMM with the contraction split over a CTA cluster and reduced through DSMEM (config 5)."""
    from ai_toolkit_b200 import cabi
    torch.manual_seed(M + K)
    dev = torch.device("cuda:0")
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    b = (torch.randn(K, N, device=dev) * 0.05).bfloat16() if trans_b else (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    ra = torch.rand(2, device=dev) + 0.5
    rps = (M + 1) // 2
    out = torch.full((M, 64), float("nan"), device=dev, dtype=torch.bfloat16)
    cabi.gemm_bf16(a, b, out, trans_b=trans_b, alpha=0.7, row_alpha=ra, rows_per_sample=rps, N=N, config=cabi.GEMM_SKINNY_CLUSTER)
    torch.cuda.synchronize()
    want = 0.7 * (a.float() @ (b.float() if trans_b else b.float().t()))
    want = want * ra[torch.arange(M, device=dev) // rps][:, None]
    assert _err(out[:, :N], want.bfloat16()) < 6e-3
    if N < 64:
        assert torch.isnan(out[:, N:].float()).all()
