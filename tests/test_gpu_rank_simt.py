"""CUDA-core rank-side kernel (csrc/rank_simt.cu, `b200_rank_gemm`): Z = X A_pack^T and T = dY B_pack for live ranks <= 16,
against an fp32 reference and against the tensor-core skinny GEMM it replaces, at the FLUX and SDXL shapes, ragged M / K,
per-sample multipliers; dead columns must come out as exact zeros over a NaN-filled buffer."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


@pytest.mark.parametrize("M,K,r,trans", [
    (4608, 3072, 16, False), (4608, 12288, 16, True), (4608, 3072, 16, True),   # FLUX r16
    (2048, 1280, 8, False), (2048, 5120, 8, True), (8192, 640, 8, False),       # SDXL r8 (K = 640: 2.5 chunks)
    (154, 2048, 8, False), (154, 1280, 8, True),                                  # SDXL cross-attention k / v of the 2 x 77 text rows
    (1001, 1288, 4, False), (1001, 1288, 4, True), (37, 264, 5, False), (37, 264, 5, True), (4608, 3072, 12, False),
])
def test_rank_gemm_matches_fp32_and_tensor_core_path(M, K, r, trans):
    from ai_toolkit_b200 import cabi, ops
    g = torch.Generator().manual_seed(M + K + r)
    x = torch.randn(M, K, generator=g).bfloat16().to(DEV)
    if trans:
        w = torch.zeros(K, 64)
        w[:, :r] = torch.randn(K, r, generator=g) * 0.05
    else:
        w = torch.zeros(64, K)
        w[:r] = torch.randn(r, K, generator=g) * 0.05
    w = w.bfloat16().to(DEV)
    alpha = 0.75
    out = torch.full((M, 64), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.rank_gemm(x, w, out, r, trans_w=trans, alpha=alpha)
    ref = alpha * (x.float() @ (w.float() if trans else w.float().t()))
    assert torch.isfinite(out.float()).all()
    assert bool((out[:, r:] == 0).all())
    assert _rel(out[:, :r], ref[:, :r]) < 3e-3          # one bf16 rounding of an fp32 sum
    tc = torch.empty((M, 64), device=DEV, dtype=torch.bfloat16)
    cabi.gemm_bf16(x, w, tc, trans_b=trans, alpha=alpha)
    # same arithmetic, different summation order: differences are single bf16 ulps on a few elements
    assert _rel(out, tc) < 2e-3
    assert _rel(out[:, :r], ref[:, :r]) < 1.5 * _rel(tc[:, :r], ref[:, :r]) + 1e-4


def test_rank_gemm_per_sample_multiplier_and_row_stride():
    from ai_toolkit_b200 import ops
    g = torch.Generator().manual_seed(3)
    M, K, r, rps = 1024, 1024, 16, 256
    big = torch.randn(M, K + 64, generator=g).bfloat16().to(DEV)
    x = big[:, 64:]                                      # a column slice: row stride K + 64, 16-byte aligned start
    w = torch.zeros(64, K)
    w[:r] = torch.randn(r, K, generator=g) * 0.05
    w = w.bfloat16().to(DEV)
    ra = torch.tensor([1.0, -0.5, 0.0, 2.0], device=DEV)
    out = torch.full((M, 64), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.rank_gemm(x, w, out, r, alpha=0.5, row_alpha=ra, rows_per_sample=rps)
    ref = 0.5 * ra.repeat_interleave(rps)[:, None] * (x.float() @ w.float().t())
    assert _rel(out, ref) < 3e-3 and bool((out[2 * rps:3 * rps] == 0).all())


def test_rank_gemm_rejects_large_rank():
    from ai_toolkit_b200 import cabi, ops
    x = torch.zeros(64, 256, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(64, 256, device=DEV, dtype=torch.bfloat16)
    out = torch.zeros(64, 64, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(cabi.B200Error):
        ops.rank_gemm(x, w, out, 17)
