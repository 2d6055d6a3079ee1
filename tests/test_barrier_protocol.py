"""The mbarrier / tcgen05 pipelines of the attention kernels under random, skewed interleavings (tools/mbar_model.py).

Checked at every parity wait: the waiter is at most one phase behind (else the parity test lies), the parity expression
written in the kernel matches the phase it means, and the data hazards the barriers protect (S / P / X buffer reuse, O
rescale vs the P V MMA, TMA stage reuse).  Negative controls show the checker finds the two bug classes met so far."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
import mbar_model as m  # noqa: E402


@pytest.mark.parametrize("n_kv,n_soft,rescale", [(8, 3, 0.3), (5, 2, 0.9), (1, 2, 0.5), (2, 2, 0.5), (3, 4, 0.0)])
def test_forward_lazy_pv_protocol(n_kv, n_soft, rescale):
    """attention_r2.cu `attn_fwd_nt_kernel`: per-buffer p_full / pv_done, lazy waits."""
    for seed in range(400):
        m.run_fwd(seed, n_kv=n_kv, n_soft=n_soft, lazy=True, double_p_full=True, rescale_prob=rescale)
    for seed in range(400):  # the row-max exchange synchronises only the warps sharing a lane quarter (groups of 2)
        m.run_fwd(seed, n_kv=n_kv, n_soft=4, lazy=True, double_p_full=True, rescale_prob=rescale, sync_group=2)


def test_forward_round1_protocol():
    """attention.cu `attn_fwd_kernel` (validated): single p_full / pv_done, softmax(j) always waits for P V(j-1)."""
    for seed in range(400):
        m.run_fwd(seed, n_kv=7, n_soft=3, lazy=False, double_p_full=False, rescale_prob=0.0)


def test_forward_lazy_with_single_p_full_is_caught():
    """Why p_full had to become per-buffer: without the P V(j-1) wait the softmax can finish tile j+1 before a delayed
    MMA warp has tested phase j — it is then two phases behind and its parity test blocks for ever."""
    hits = 0
    for seed in range(600):
        try:
            m.run_fwd(seed, lazy=True, double_p_full=False)
        except m.ProtocolError as e:
            assert "p_full" in str(e)
            hits += 1
    assert hits > 0


@pytest.mark.parametrize("n_t,group_size,stages", [(9, 2, 3), (6, 3, 3), (1, 2, 3), (2, 2, 3), (72, 1, 3)])
def test_backward_two_group_protocol(n_t, group_size, stages):
    """attention.cu `attn_bwd_kernel` and attention_r2.cu `attn_bwd_r2_kernel`: group g <-> TMEM buffer g <-> pb_full[g]."""
    for seed in range(60 if n_t > 20 else 400):
        m.run_bwd(seed, n_t=n_t, group_size=group_size, stages=stages)


def test_backward_shared_arrival_barrier_is_caught():
    for seed in range(20):
        with pytest.raises(m.ProtocolError):
            m.run_bwd(seed, shared_pb_full=True)


@pytest.mark.parametrize("n_t,group_size,stages", [(9, 2, 4), (6, 3, 3), (1, 2, 4), (2, 2, 4), (3, 2, 4), (72, 1, 4)])
def test_backward_dq_early_issue_protocol(n_t, group_size, stages):
    """attention_r2.cu `attn_bwd_q2_kernel`: A(i+2) is issued on x_free (the group has loaded S / dP), dS lives in D[g],
    softmax(i) waits d_free = B(i-2) before rewriting D[g]."""
    for seed in range(60 if n_t > 20 else 400):
        m.run_bwd_q2(seed, n_t=n_t, group_size=group_size, stages=stages)


def test_backward_dq_without_d_free_wait_is_caught():
    hits = 0
    for seed in range(200):
        try:
            m.run_bwd_q2(seed, wait_d_free=False)
        except m.ProtocolError as e:
            assert "overwrites dS" in str(e)
            hits += 1
    assert hits > 0


@pytest.mark.parametrize("n_tiles,kb,stages,zero", [(5, 4, 3, ()), (7, 1, 6, ()), (3, 48, 6, ()), (6, 3, 2, (1, 4)), (1, 2, 6, ())])
def test_gemm_pipeline_protocol(n_tiles, kb, stages, zero):
    """gemm_tcgen05.cu: TMA ring (full / empty), two TMEM accumulators (tfull / tempty), including split-K slices that have
    no k-block at all (the handshake must still run, and the epilogue must see a zero partial)."""
    for seed in range(60 if kb > 20 else 300):
        m.run_gemm(seed, n_tiles=n_tiles, kb_per_tile=kb, stages=stages, n_epi=3, tiles_with_zero_k=zero)
