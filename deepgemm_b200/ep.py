"""Expert-sharded M-grouped FP8 GEMM (BASELINE config 5).

The grouped-contiguous path shards naturally by experts: rank r of P owns experts [r*G/P, (r+1)*G/P). Tokens start
evenly distributed over the ranks, each with one destination expert. One step =

  1. bucket the local tokens by destination expert (stable sort)                        -- device, no host sync
  2. exchange per-(rank, expert) counts:   all_to_all_single of a [P, G/P] int32 table  -- 4*G bytes per rank
  3. ONE payload all-to-all: every token travels as one row of K + 4*ceil(K/512) bytes = FP8 values followed by its
     packed UE8M0 scale factors (7168 + 56 B for DeepSeek-V3) over NCCL (NVLink 5 / NVSwitch: every peer at full
     bandwidth, so a flat variable-count all-to-all is the idiomatic dispatch; no topology-aware ring)
  4. lay the received rows out in the contiguous-grouped format the GEMM consumes (expert segments aligned to
     `get_mk_alignment_for_contiguous_layout()`, psum layout = end row per expert) and transpose the scale factors to
     the MN-major wire format
  5. local `m_grouped_fp8_gemm_nt_contiguous` on this rank's experts.

This mirrors how the reference's grouped GEMM sits inside expert parallelism in its own baseline
(tests/test_mega_moe.py:148-205: DeepEP dispatch -> m_grouped_fp8_fp4_gemm_nt_contiguous(use_psum_layout=True) ->
combine). The reverse all-to-all + weighted reduce ("combine") is the next row.

Steps 1-4 are torch plumbing and run on any backend (the world_size-2 `gloo` CPU test covers them); step 5 needs CUDA.
"""
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def _ceil_div(a: int, b: int) -> int:
    return -(-a // b)


@dataclass
class DispatchResult:
    a: torch.Tensor               # [m_aligned, K] e4m3 (uint8 view on CPU backends), expert segments, zero padding rows
    sfa: torch.Tensor             # int32 [m_aligned, ceil(K/512)] MN-major packed UE8M0 (strides (1, m_aligned))
    psum_layout: torch.Tensor     # int32 [experts_per_rank]: end row (unaligned) of each local expert segment
    grouped_layout: torch.Tensor  # int32 [m_aligned]: local expert id per row, -1 on padding rows
    recv_counts: torch.Tensor     # int64 [P, experts_per_rank] tokens received from each rank for each local expert
    src_order: torch.Tensor       # int64 [num local tokens]: permutation that sorted the local tokens by expert
    num_recv: int


def pack_rows(x_fp8: torch.Tensor, sf_packed: torch.Tensor) -> torch.Tensor:
    """One wire row per token: K FP8 bytes followed by the token's packed UE8M0 words (K-major int32)."""
    t, k = x_fp8.shape
    assert sf_packed.dtype == torch.int32 and sf_packed.shape == (t, _ceil_div(k, 512))
    return torch.cat([x_fp8.contiguous().view(torch.uint8), sf_packed.contiguous().view(torch.uint8).view(t, -1)], dim=1)


def dispatch(x_fp8: torch.Tensor, sf_packed: torch.Tensor, expert_ids: torch.Tensor, num_experts: int, alignment: int,
             group: Optional[dist.ProcessGroup] = None) -> DispatchResult:
    """Steps 1-4. `x_fp8` [T,K] e4m3 (or uint8), `sf_packed` [T, ceil(K/512)] int32 (K-major, as
    per_token_cast_to_fp8(..., use_packed_ue8m0=True) returns), `expert_ids` [T] int64 in [0, num_experts)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    assert num_experts % world == 0
    epr = num_experts // world
    t, k = x_fp8.shape
    kp = sf_packed.shape[1]
    dev = x_fp8.device

    # 1. bucket by destination expert (=> also by destination rank)
    order = torch.argsort(expert_ids, stable=True)
    rows = pack_rows(x_fp8, sf_packed)[order]
    counts = torch.bincount(expert_ids, minlength=num_experts).view(world, epr)

    # 2. counts: what every rank sends me for each of my experts
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    send_split = counts.sum(dim=1).tolist()                 # host sync: split sizes of the payload exchange
    recv_split = recv_counts.sum(dim=1).tolist()
    num_recv = int(sum(recv_split))

    # 3. the payload: ONE all-to-all of (K + 4*kp)-byte rows
    recv_rows = torch.empty((num_recv, rows.shape[1]), dtype=torch.uint8, device=dev)
    dist.all_to_all_single(recv_rows, rows, output_split_sizes=recv_split, input_split_sizes=send_split, group=group)

    # 4. contiguous-grouped layout. Received rows are ordered (source rank, expert); row j of source s for expert e
    #    goes to  seg_start[e] + (rows for e from sources < s) + j.
    per_expert = recv_counts.sum(dim=0)                                        # [epr]
    aligned = (per_expert + alignment - 1) // alignment * alignment
    seg_start = torch.cumsum(aligned, 0) - aligned
    m_aligned = int(aligned.sum())
    before = torch.cumsum(recv_counts, 0) - recv_counts                        # [P, epr] rows for e from earlier sources
    block_start = (seg_start.unsqueeze(0) + before).reshape(-1)                # destination of each (s, e) block
    block_len = recv_counts.reshape(-1)
    src_block_start = torch.cumsum(block_len, 0) - block_len
    block_of_row = torch.repeat_interleave(torch.arange(world * epr, device=dev), block_len, output_size=num_recv)
    dest = block_start[block_of_row] + (torch.arange(num_recv, device=dev) - src_block_start[block_of_row])

    a = torch.zeros((m_aligned, k), dtype=torch.uint8, device=dev)
    a[dest] = recv_rows[:, :k]
    sf_rows = recv_rows[:, k:].contiguous().view(torch.int32).view(num_recv, kp)
    sfa_t = torch.zeros((kp, m_aligned), dtype=torch.int32, device=dev)       # MN-major storage
    sfa_t[:, dest] = sf_rows.t()
    grouped_layout = torch.full((m_aligned,), -1, dtype=torch.int32, device=dev)
    grouped_layout[dest] = (block_of_row % epr).to(torch.int32)
    psum = (seg_start + per_expert).to(torch.int32)
    if x_fp8.dtype == torch.float8_e4m3fn:
        a = a.view(torch.float8_e4m3fn)
    return DispatchResult(a=a, sfa=sfa_t.t(), psum_layout=psum, grouped_layout=grouped_layout, recv_counts=recv_counts,
                          src_order=order, num_recv=num_recv)


def dispatch_local(x_fp8: torch.Tensor, sf_packed: torch.Tensor, expert_ids: torch.Tensor, num_experts: int,
                   alignment: int) -> DispatchResult:
    """World size 1: the same bucketing + layout steps without any communication (all experts are local)."""
    t, k = x_fp8.shape
    kp = sf_packed.shape[1]
    dev = x_fp8.device
    order = torch.argsort(expert_ids, stable=True)
    counts = torch.bincount(expert_ids, minlength=num_experts)
    aligned = (counts + alignment - 1) // alignment * alignment
    seg_start = torch.cumsum(aligned, 0) - aligned
    src_start = torch.cumsum(counts, 0) - counts
    m_aligned = int(aligned.sum())
    e_sorted = expert_ids[order]
    dest = seg_start[e_sorted] + (torch.arange(t, device=dev) - src_start[e_sorted])
    a = torch.zeros((m_aligned, k), dtype=torch.uint8, device=dev)
    a[dest] = x_fp8.contiguous().view(torch.uint8)[order]
    sfa_t = torch.zeros((kp, m_aligned), dtype=torch.int32, device=dev)
    sfa_t[:, dest] = sf_packed[order].t()
    layout = torch.full((m_aligned,), -1, dtype=torch.int32, device=dev)
    layout[dest] = e_sorted.to(torch.int32)
    if x_fp8.dtype == torch.float8_e4m3fn:
        a = a.view(torch.float8_e4m3fn)
    return DispatchResult(a=a, sfa=sfa_t.t(), psum_layout=(seg_start + counts).to(torch.int32), grouped_layout=layout,
                          recv_counts=counts.view(1, -1), src_order=order, num_recv=t)


def expert_sharded_grouped_gemm(x_fp8: torch.Tensor, sf_packed: torch.Tensor, expert_ids: torch.Tensor,
                                w_local: Tuple[torch.Tensor, torch.Tensor], num_experts: int,
                                group: Optional[dist.ProcessGroup] = None,
                                use_psum_layout: bool = True) -> Tuple[torch.Tensor, DispatchResult]:
    """Dispatch + local grouped GEMM. `w_local` = (B [G/P, N, K] e4m3, SFB) for THIS rank's experts (SFB FP32
    [G/P, N/128, K/128] or pre-packed int32). Returns (D [m_aligned, N] bf16 on the expert rank, dispatch record)."""
    from . import gemm, runtime
    alignment = runtime.get_mk_alignment_for_contiguous_layout()
    r = dispatch(x_fp8, sf_packed, expert_ids, num_experts, alignment, group)
    n = w_local[0].shape[1]
    d = torch.empty((r.a.shape[0], n), dtype=torch.bfloat16, device=x_fp8.device)
    if r.a.shape[0] > 0:
        layout = r.psum_layout if use_psum_layout else r.grouped_layout
        gemm.m_grouped_fp8_gemm_nt_contiguous((r.a, r.sfa), w_local, d, layout, use_psum_layout=use_psum_layout)
    return d, r
