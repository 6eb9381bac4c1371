"""Multi-GPU check of the expert-sharded grouped GEMM (run under torchrun, one rank per GPU; also works with 1 GPU).
Every rank verifies, over several dispatches with different routings (buffer reuse / epochs):
 (a) the peer-memory dispatch (EpBuffer, csrc/ep_dispatch.cuh) lands bit-identical rows, scale factors and psum layout
     to the NCCL all-to-all baseline (ep.dispatch_alltoall),
 (b) token_row maps every locally-owned token to its row,
 (c) the grouped GEMM output against an FP32 matmul of the dequantised received rows,
 (d) a CUDA graph of dispatch + GEMM replays to the same result."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if 'RANK' not in os.environ:
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT='29533', RANK='0', WORLD_SIZE='1')
    dist.init_process_group('nccl', device_id=dev)
    import deepgemm_b200 as dg
    from deepgemm_b200 import ep
    from deepgemm_b200.testing import calc_diff
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    g, n, k, t_local = 16, 512, 1024, 1000
    epr = g // world
    align = dg.get_mk_alignment_for_contiguous_layout()
    capacity = 2 * t_local * world + epr * align
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    w = torch.randn((epr, n, k), device=dev, dtype=torch.bfloat16, generator=gen)
    qs = [per_block_cast_to_fp8(w[i], True) for i in range(epr)]
    wq = (torch.stack([q[0] for q in qs]), torch.stack([q[1] for q in qs]))
    buf = ep.EpBuffer(g, capacity, k)
    torch.backends.cuda.matmul.allow_tf32 = False

    def check(xq, sf, ids, d, r, tag):
        torch.cuda.synchronize()
        ref = ep.dispatch_alltoall(xq, sf, ids, g, align)
        m_al = ref.a.shape[0]
        assert buf.num_rows() == m_al, (buf.num_rows(), m_al)
        assert not buf.overflowed()
        assert torch.equal(r.psum_layout, ref.psum_layout), tag
        valid = ref.grouped_layout >= 0
        assert torch.equal(r.a[:m_al].view(torch.uint8)[valid], ref.a.view(torch.uint8)[valid]), tag
        assert torch.equal(r.sfa[:m_al][valid], ref.sfa[valid]), tag
        # (b) my tokens that I own myself
        mine = (ids // epr) == rank
        rows = r.token_row[mine].long()
        assert torch.equal(r.a.view(torch.uint8)[rows], xq.view(torch.uint8)[mine]), tag
        # (c) GEMM on the received rows
        dense_sf = torch.empty((m_al, r.sfa.shape[1]), dtype=torch.int32, device=dev).copy_(r.sfa[:m_al])
        sfa = (dense_sf.view(torch.uint8).to(torch.int32) << 23).view(torch.float32)
        a_deq = r.a[:m_al].float() * sfa[:, :k // 128].repeat_interleave(128, 1)
        ok_rows, start = 0, 0
        for e in range(epr):
            end = int(r.psum_layout[e])
            if end > start:
                w_deq = wq[0][e].float() * wq[1][e].repeat_interleave(128, 0).repeat_interleave(128, 1)
                diff = calc_diff(d[start:end], a_deq[start:end] @ w_deq.t())
                assert diff < 1e-5, (rank, e, diff, tag)
                ok_rows += end - start
            start = (end + align - 1) // align * align
        assert ok_rows == ref.num_recv
        return ok_rows

    d = buf.output(n)                       # symmetric: combine() gathers from it
    # (capacity must also hold the top-8 round below: t_local / 4 tokens x 8 slots per rank)
    # every rank can rebuild every expert's weights (seeded by the owner's rank) to check what combine brings back
    w_all = []
    for o in range(world):
        go = torch.Generator(device=dev).manual_seed(1 + o)
        wo = torch.randn((epr, n, k), device=dev, dtype=torch.bfloat16, generator=go)
        w_all.append([per_block_cast_to_fp8(wo[i], True) for i in range(epr)])

    def check_combine(xq, sf, ids, r, tag):
        out = buf.combine(r.token_row, ids)
        torch.cuda.synchronize()
        sfx = (sf.contiguous().view(torch.uint8).to(torch.int32) << 23).view(torch.float32)
        x_deq = xq.float() * sfx[:, :k // 128].repeat_interleave(128, 1)
        for e in ids.unique().tolist():
            sel = ids == e
            wq_e, sw_e = w_all[e // epr][e % epr]
            w_deq = wq_e.float() * sw_e.repeat_interleave(128, 0).repeat_interleave(128, 1)
            diff = calc_diff(out[sel], x_deq[sel] @ w_deq.t())
            assert diff < 1e-5, (rank, e, diff, tag)
    total = 0
    for it in range(4):
        x = torch.randn((t_local, k), device=dev, dtype=torch.bfloat16, generator=gen)
        xq, sf = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
        hi = g if it != 2 else max(1, g // 4)          # iteration 2: skewed routing (few experts, some ranks idle)
        ids = torch.randint(0, hi, (t_local,), device=dev, generator=gen)
        if it == 3:
            ids = ids.to(torch.int32)
        d.fill_(float('nan'))
        _, r = ep.expert_sharded_grouped_gemm(xq, sf, ids, wq, buf, d, overlap=(it % 2 == 0))
        total += check(xq, sf, ids.long(), d, r, f'iter {it}')
        check_combine(xq, sf, ids.long() if ids.dtype != torch.int64 else ids, r, f'combine {it}')

    # (c') top-8 routing + weighted combine, bit-checked: out[t] = sum_j w[t,j] * D_owner[row(t,j)] with FP32 products and an
    #      FP32 running sum in slot order (the owners' D rows are fetched here through torch from an all-gather of D)
    topk = 8
    for it in range(2):
        x = torch.randn((t_local // 4, k), device=dev, dtype=torch.bfloat16, generator=gen)
        xq, sf = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
        tl = x.shape[0]
        ids8 = torch.stack([torch.randperm(g, device=dev, generator=gen)[:topk] for _ in range(tl)])
        if it == 1:
            ids8[::5, 3] = -1
        wts = torch.rand((tl, topk), device=dev, generator=gen)
        d.fill_(float('nan'))
        r = buf.dispatch(xq, sf, ids8)
        buf.grouped_gemm(wq, d, r.expected_m, overlap=False)
        out = buf.combine(r.token_row, ids8, weights=wts)
        torch.cuda.synchronize()
        assert not buf.overflowed()
        d_all = [torch.empty_like(d) for _ in range(world)]
        dist.all_gather(d_all, d.contiguous())
        d_all = torch.stack(d_all)                                       # [world, capacity, n]
        rows = r.token_row.view(tl, topk).long()
        routed = ids8 >= 0
        owner = (ids8.clamp(min=0) // epr).long()
        acc = torch.zeros((tl, n), dtype=torch.float32, device=dev)
        for j in range(topk):
            term = wts[:, j:j + 1] * d_all[owner[:, j], rows[:, j].clamp(min=0)].float()
            acc = torch.where(routed[:, j:j + 1], acc + term, acc)
        assert torch.equal(out, acc.to(torch.bfloat16)), f'weighted top-{topk} combine {it}'
        # and the rows are right: FP32 matmul of the dequantised token with the owner's expert weights
        sfx = (sf.contiguous().view(torch.uint8).to(torch.int32) << 23).view(torch.float32)
        x_deq = xq.float() * sfx[:, :k // 128].repeat_interleave(128, 1)
        for e in (0, g // 2, g - 1):
            sel = (ids8 == e).any(dim=1)
            if bool(sel.any()):
                wq_e, sw_e = w_all[e // epr][e % epr]
                w_deq = wq_e.float() * sw_e.repeat_interleave(128, 0).repeat_interleave(128, 1)
                slot = (ids8[sel] == e).float().argmax(dim=1)
                got = d_all[e // epr][rows[sel].gather(1, slot.unsqueeze(1)).squeeze(1)]
                assert calc_diff(got, x_deq[sel] @ w_deq.t()) < 1e-5, (rank, e)
        dist.barrier()                                                   # nobody re-dispatches while a peer still reads D
    total += 0

    # (d) CUDA graph: capture dispatch + GEMM once, replay with new inputs
    x = torch.randn((t_local, k), device=dev, dtype=torch.bfloat16, generator=gen)
    xq, sf = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
    ids = torch.randint(0, g, (t_local,), device=dev, generator=gen)
    sx, ssf, sids = torch.empty_like(xq), torch.empty_like(sf), torch.empty(t_local, dtype=torch.int64, device=dev)
    row = torch.empty(t_local, dtype=torch.int32, device=dev)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(2):                              # warm-up outside capture (same count on every rank)
            r = buf.dispatch(sx.copy_(xq), ssf.copy_(sf), sids.copy_(ids), row, wait=False)
            buf.grouped_gemm(wq, d, r.expected_m, overlap=True)
    torch.cuda.synchronize()
    dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        r = buf.dispatch(sx, ssf, sids, row, wait=False)
        buf.grouped_gemm(wq, d, r.expected_m, overlap=True)
    for it in range(2):
        x = torch.randn((t_local, k), device=dev, dtype=torch.bfloat16, generator=gen)
        xq, sf = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
        ids = torch.randint(0, g, (t_local,), device=dev, generator=gen)
        sx.copy_(xq), ssf.copy_(sf), sids.copy_(ids)
        d.fill_(float('nan'))
        graph.replay()
        total += check(xq, sf, ids, d, r, f'graph {it}')
    print(f'rank {rank}/{world}: ep check ok, {total} rows verified over 6 dispatches (2 under CUDA graph), top-8 weighted combine bit-exact', flush=True)
    del graph
    buf.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
