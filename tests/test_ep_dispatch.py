"""CPU, world_size 2 over gloo: the expert-parallel dispatch (deepgemm_b200/ep.py steps 1-4) delivers every token, with
its scale factors, to the rank that owns its expert, laid out in the contiguous-grouped format the GEMM consumes."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _make_rank_inputs(rank, t, k, num_experts):
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randint(0, 255, (t, k), dtype=torch.uint8, generator=g)
    sf = torch.randint(0, 2 ** 31 - 1, (t, (k + 511) // 512), dtype=torch.int32, generator=g)
    ids = torch.randint(0, num_experts, (t,), generator=g)
    return x, sf, ids


def _worker(rank, world, port, t, k, num_experts, alignment, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from deepgemm_b200 import ep
        x, sf, ids = _make_rank_inputs(rank, t, k, num_experts)
        r = ep.dispatch_alltoall(x, sf, ids, num_experts, alignment)
        res = {k_: getattr(r, k_) for k_ in ('a', 'psum_layout', 'grouped_layout', 'recv_counts', 'num_recv')}
        res['sfa'] = torch.empty(r.sfa.shape, dtype=torch.int32).copy_(r.sfa)
        res['sfa_stride'] = tuple(r.sfa.stride())
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_dispatch_world2_gloo():
    world, t, k, num_experts, alignment = 2, 37, 1024, 6, 16
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, t, k, num_experts, alignment, out), nprocs=world, join=True)
        out = dict(out)
    inputs = [_make_rank_inputs(r, t, k, num_experts) for r in range(world)]
    epr = num_experts // world
    for rank in range(world):
        res = out[rank]
        a, sfa, psum, layout = res['a'], res['sfa'], res['psum_layout'], res['grouped_layout']
        assert res['sfa_stride'] == (1, a.shape[0])                      # MN-major wire format
        start = 0
        total = 0
        for le in range(epr):
            e = rank * epr + le
            # expected rows: tokens of expert e from source rank 0 (in original order), then from source rank 1
            exp_x = torch.cat([inp[0][inp[2] == e] for inp in inputs])
            exp_sf = torch.cat([inp[1][inp[2] == e] for inp in inputs])
            cnt = exp_x.shape[0]
            end = int(psum[le])
            assert end - start == cnt
            assert torch.equal(a[start:end], exp_x)
            assert torch.equal(sfa[start:end], exp_sf)
            assert bool((layout[start:end] == le).all())
            aligned_end = start + (cnt + alignment - 1) // alignment * alignment
            assert bool((layout[end:aligned_end] == -1).all())
            assert bool((a[end:aligned_end] == 0).all())                 # zero padding rows
            start = aligned_end
            total += cnt
        assert start == a.shape[0] and total == res['num_recv']
        assert int(res['recv_counts'].sum()) == total
    assert sum(out[r]['num_recv'] for r in range(world)) == world * t  # every token arrived exactly once
