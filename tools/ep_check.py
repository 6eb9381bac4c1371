"""Multi-GPU check of the expert-sharded grouped GEMM (run under torchrun, one rank per GPU).
Every rank verifies (a) conservation: the rows it received are exactly the rows the others sent for its experts
(global byte checksum), (b) its grouped GEMM output against an FP32 matmul of the dequantised received rows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    import deepgemm_b200 as dg  # noqa: F401
    from deepgemm_b200 import ep
    from deepgemm_b200.testing import calc_diff
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    g, n, k, t_local = 16, 512, 1024, 1000
    epr = g // world
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    w = torch.randn((epr, n, k), device=dev, dtype=torch.bfloat16, generator=gen)
    qs = [per_block_cast_to_fp8(w[i], True) for i in range(epr)]
    wq = (torch.stack([q[0] for q in qs]), torch.stack([q[1] for q in qs]))
    x = torch.randn((t_local, k), device=dev, dtype=torch.bfloat16, generator=gen)
    xq, sf = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
    ids = torch.randint(0, g, (t_local,), device=dev, generator=gen)
    d, r = ep.expert_sharded_grouped_gemm(xq, sf, ids, wq, g)
    torch.cuda.synchronize()
    # (a) conservation of tokens and bytes
    sent = torch.tensor([float(t_local), float(xq.view(torch.uint8).double().sum())], device=dev, dtype=torch.float64)
    recv = torch.tensor([float(r.num_recv), float(r.a.view(torch.uint8).double().sum())], device=dev, dtype=torch.float64)
    dist.all_reduce(sent), dist.all_reduce(recv)
    assert torch.equal(sent, recv), (sent, recv)
    # (b) GEMM on the received rows
    torch.backends.cuda.matmul.allow_tf32 = False
    dense_sf = torch.empty(r.sfa.shape, dtype=torch.int32, device=dev).copy_(r.sfa)
    sfa = (dense_sf.view(torch.uint8).to(torch.int32) << 23).view(torch.float32)      # [m, 4*kp] per-128 scales
    a_deq = r.a.float() * sfa[:, :k // 128].repeat_interleave(128, 1)
    ok_rows, start = 0, 0
    for e in range(epr):
        end = int(r.psum_layout[e])
        if end > start:
            w_deq = wq[0][e].float() * wq[1][e].repeat_interleave(128, 0).repeat_interleave(128, 1)
            want = a_deq[start:end] @ w_deq.t()
            diff = calc_diff(d[start:end], want)
            assert diff < 1e-5, (rank, e, diff)
            ok_rows += end - start
        start = (end + 127) // 128 * 128
    assert ok_rows == r.num_recv
    print(f'rank {rank}/{world}: ep check ok, received {r.num_recv} rows', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
