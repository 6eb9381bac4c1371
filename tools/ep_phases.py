"""Where the time of the expert-parallel dispatch / combine goes on this rank (run under torchrun, one rank per GPU): the bench's
config-5 problem (256 experts, 32768 tokens in total, K=7168, N=4096), phase durations from the kernels' own globaltimer stamps.
Development tool."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if 'RANK' not in os.environ:
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT='29537', RANK='0', WORLD_SIZE='1')
    dist.init_process_group('nccl', device_id=dev)
    import deepgemm_b200 as dg
    from deepgemm_b200 import ep
    g, n, k, tokens_total = 256, 4096, 7168, 32768
    b, sfb_p, xq, sf_packed, ids, capacity = bench.ep_problem(rank, world, dev, dg, g, n, k, tokens_total)
    buf = ep.EpBuffer(g, capacity, k)
    d = buf.output(n)
    t_local = xq.shape[0]
    token_row = torch.empty(t_local, dtype=torch.int32, device=dev)
    out_tokens = torch.empty((t_local, n), device=dev, dtype=torch.bfloat16)
    names = ['rank+barrier', 'publish counts', 'wait counts', 'layout+barrier', 'scatter (CTA 0)', 'fence+last CTA', 'wait peers']

    def step():
        r = buf.dispatch(xq, sf_packed, ids, token_row)
        buf.grouped_gemm((b, sfb_p), d, r.expected_m, overlap=False)
        buf.combine(token_row, ids, out_tokens)

    rows = {'synced': [], 'free_running': []}
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    for mode in ('synced', 'free_running'):
        for it in range(8):
            if mode == 'synced':
                torch.cuda.synchronize()
                dist.barrier()
                step()
            else:
                for _ in range(6):
                    step()
            torch.cuda.synchronize()
            ts = buf.debug_timestamps()
            ph = [(ts[i + 1] - ts[i]) / 1e3 for i in range(7)]
            rows[mode].append({'dispatch_us': (ts[7] - ts[0]) / 1e3, 'phases_us': ph, 'dispatch_end_to_combine_start_us': (ts[8] - ts[7]) / 1e3,
                               'combine_wait_us': (ts[9] - ts[8]) / 1e3, 'combine_gather_cta0_us': (ts[10] - ts[9]) / 1e3})
    med = lambda xs: sorted(xs)[len(xs) // 2]   # noqa: E731
    res = {'rank': rank, 'world': world}
    for mode, rs in rows.items():
        res[mode] = {'dispatch_us': round(med([r['dispatch_us'] for r in rs]), 1),
                     'phases_us': {nm: round(med([r['phases_us'][i] for r in rs]), 1) for i, nm in enumerate(names)},
                     'gemm_window_us': round(med([r['dispatch_end_to_combine_start_us'] for r in rs]), 1),
                     'combine_wait_us': round(med([r['combine_wait_us'] for r in rs]), 1),
                     'combine_gather_cta0_us': round(med([r['combine_gather_cta0_us'] for r in rs]), 1)}
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        for r in gathered:
            print(json.dumps(r), flush=True)
    buf.close()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
