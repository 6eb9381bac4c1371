"""Where do the microseconds between 'kernel time by the profiler' and 'CUDA events around one launch' go? (development tool)
For ours (a few knob settings) and the reference: event-timed (L2 flushed) and kineto-timed duration of the same launch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools.bringup import import_reference, make_inputs  # noqa: E402
import deepgemm_b200 as dg  # noqa: E402
from deepgemm_b200 import _lib  # noqa: E402
from deepgemm_b200.testing import bench_events, bench_kineto  # noqa: E402

ref = import_reference()
KEYS = ('DGB200_STAGES', 'DGB200_TMA_STORE', 'DGB200_BLOCK_M', 'DGB200_CSPLIT')
for (m, n, k) in [(4096, 4096, 7168), (64, 4096, 7168), (512, 4096, 7168), (128, 128, 128)]:
    a, b, qa, qb = make_inputs(m, n, k)
    sfa = dg.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
    sfb = dg.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
    d = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    fr = lambda: ref.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d)  # noqa: E731
    fo = lambda: dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d)  # noqa: E731
    rows = []
    fr()
    ev = sorted(bench_events(fr, 3, 20))
    rows.append(('ref', round(bench_kineto(fr, 'gemm_', num_tests=10) * 1e6, 2), round(ev[10] * 1e6, 2), round(ev[0] * 1e6, 2)))
    for env in ({}, {'DGB200_STAGES': '3'}, {'DGB200_TMA_STORE': '0'}, {'DGB200_CSPLIT': '0'}):
        for kk in KEYS:
            os.environ.pop(kk, None)
        os.environ.update(env)
        fo()
        cfg = _lib.last_config()
        ev = sorted(bench_events(fo, 3, 20))
        rows.append((json.dumps(env), round(bench_kineto(fo, 'fp8_gemm_kernel', num_tests=10) * 1e6, 2), round(ev[10] * 1e6, 2), round(ev[0] * 1e6, 2),
                     cfg['smem_bytes'], cfg['num_stages'], cfg['cluster']))
    for kk in KEYS:
        os.environ.pop(kk, None)
    # an empty-ish kernel for scale: torch's smallest elementwise op
    z = torch.zeros(32, device='cuda')
    ev = sorted(bench_events(lambda: z.add_(1), 3, 20))
    rows.append(('torch add_ (32 elements)', None, round(ev[10] * 1e6, 2), round(ev[0] * 1e6, 2)))
    print(f'== {m}x{n}x{k}   (name, kineto us, event median us, event min us, ...)', flush=True)
    for r in rows:
        print('  ', *r, flush=True)
