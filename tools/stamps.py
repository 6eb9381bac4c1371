"""Dump CTA-0 life-cycle timestamps (clock64) of one GEMM launch for a few shapes. Development tool."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools.bringup import make_inputs  # noqa: E402

import deepgemm_b200 as dg  # noqa: E402
from deepgemm_b200 import _lib  # noqa: E402

NAMES = ['entry', 'setup_done', 'first_tma', 'first_data', 'producer_start', 'last_mma', 'acc_ready', 'stores_issued',
         'teardown_begin', 'exit', 'cs_outbox', 'cs_bar', 'cs_sent', 'cs_recv', 'first_tile_known']
COLD = '--cold' in sys.argv
SHAPES = [(4096, 4096, 7168), (4096, 7168, 2048)] if '--big' in sys.argv else [(64, 4096, 7168), (128, 4096, 7168)]
for arg in sys.argv[1:]:
    if arg.startswith('--shapes='):        # --shapes=64x7168x2048,128x24576x1536
        SHAPES = [tuple(int(v) for v in s.split('x')) for s in arg.split('=', 1)[1].split(',')]
flush = torch.empty(256 << 20, dtype=torch.int32, device='cuda') if COLD else None
ts = torch.zeros(16 + 2 * 160, dtype=torch.int64, device='cuda')
_lib.lib().dgb200_debug_set_timestamps(ts.data_ptr())
for (m, n, k) in SHAPES:
    for splits in ('1',):
        a, b, qa, qb = make_inputs(m, n, k)
        sfa = dg.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
        sfb = dg.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
        d = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
        for _ in range(3):
            ts.zero_()
            if COLD:
                flush.zero_()
            dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d)
            torch.cuda.synchronize()
        v = ts.cpu().tolist()
        rel = {nm: v[i] - v[0] for i, nm in enumerate(NAMES)}
        per = torch.tensor(v[16:16 + 2 * 148]).view(148, 2)
        t0 = int(per[:, 0][per[:, 0] > 0].min())
        starts = sorted(int(x) - t0 for x in per[:, 0].tolist() if x > 0)
        ends = sorted(int(x) - t0 for x in per[:, 1].tolist() if x > 0)
        q = lambda arr, f: arr[min(len(arr) - 1, int(f * len(arr)))]
        print(json.dumps({'m': m, 'n': n, 'k': k, 'cfg': _lib.last_config(), 'cycles_from_entry': rel,
                          'cta_start_ns': [starts[0], q(starts, 0.5), starts[-1]],
                          'cta_end_ns': [ends[0], q(ends, 0.25), q(ends, 0.5), q(ends, 0.75), ends[-1]]}), flush=True)
_lib.lib().dgb200_debug_set_timestamps(None)
