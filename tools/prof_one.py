"""Run one dense GEMM a few times (for ncu). usage: prof_one.py {ours|ref} M N K [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools.bringup import import_reference, make_inputs  # noqa: E402

which, m, n, k = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
a, b, qa, qb = make_inputs(m, n, k)
d = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
if which == 'ref':
    lib = import_reference()
else:
    import deepgemm_b200 as lib
sfa = lib.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
sfb = lib.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
flush = torch.empty(256 << 20, dtype=torch.int32, device='cuda')
for _ in range(iters):
    flush.zero_()
    lib.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d)
torch.cuda.synchronize()
print('done', which, m, n, k)
