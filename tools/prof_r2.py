"""One workload a few times, for ncu (round 2). usage:
  prof_r2.py dense {ours|ref} M N K | contiguous {ours|ref} mean_m | masked {ours|ref} mean_m | quant M K | ep"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden'))
import torch  # noqa: E402

kind = sys.argv[1]
flush = torch.empty(256 << 20, dtype=torch.int32, device='cuda')
iters = 3


def lib_of(which):
    if which == 'ref':
        from tools.bringup import import_reference
        return import_reference()
    import deepgemm_b200
    return deepgemm_b200


if kind == 'dense':
    from tools.bringup import make_inputs
    lib = lib_of(sys.argv[2])
    m, n, k = (int(x) for x in sys.argv[3:6])
    a, b, qa, qb = make_inputs(m, n, k)
    d = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    sfa = lib.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
    sfb = lib.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
    for _ in range(iters):
        flush.zero_()
        lib.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d)
elif kind in ('contiguous', 'masked'):
    import cases
    from deepgemm_b200 import utils
    lib = lib_of(sys.argv[2])
    mean_m = int(sys.argv[3])
    if kind == 'contiguous':
        p = cases.make_contiguous(mean_m, utils)
        sfa = lib.transform_sf_into_required_layout(p['a'][1], p['m'], 7168, (1, 128, 128), None, True)
        sfb = lib.transform_sf_into_required_layout(p['b'][1], 4096, 7168, (1, 128, 128), 256, False)
        d = torch.empty((p['m'], 4096), device='cuda', dtype=torch.bfloat16)
        for _ in range(iters):
            flush.zero_()
            lib.m_grouped_fp8_gemm_nt_contiguous((p['a'][0], sfa), (p['b'][0], sfb), d, p['layout'])
    else:
        p = cases.make_masked(mean_m, utils)
        sfa = lib.transform_sf_into_required_layout(p['a'][1], 128, 2048, (1, 128, 128), 256, True)
        sfb = lib.transform_sf_into_required_layout(p['b'][1], 7168, 2048, (1, 128, 128), 256, False)
        d = torch.zeros((256, 128, 7168), device='cuda', dtype=torch.bfloat16)
        for _ in range(iters):
            flush.zero_()
            lib.m_grouped_fp8_gemm_nt_masked((p['a'][0], sfa), (p['b'][0], sfb), d, p['masked_m'], p['expected_m'])
elif kind == 'quant':
    import deepgemm_b200 as dg
    m, k = int(sys.argv[2]), int(sys.argv[3])
    x = torch.randn((m, k), device='cuda', dtype=torch.bfloat16)
    for _ in range(iters):
        flush.zero_()
        dg.per_token_cast_to_fp8_packed(x)
elif kind == 'ep':
    import deepgemm_b200 as dg
    from deepgemm_b200 import ep
    import bench
    b, sfb_p, xq, sf_packed, ids, capacity = bench.ep_problem(0, 1, torch.device('cuda', 0), dg, 256, 4096, 7168, 32768)
    buf = ep.EpBuffer(256, capacity, 7168)
    d = buf.output(4096)
    row = torch.empty(xq.shape[0], dtype=torch.int32, device='cuda')
    out = torch.empty((xq.shape[0], 4096), device='cuda', dtype=torch.bfloat16)
    for _ in range(iters):
        flush.zero_()
        r = buf.dispatch(xq, sf_packed, ids, row)
        buf.grouped_gemm((b, sfb_p), d, r.expected_m, overlap=False)
        buf.combine(row, ids, out)
    torch.cuda.synchronize()
    buf.close()
torch.cuda.synchronize()
print('done', sys.argv[1:])
