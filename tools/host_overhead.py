"""Host time per GEMM call (the GPU work is a tiny launch, so the loop is host-bound): the Python wrapper, the bare C-ABI call
through ctypes, and the reference's pybind path on the same tensors. Development tool."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools.bringup import import_reference, make_inputs  # noqa: E402
import deepgemm_b200 as dg  # noqa: E402
from deepgemm_b200 import _lib, gemm  # noqa: E402

m, n, k = 16, 128, 512
a, b, qa, qb = make_inputs(m, n, k)
sfa = dg.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
sfb = dg.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
d = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
N = 20000


def loop(fn):
    for _ in range(200):
        fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(N):
        fn()
    dt = time.perf_counter() - t
    torch.cuda.synchronize()
    return round(dt / N * 1e6, 2)


out = {'shape': [m, n, k], 'calls': N}
out['ours_python_us'] = loop(lambda: dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d))
lib = _lib.lib()
stream = torch.cuda.current_stream().cuda_stream
args = (qa[0].data_ptr(), sfa.data_ptr(), qb[0].data_ptr(), sfb.data_ptr(), d.data_ptr(), m, n, k, k, k, n, 0, 0,
        sfa.stride(-1), sfb.stride(-1), 128, 128, 0, 0, 0, 0, stream)
out['ours_ctypes_call_us'] = loop(lambda: lib.dgb200_fp8_gemm_nt(*args))
try:
    ref = import_reference()
    sfa_r = ref.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
    sfb_r = ref.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
    out['reference_pybind_us'] = loop(lambda: ref.fp8_gemm_nt((qa[0], sfa_r), (qb[0], sfb_r), d))
except Exception as e:  # noqa: BLE001
    out['reference_pybind_us'] = str(e)[:100]
print(json.dumps(out))
