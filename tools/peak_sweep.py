"""Issue-only FP8 tcgen05.mma rate versus UMMA N (development tool; the headline probe lives in bench.py `fp8_peak`)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deepgemm_b200 import _lib  # noqa: E402

lib = _lib.lib()
sms = torch.cuda.get_device_properties(0).multi_processor_count & ~1
stream = torch.cuda.current_stream().cuda_stream
for n in (16, 32, 64, 96, 128, 144, 160, 176, 192, 208, 224, 240, 256):
    iters = 4096
    flops = (sms // 2) * iters * 4 * 2.0 * 256 * n * 32
    for _ in range(3):
        _lib.check(lib.dgb200_debug_fp8_peak(n, iters, sms, stream))
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.dgb200_debug_fp8_peak(n, iters, sms, stream))
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    cyc = best * 1e-3 * 1.965e9 / (iters * 4)
    print(json.dumps({'umma_n': n, 'burst_tflops': round(flops / (best * 1e-3) / 1e12, 1), 'ms': round(best, 3),
                      'cycles_per_umma_at_1965MHz': round(cyc, 1)}), flush=True)
