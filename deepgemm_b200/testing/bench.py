"""Device-side timing helpers.

`bench_events` times a callable with CUDA events on the current stream after warm-up, flushing L2 between
iterations by writing a buffer larger than the 126 MB L2 (the reference's bench_kineto zeroes 8 GB for the same
purpose, deep_gemm/testing/bench.py:92-108; we flush with 512 MB which is > 4x L2).
"""
from typing import Callable, List

import torch

_flush_buf = None


def flush_l2(num_bytes: int = 512 << 20) -> None:
    global _flush_buf
    if _flush_buf is None or _flush_buf.numel() < num_bytes // 4:
        _flush_buf = torch.empty(num_bytes // 4, dtype=torch.int32, device='cuda')
    _flush_buf.zero_()


def bench_events(fn: Callable[[], None], num_warmups: int = 3, num_tests: int = 10, flush: bool = True) -> List[float]:
    """Per-iteration device time in seconds (list of `num_tests` samples)."""
    for _ in range(num_warmups):
        fn()
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(num_tests)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(num_tests)]
    for i in range(num_tests):
        if flush:
            flush_l2()
        starts[i].record()
        fn()
        ends[i].record()
    torch.cuda.synchronize()
    return [s.elapsed_time(e) * 1e-3 for s, e in zip(starts, ends)]


def bench_kineto(fn: Callable[[], None], kernel_names, num_tests: int = 30, suppress_kineto_output: bool = True,
                 flush_l2_size: int = 512 << 20, **_):
    """Mean device time of the kernels whose name contains `kernel_names`, measured with torch.profiler the way the
    reference does (deep_gemm/testing/bench.py:79-146) so numbers are directly comparable."""
    is_tuple = isinstance(kernel_names, (tuple, list))
    names = tuple(kernel_names) if is_tuple else (kernel_names,)
    fn()
    schedule = torch.profiler.schedule(wait=0, warmup=1, active=1, repeat=1)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA], schedule=schedule) as prof:
        for _ in range(2):
            for _ in range(num_tests):
                flush_l2(flush_l2_size)
                fn()
            torch.cuda.synchronize()
            prof.step()
    totals = {n: [0.0, 0] for n in names}
    for evt in prof.key_averages():
        for n in names:
            if n in evt.key:
                totals[n][0] += evt.device_time_total * 1e-6
                totals[n][1] += evt.count
    out = tuple((t / c if c else 0.0) for t, c in totals.values())
    return out if is_tuple else out[0]
