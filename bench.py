#!/usr/bin/env python
"""bench.py -- one JSON line per run (see the round contract).

Default workload (`--workload dense`, BASELINE.json configs[1]): one STEP = one pass of `fp8_gemm_nt` over the four
DeepSeek-V3 dense shapes M in {64,128,512,4096}, N=4096, K=7168 (synthetic BF16 randn, quantised like the
reference's tests: 1x128 UE8M0 token scales, 128x128 UE8M0 weight scales). `value` = sum(2MNK) / sum(device time) in
TFLOPS with operands resident in HBM and packed scale factors prepared (kernel-only, cold L2: a 512 MB flush
precedes every timed launch, timed with CUDA events on the launching stream). `e2e` = the same step through the
public API from PINNED HOST buffers: H2D of A/B/scale factors, FP32->UE8M0 scale packing, the GEMMs, D2H of D.

Other workloads (not the driver's default): `--workload contiguous` (config 3), `--workload masked` (config 4,
CUDA-graph replay). With --gpus N > 1 every rank runs an independent replica (the dense GEMM does not shard:
"replicas only"); `value` is the sum over ranks / max-over-ranks time.

`--impl reference` times the CPU arm: the torch-CPU BF16-emulated blockwise GEMM (oracle port; the reference has no
CPU implementation of this path) on a bounded sample of the same step with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

DENSE_SHAPES = [(64, 4096, 7168), (128, 4096, 7168), (512, 4096, 7168), (4096, 4096, 7168)]
NOMINAL_FP8_TFLOPS = 4500.0
FALLBACK_PEAKS = {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}


def load_peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), 'measured'
    return dict(FALLBACK_PEAKS), 'fallback'


def committed_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed `ncu --set full` capture
    (profiles/traffic.json, written by tools/ncu_summary.py); None if that kernel has no capture."""
    path = os.path.join(REPO, 'profiles', 'traffic.json')
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f).get(key)
    return None if t is None else int(t['dram_read_bytes'] + t['dram_write_bytes'])


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clocks / throttle reasons while the timed region runs: NVML in-process every ~2 ms (the timed region of
    the default run lasts ~10 ms, one `nvidia-smi` fork takes longer than that); falls back to forking nvidia-smi."""
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(visible.split(',')[index]) if visible and visible.split(',')[index].isdigit() else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:  # noqa: BLE001
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:  # noqa: BLE001
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        flag = lambda bit: 'Active' if mask & bit else 'Not Active'   # noqa: E731
        try:
            watts = n.nvmlDeviceGetPowerUsage(self._h) / 1000.0
        except Exception:  # noqa: BLE001
            watts = 0.0
        return [str(sm), str(self._max), f'{watts:.0f}', flag(0x8), flag(0x40), flag(0x20), flag(0x4)]

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(',')])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.002 if self._nvml is not None else 0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        mx = max((float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()), default=None)
        watts = sorted(float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace('.', '').isdigit())
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_min_mhz': sm[0] if sm else None, 'sm_max_mhz': mx,
                'reasons': sorted(reasons), 'samples': len(self.rows), 'power_w_median': watts[len(watts) // 2] if watts else None,
                'source': 'nvml' if self._nvml is not None else 'nvidia-smi'}


# ------------------------------------------------------------------------------------------------ workloads
def make_dense_problem(m, n, k, device, seed):
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    g = torch.Generator(device=device).manual_seed(seed)
    a = torch.randn((m, k), device=device, dtype=torch.bfloat16, generator=g)
    b = torch.randn((n, k), device=device, dtype=torch.bfloat16, generator=g)
    qa, qb = per_token_cast_to_fp8(a, True), per_block_cast_to_fp8(b, True)
    return qa, qb


def run_dense(args, rank, world, device):
    import deepgemm_b200 as dg
    from deepgemm_b200 import _lib
    from deepgemm_b200.testing import flush_l2
    peaks, peak_kind = load_peaks()
    probs = []
    for i, (m, n, k) in enumerate(DENSE_SHAPES):
        qa, qb = make_dense_problem(m, n, k, device, seed=i)
        sfa = dg.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
        sfb = dg.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
        d = torch.empty((m, n), device=device, dtype=torch.bfloat16)
        probs.append(dict(m=m, n=n, k=k, qa=qa, qb=qb, sfa=sfa, sfb=sfb, d=d))

    def step_kernel_only(record=None):
        for i, p in enumerate(probs):
            flush_l2()
            if record is not None:
                record[i][0].record()
            dg.fp8_gemm_nt((p['qa'][0], p['sfa']), (p['qb'][0], p['sfb']), p['d'])
            if record is not None:
                record[i][1].record()
            elif 'tile' not in p:
                cfg = _lib.last_config()
                p['tile'] = {k_: cfg[k_] for k_ in ('block_m', 'cluster', 'num_stages', 'num_splits', 'cluster_split')}

    # ---- kernel-only (inputs resident) ------------------------------------------------------------
    # The clock sampler (a thread that forks nvidia-smi) starts before the warm-up so that its start-up noise and the
    # GPU's idle->busy clock ramp fall outside the timed region.
    with ClockSampler(torch.cuda.current_device()) as clocks:
        for _ in range(args.warmup):
            step_kernel_only()
        torch.cuda.synchronize()
        barrier(world)
        events = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in probs]
                  for _ in range(args.steps)]
        launches0 = _lib.launch_count()
        clocks.rows.clear()
        t_wall0 = time.perf_counter()
        for s in range(args.steps):
            step_kernel_only(events[s])
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
    launches = _lib.launch_count() - launches0
    barrier(world)
    per_step_ms = [[e[i][0].elapsed_time(e[i][1]) for i in range(len(probs))] for e in events]
    per_shape_ms = [sum(st[i] for st in per_step_ms) / args.steps for i in range(len(probs))]
    step_ms = sum(per_shape_ms)
    step_ms = allreduce_max(step_ms, world, device)
    flops = sum(2.0 * p['m'] * p['n'] * p['k'] for p in probs)
    value = flops * world / (step_ms * 1e-3) / 1e12

    fp8_peak = 2.0 * peaks['bf16_tflops']  # measured proxy: FP8 tensor rate = 2x the measured cuBLAS BF16 burst
    per_shape = []
    for p, ms in zip(probs, per_shape_ms):
        m, n, k = p['m'], p['n'], p['k']
        fl = 2.0 * m * n * k
        byts = m * k + n * k + m * n * 2 + (m + n) * ((k + 511) // 512) * 4
        tf, gbs = fl / (ms * 1e-3) / 1e12, byts / (ms * 1e-3) / 1e9
        bound = 'tensor' if fl / byts > fp8_peak * 1e12 / (peaks['hbm_gbs'] * 1e9) else 'hbm'
        frac = tf / fp8_peak if bound == 'tensor' else gbs / peaks['hbm_gbs']
        per_shape.append({'m': m, 'n': n, 'k': k, 'us': round(ms * 1e3, 2), 'tflops': round(tf, 1), 'gbs': round(gbs, 1),
                          'bound': bound, 'frac_of_' + peak_kind: round(frac, 4), 'tile': p['tile']})

    # dominant kernel of the step = the M=4096 launch (tensor bound)
    dom = per_shape[-1]
    roofline = {'bound': 'tensor', 'kernel': 'fp8_gemm_kernel<dense> M=4096 N=4096 K=7168',
                'achieved': dom['tflops'], 'peak': round(fp8_peak, 1), 'unit': 'TFLOP/s',
                'frac': round(dom['tflops'] / fp8_peak, 4),
                'peak_source': f'2 x {peak_kind} bf16_tflops (MEASURED_PEAKS.json); nominal dense FP8 = {NOMINAL_FP8_TFLOPS}',
                'frac_of_nominal': round(dom['tflops'] / NOMINAL_FP8_TFLOPS, 4), 'traffic': committed_traffic('dense_m4096'),
                'algorithmic_bytes': int(4096 * 7168 + 4096 * 7168 + 4096 * 4096 * 2 + 2 * 4096 * 14 * 4),
                'share_of_step': round(per_shape_ms[-1] / sum(per_shape_ms), 4)}

    # ---- end to end: pinned host buffers -> H2D -> SF pack -> GEMM -> D2H ---------------------------
    host = []
    for p in probs:
        h = {k_: v.cpu().pin_memory() for k_, v in (('a', p['qa'][0].view(torch.uint8)), ('sfa', p['qa'][1]))}
        h['d'] = torch.empty((p['m'], p['n']), dtype=torch.bfloat16).pin_memory()
        host.append(h)
    hb = probs[0]['qb'][0].view(torch.uint8).cpu().pin_memory()
    hsfb = probs[0]['qb'][1].cpu().pin_memory()
    # NOTE: all four shapes share N, K: the weight matrix travels once per step
    dev_a = [torch.empty_like(p['qa'][0].view(torch.uint8)) for p in probs]
    dev_sfa = [torch.empty_like(p['qa'][1]) for p in probs]
    dev_b, dev_sfb = torch.empty_like(probs[0]['qb'][0].view(torch.uint8)), torch.empty_like(probs[0]['qb'][1])
    h2d = hb.numel() + hsfb.numel() * 4 + sum(h['a'].numel() + h['sfa'].numel() * 4 for h in host)
    d2h = sum(h['d'].numel() * 2 for h in host)

    def step_e2e():
        dev_b.copy_(hb, non_blocking=True)
        dev_sfb.copy_(hsfb, non_blocking=True)
        for i, p in enumerate(probs):
            dev_a[i].copy_(host[i]['a'], non_blocking=True)
            dev_sfa[i].copy_(host[i]['sfa'], non_blocking=True)
            dg.fp8_gemm_nt((dev_a[i].view(torch.float8_e4m3fn), dev_sfa[i]), (dev_b.view(torch.float8_e4m3fn), dev_sfb), p['d'])
            host[i]['d'].copy_(p['d'], non_blocking=True)

    for _ in range(max(args.warmup, 3)):
        step_e2e()
    torch.cuda.synchronize()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    e2e_ms = allreduce_max(e0.elapsed_time(e1) / args.steps, world, device)
    e2e_value = flops * world / (e2e_ms * 1e-3) / 1e12

    out = {
        'metric': 'FP8 TFLOPS over the DeepSeek-V3 dense shapes (sum 2MNK / sum kernel time)', 'value': round(value, 2),
        'unit': 'TFLOPS', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(step_ms, 4),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp8_e4m3 (fp32 accumulate, bf16 out)',
        'data': 'synthetic', 'impl': 'deepgemm_b200',
        'config': {'workload': 'dense fp8_gemm_nt M in {64,128,512,4096} N=4096 K=7168, 1x128/128x128 UE8M0 SF',
                   'l2': 'flushed (512 MB write) before every timed launch', 'parallelism': f'replicas x{world}'},
        'per_shape': per_shape, 'roofline': roofline, 'clocks': clocks.summary(),
        'e2e': {'value': round(e2e_value, 3), 'unit': 'TFLOPS', 'ms_per_step': round(e2e_ms, 4),
                'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h)},
        'gpu_launches': int(launches), 'wall_ms_per_step_incl_flush': round(t_wall * 1e3 / args.steps, 3),
        'per_step_us': [[round(x * 1e3, 1) for x in st] for st in per_step_ms[:8]],
    }
    return out


# ------------------------------------------------------------------------------------------------ grouped workloads
def _grouped_weights(g, n, k, device, seed):
    """[G,N,K] FP8 weights + 128x128 scales, generated expert by expert to bound memory."""
    from deepgemm_b200.utils import per_block_cast_to_fp8
    gen = torch.Generator(device=device).manual_seed(seed)
    b = torch.empty((g, n, k), device=device, dtype=torch.float8_e4m3fn)
    sfb = torch.empty((g, (n + 127) // 128, (k + 127) // 128), device=device, dtype=torch.float32)
    for i in range(g):
        b[i], sfb[i] = per_block_cast_to_fp8(torch.randn((n, k), device=device, dtype=torch.bfloat16, generator=gen), True)
    return b, sfb


def _time_events(fn, steps, warmup, world):
    from deepgemm_b200.testing import flush_l2
    for _ in range(warmup):
        fn()
    barrier(world)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for s0, s1 in evs:
        flush_l2()
        s0.record()
        fn()
        s1.record()
    barrier(world)
    return sum(a.elapsed_time(b) for a, b in evs) / steps


def run_grouped(args, rank, world, device):
    """BASELINE configs 3 (contiguous prefill, 256 experts, N=4096, K=7168) and 4 (masked decode under a CUDA graph,
    256 experts, M_max=128, N=7168, K=2048). HBM-bound: roofline = algorithmic bytes / measured HBM copy bandwidth."""
    import random
    import deepgemm_b200 as dg
    from deepgemm_b200 import _lib
    from deepgemm_b200.utils import per_token_cast_to_fp8
    peaks, peak_kind = load_peaks()
    random.seed(0)
    masked = args.workload == 'masked'
    g = 256
    if masked:
        m_max, n, k, mean_m = 128, 7168, 2048, args.mean_m or 64
    else:
        n, k, mean_m = 4096, 7168, args.mean_m or 128
    b, sfb = _grouped_weights(g, n, k, device, seed=0)
    sfb_p = dg.transform_sf_into_required_layout(sfb, n, k, (1, 128, 128), g, False)
    if masked:
        a = torch.randn((g, m_max, k), device=device, dtype=torch.bfloat16)
        qs = [per_token_cast_to_fp8(a[i], True) for i in range(g)]
        qa = (torch.stack([q[0] for q in qs]), torch.stack([q[1] for q in qs]))
        sfa = dg.transform_sf_into_required_layout(qa[1], m_max, k, (1, 128, 128), g, True)
        counts = torch.tensor([min(m_max, int(mean_m * random.uniform(0.7, 1.3))) for _ in range(g)], device=device, dtype=torch.int32)
        d = torch.zeros((g, m_max, n), device=device, dtype=torch.bfloat16)
        valid = int(counts.sum())
        call = lambda: dg.m_grouped_fp8_gemm_nt_masked((qa[0], sfa), (b, sfb_p), d, counts, int(1.2 * mean_m))  # noqa: E731
        call()
        graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                call()
        fn = graph.replay
        rows_total = valid
        name = f'masked grouped decode (CUDA graph replay), G=256 M_max=128 N=7168 K=2048 mean_m={mean_m}'
    else:
        alignment = dg.get_mk_alignment_for_contiguous_layout()
        ms = [int(mean_m * random.uniform(0.7, 1.3)) for _ in range(g)]
        aligned = [(x + alignment - 1) // alignment * alignment for x in ms]
        m = sum(aligned)
        a = torch.randn((m, k), device=device, dtype=torch.bfloat16)
        layout = torch.empty(m, device=device, dtype=torch.int32)
        s0 = 0
        for i, (mi, ai) in enumerate(zip(ms, aligned)):
            layout[s0:s0 + mi] = i
            layout[s0 + mi:s0 + ai] = -1
            a[s0 + mi:s0 + ai] = 0
            s0 += ai
        qa = per_token_cast_to_fp8(a, True)
        sfa = dg.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
        d = torch.empty((m, n), device=device, dtype=torch.bfloat16)
        valid, rows_total = sum(ms), m
        fn = lambda: dg.m_grouped_fp8_gemm_nt_contiguous((qa[0], sfa), (b, sfb_p), d, layout)  # noqa: E731
        name = f'm_grouped contiguous prefill, G=256 N=4096 K=7168 mean_m={mean_m} (sum M={m}, valid {valid}, alignment {alignment})'
    launches0 = _lib.launch_count()
    with ClockSampler(torch.cuda.current_device()) as clocks:
        ms_step = _time_events(fn, args.steps, args.warmup, world)
    ms_step = allreduce_max(ms_step, world, device)
    launches = (_lib.launch_count() - launches0) if not masked else args.steps + args.warmup  # graph replays relaunch the kernel
    flops = 2.0 * valid * n * k
    byts = rows_total * k + g * n * k + rows_total * n * 2 + (rows_total + g * n) * ((k + 511) // 512) * 4
    gbs = byts / (ms_step * 1e-3) / 1e9
    return {
        'metric': 'grouped FP8 GEMM tokens/s (valid rows / kernel time)', 'value': round(valid * world / (ms_step * 1e-3), 1),
        'unit': 'tokens/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(ms_step, 4),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp8_e4m3 (fp32 accumulate, bf16 out)',
        'data': 'synthetic', 'impl': 'deepgemm_b200', 'tflops': round(flops / (ms_step * 1e-3) / 1e12, 1),
        'config': {'workload': name, 'l2': 'flushed (512 MB write) before every timed launch', 'parallelism': f'replicas x{world}',
                   'tile': _lib.last_config()},
        'roofline': {'bound': 'hbm', 'achieved': round(gbs, 1), 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                     'frac': round(gbs / peaks['hbm_gbs'], 4), 'peak_source': peak_kind + ' hbm_gbs (MEASURED_PEAKS.json)', 'traffic': None},
        'clocks': clocks.summary(), 'gpu_launches': int(launches),
        'e2e': {'value': None, 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0,
                'note': 'grouped workloads keep the 7.5 / 3.8 GB of expert weights resident; see the dense workload for e2e'},
    }


def run_ep(args, rank, world, device):
    """BASELINE config 5: 256 experts sharded over the ranks; tokens written straight into the owners' GEMM input
    buffers by the peer-memory dispatch kernels (NVLink stores), then the local grouped GEMM. No host sync per step."""
    import deepgemm_b200 as dg
    from deepgemm_b200 import _lib, ep
    from deepgemm_b200.utils import per_token_cast_to_fp8
    g, n, k, tokens_total = 256, 4096, 7168, 32768
    epr = g // world
    t_local = tokens_total // world
    align = dg.get_mk_alignment_for_contiguous_layout()
    b, sfb = _grouped_weights(epr, n, k, device, seed=1000 + rank)
    sfb_p = dg.transform_sf_into_required_layout(sfb, n, k, (1, 128, 128), epr, False)
    gen = torch.Generator(device=device).manual_seed(rank)
    x = torch.randn((t_local, k), device=device, dtype=torch.bfloat16, generator=gen)
    xq, sf_packed = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
    ids = torch.randint(0, g, (t_local,), device=device, generator=gen)
    # capacity: balanced routing + 25% head room + alignment padding (a production caller sizes for its worst case)
    capacity = (int(t_local * 1.25) + epr * align + 127) // 128 * 128
    buf = ep.EpBuffer(g, capacity, k)
    d = buf.output(n)                      # symmetric (peer-mapped) output, so that the combine can be timed too
    token_row = torch.empty(t_local, dtype=torch.int32, device=device)
    overlap = os.environ.get('DGB200_EP_OVERLAP', '0') != '0'   # GEMM beside the scatter (per-expert arrival counters)

    def step(record=None):
        if record:
            record[0].record()
        r = buf.dispatch(xq, sf_packed, ids, token_row, wait=not overlap)
        if record:
            if overlap:
                record[1] = record[0]      # nothing may sit between the scatter and its dependent GEMM launch
            else:
                record[1].record()
        buf.grouped_gemm((b, sfb_p), d, r.expected_m, overlap=overlap)
        if record:
            record[2].record()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    assert not buf.overflowed(), 'dispatch buffer capacity exceeded'
    barrier(world)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    launches0 = _lib.launch_count()
    with ClockSampler(torch.cuda.current_device()) as clocks:
        for e in evs:
            step(e)
        torch.cuda.synchronize()
    launches = _lib.launch_count() - launches0
    barrier(world)
    disp = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    gemm = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    total = allreduce_max(disp + gemm, world, device)
    disp, gemm = allreduce_max(disp, world, device), allreduce_max(gemm, world, device)

    # the way back (top-1 combine: every source pulls its tokens' output rows over NVLink), timed on its own
    out_tokens = torch.empty((t_local, n), device=device, dtype=torch.bfloat16)
    step()
    buf.combine(token_row, ids, out_tokens)
    torch.cuda.synchronize()
    barrier(world)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    comb = 0.0
    for _ in range(3):
        step()
        c0.record()
        buf.combine(token_row, ids, out_tokens)
        c1.record()
        torch.cuda.synchronize()
        comb += c0.elapsed_time(c1) / 3
    combine_ms = allreduce_max(comb, world, device)

    # library baseline for the same dispatch: NCCL all-to-all + torch re-layout (outside the timed region)
    group = torch.distributed.group.WORLD if world > 1 else None
    def baseline():
        return ep.dispatch_alltoall(xq, sf_packed, ids, g, align, group) if world > 1 else \
            ep.dispatch_local(xq, sf_packed, ids, g, align)
    baseline()
    torch.cuda.synchronize()
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        baseline()
    e1.record()
    torch.cuda.synchronize()
    base_ms = allreduce_max(e0.elapsed_time(e1) / 3, world, device)

    row_bytes = k + 4 * ((k + 511) // 512)
    wire = t_local * row_bytes                                    # bytes this rank's scatter kernel moves (read + write each)
    remote = wire * (world - 1) / world
    peaks, peak_kind = load_peaks()
    gbs = 2.0 * wire / (disp * 1e-3) / 1e9 if disp > 0 else 0.0
    buf_rows = buf.num_rows()
    buf.close()
    return {
        'metric': 'expert-sharded grouped FP8 GEMM tokens/s (dispatch + GEMM)', 'value': round(tokens_total / (total * 1e-3), 1),
        'unit': 'tokens/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': round(total, 4),
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'fp8_e4m3 (fp32 accumulate, bf16 out)',
        'data': 'synthetic', 'impl': 'deepgemm_b200',
        'config': {'workload': f'expert-sharded grouped GEMM: 256 experts over {world} GPU(s), 32768 tokens, N=4096 K=7168, '
                               'peer-memory dispatch (NVLink stores of FP8 rows + packed UE8M0 SFs into the owner\'s GEMM buffer)',
                   'parallelism': f'ep{world}', 'l2': 'inputs (>= 0.9 GB of expert weights per rank) exceed L2'},
        'dispatch_ms': round(disp, 4), 'gemm_ms': round(gemm, 4), 'dispatch_alltoall_baseline_ms': round(base_ms, 4),
        'combine_ms': round(combine_ms, 4), 'overlap': bool(overlap), 'overlap_note': 'with overlap the two phases share the GPU: dispatch_ms/gemm_ms are stream-event splits, only ms_per_step is meaningful',
        'tflops': round(2.0 * tokens_total * n * k / (total * 1e-3) / 1e12, 1),
        'wire_bytes_per_rank': int(wire), 'remote_bytes_per_rank': int(remote), 'rows_received_rank0': buf_rows,
        'roofline': {'kernel': 'ep::scatter_kernel (+bucket/exchange/wait)', 'bound': 'hbm', 'achieved': round(gbs, 1),
                     'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': round(gbs / peaks['hbm_gbs'], 4),
                     'peak_source': peak_kind + ' hbm_gbs (MEASURED_PEAKS.json)', 'traffic': None,
                     'note': 'algorithmic bytes = read + write of every local token row once; the remote share crosses NVLink'},
        'clocks': clocks.summary(), 'gpu_launches': int(launches),
        'e2e': {'value': round(tokens_total / (total * 1e-3), 1), 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0,
                'note': 'tokens originate on the GPUs (output of the previous layer); nothing crosses PCIe in this path'},
    }


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(shapes, threads):
    """torch-CPU BF16-emulated blockwise GEMM (oracle port) over `shapes`; returns (seconds, flops)."""
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    from oracle import blockwise
    torch.set_num_threads(threads)
    total, flops = 0.0, 0.0
    for i, (m, n, k) in enumerate(shapes):
        g = torch.Generator().manual_seed(i)
        a = torch.randn((m, k), generator=g).to(torch.bfloat16)
        b = torch.randn((n, k), generator=g).to(torch.bfloat16)
        qa, qb = per_token_cast_to_fp8(a, True), per_block_cast_to_fp8(b, True)
        best = float('inf')
        for _ in range(2):
            t0 = time.perf_counter()
            blockwise.bf16_emulated_gemm_nt(qa, qb)
            best = min(best, time.perf_counter() - t0)
        total += best
        flops += 2.0 * m * n * k
    return total, flops


def cpu_baseline_block():
    threads = os.cpu_count() or 1
    sample = DENSE_SHAPES[:3]
    sec, fl = cpu_reference_step(sample, threads)
    return {'value': round(fl / sec / 1e12, 4), 'unit': 'TFLOPS', 'cores': threads, 'kind': 'port',
            'sample': 'M in {64,128,512} of the dense step (N=4096, K=7168), dequantise-to-BF16 + torch.matmul, best of 2',
            'seconds': round(sec, 3)}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return None
    threads = os.cpu_count() or 1
    sample = DENSE_SHAPES[:3]
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(sample[:1], threads)
    times = []
    fl = 0.0
    for _ in range(args.steps):
        sec, fl = cpu_reference_step(sample, threads)
        times.append(sec)
    sec = sum(times) / len(times)
    v = round(fl / sec / 1e12, 4)
    return {
        'impl': 'reference', 'metric': 'FP8 TFLOPS over the DeepSeek-V3 dense shapes (sum 2MNK / sum kernel time)',
        'value': v, 'unit': 'TFLOPS', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': round(sec * 1e3, 3), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16 emulation of fp8_e4m3 x ue8m0 (fp32 accumulate)', 'data': 'synthetic',
        'config': {'workload': 'dense fp8_gemm_nt M in {64,128,512,4096} N=4096 K=7168, 1x128/128x128 UE8M0 SF',
                   'sample': 'bounded: M in {64,128,512} per step'},
        'cpu_baseline': {'value': v, 'unit': 'TFLOPS', 'cores': threads, 'kind': 'port',
                         'sample': 'M in {64,128,512} of the dense step, torch CPU BF16 matmul of the dequantised operands'},
        'e2e': {'value': v, 'unit': 'TFLOPS', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }


# ------------------------------------------------------------------------------------------------ distributed glue
def barrier(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def allreduce_max(x, world, device):
    if world == 1:
        return x
    t = torch.tensor([x], device=device, dtype=torch.float64)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='dense', choices=['dense', 'contiguous', 'masked', 'ep'])
    ap.add_argument('--mean-m', type=int, default=0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        out = run_reference_arm(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (the FP8 GEMM path has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.distributed.init_process_group(backend='nccl', device_id=device)
    runner = {'dense': run_dense, 'contiguous': run_grouped, 'masked': run_grouped, 'ep': run_ep}[args.workload]
    out = runner(args, rank, world, device)
    if rank == 0:
        if not args.no_cpu_baseline and args.workload == 'dense':
            out['cpu_baseline'] = cpu_baseline_block()
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
