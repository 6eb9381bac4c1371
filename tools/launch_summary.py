"""Per-kernel totals and shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_summary.py launches.csv out.md "<command line that was profiled>" """
import csv
import sys
from collections import OrderedDict

src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
name_i, val_i, unit_i = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = OrderedDict()
order = []
for r in rows[1:]:
    v = float(r[val_i].replace(',', ''))
    v = v / 1e3 if r[unit_i] in ('ns', 'nsecond') else (v * 1e3 if r[unit_i] in ('ms', 'msecond') else v)
    a = agg.setdefault(r[name_i], [0, 0.0])
    a[0] += 1
    a[1] += v
    order.append((r[name_i], v))
total = sum(a[1] for a in agg.values())
lines = [f'# ncu launch list of `{cmd}` (gpu__time_duration.sum, --clock-control none)', '',
         'Per-launch times are cold-cache and serialised under ncu; what must agree with bench.py is each kernel\'s SHARE.', '',
         '| kernel | launches | total us | share |', '|---|---|---|---|']
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    lines.append(f'| `{name[:96]}` | {n} | {t:.1f} | {100 * t / total:.1f}% |')
ours = {k: v for k, v in agg.items() if 'dgb200::' in k}
gemm = {k: v for k, v in ours.items() if 'fp8_gemm_kernel' in k}
gt = sum(v[1] for v in gemm.values())
lines += ['', f'Kernels of this library: {sum(v[0] for v in ours.values())} launches, {sum(v[1] for v in ours.values()):.1f} us; the GEMM kernels among them '
              f'{sum(v[0] for v in gemm.values())} launches, {gt:.1f} us = {100 * gt / total:.1f}% of all GPU time in the command (the rest is the 512 MB L2 flush '
              'memsets, input generation and the end-to-end copies, all outside the timed CUDA-event brackets).', '',
          'Share of each GEMM instantiation inside the GEMM time (compare with `roofline.share_of_step` in the bench line):', '']
for name, (n, t) in sorted(gemm.items(), key=lambda kv: -kv[1][1]):
    lines.append(f'* `{name[:110]}`: {n} launches, {t:.1f} us, {100 * t / gt:.1f}%')
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[-8:]))
