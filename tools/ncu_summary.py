"""Summarise .ncu-rep files (or their `--page raw --csv` dumps) (key metrics per kernel launch) as markdown. usage: ncu_summary.py out.md rep1 [rep2 ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    ('gpu__time_duration.sum', 'duration'),
    ('sm__cycles_elapsed.avg.per_second', 'SM clock'),
    ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('dram__bytes_read.sum.per_second', 'DRAM read rate'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput % of peak'),
    ('l1tex__m_xbar2l1tex_read_bytes.sum', 'L2->SM bytes'),
    ('l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'L2->SM rate'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit rate'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput % of peak'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem'),
    ('launch__cluster_size', 'cluster size'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy'),
]


KEYS = {'ours_4096': 'dense_m4096', 'ours_512': 'dense_m512', 'ours_64': 'dense_m64', 'ours_128': 'dense_m128', 'ours_192': 'dense_m192',
        'ours_k2048': 'dense_4096x7168x2048', 'ref_k2048': 'reference_dense_4096x7168x2048',
        'ours_contig': 'contiguous_g256_m128', 'ref_contig': 'reference_contiguous_g256_m128', 'ours_masked': 'masked_g256_m64',
        'ours_quant': 'per_token_cast_to_fp8_4096x7168', 'ep_dispatch': 'ep_dispatch', 'ep_combine': 'ep_dispatch',
        'ref_4096': 'reference_dense_m4096', 'ref_64': 'reference_dense_m64'}
TRAFFIC = {}
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def to_bytes(value, unit):
    return float(value.replace(',', '')) * UNIT.get(unit, 1)


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    lines = ['# ncu summaries (`ncu --set full --clock-control none --import-source on`, B200)\n']
    for rep in reps:
        if rep.endswith('.csv'):
            raw = open(rep).read()
        else:
            raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            lines.append(f'## {rep}\n(no data)\n')
            continue
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
            name = d.get('Kernel Name', ('?', ''))[0]
            lines.append(f'## {rep.split("/")[-1]} — `{name[:110]}`\n')
            lines.append('| metric | value |\n|---|---|')
            for key, label in WANT:
                if key in d:
                    v, u = d[key]
                    lines.append(f'| {label} (`{key}`) | {v} {u} |')
            rd, wr = d.get('dram__bytes_read.sum'), d.get('dram__bytes_write.sum')
            key = KEYS.get(rep.split('/')[-1].split('.')[0])
            if key == 'ep_dispatch':
                key = 'ep_' + name.split('(')[0].split('<')[0].split()[-1]
            if key and rd and wr:
                TRAFFIC[key] = {'dram_read_bytes': to_bytes(*rd), 'dram_write_bytes': to_bytes(*wr),
                                'duration_us': float(d['gpu__time_duration.sum'][0]), 'kernel': name[:80]}
            lines.append('')
    open(out_path, 'w').write('\n'.join(lines) + '\n')
    print('wrote', out_path)
    if TRAFFIC:
        import json
        import os
        path = os.path.join(os.path.dirname(out_path), 'traffic.json')
        json.dump(TRAFFIC, open(path, 'w'), indent=1, sort_keys=True)
        print('wrote', path)


if __name__ == '__main__':
    main()
