"""Print the hottest SASS lines (by warp-stall samples) of an ncu source-page CSV, with dominant stall reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        s = int(r[idx['# Samples']])
    except ValueError:
        continue
    data.append((s, r))
total = sum(s for s, _ in data)
print('total samples', total)
top = sorted(range(len(data)), key=lambda i: -data[i][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(top):
    s, r = data[i]
    reasons = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f'{i:5d} {s:7d} {100.0 * s / total:5.1f}%  {r[idx["Source"]].strip():70s} ex={r[idx["Instructions Executed"]]:>8s} ' +
          ' '.join(f'{n}:{c}' for c, n in reasons if c))
