"""Summarises an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of bench.py into
(a) per-kernel time shares of the captured window and (b) mean DRAM bytes per launch of the three-term convolution kernels (bench.py `roofline.traffic`).
python tools/ncu_step_summary.py <launches.csv> <out.txt> [traffic.json batch_key]"""
import collections
import csv
import json
import re
import sys

src, out = sys.argv[1], sys.argv[2]
rows = list(csv.reader(l for l in open(src, errors='replace') if l.startswith('"')))
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
per = collections.defaultdict(dict)
names = {}
for r in rows[1:]:
    per[r[ix['ID']]][r[ix['Metric Name']]] = float(r[ix['Metric Value']].replace(',', ''))
    names[r[ix['ID']]] = r[ix['Kernel Name']]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for i, m in per.items():
    k = re.sub(r'\(.*', '', names[i]).replace('void ', '').replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
    a = agg[k]; a[0] += 1; a[1] += m.get('gpu__time_duration.sum', 0.0); a[2] += m.get('dram__bytes_read.sum', 0.0) + m.get('dram__bytes_write.sum', 0.0)
tot = sum(a[1] for a in agg.values())
unit = 'ns'
lines = [f'# {src}: {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.2f} ms of kernel time (serialised, cold-cache replays: SHARES are meaningful, not absolutes)',
         f'# {"kernel":88s} {"launches":>8s} {"time ms":>10s} {"share":>7s} {"DRAM MB/launch":>15s}']
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    lines.append(f'{k[:90]:90s} {a[0]:8d} {a[1] / 1e6:10.3f} {a[1] / tot * 100:6.2f}% {a[2] / a[0] / 1e6:15.1f}')
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:20]))
if len(sys.argv) > 4:
    three = [(k, a) for k, a in agg.items() if re.search(r'conv_nhwc_bf16_kernel<\d+, 3', k)]
    n = sum(a[0] for _, a in three); b = sum(a[2] for _, a in three)
    try:
        j = json.load(open(sys.argv[3]))
    except Exception:
        j = {}
    j[sys.argv[4]] = b / max(n, 1)
    json.dump(j, open(sys.argv[3], 'w'))
    print('three-term conv: launches', n, 'mean DRAM bytes / launch', b / max(n, 1))
