"""Summarise an ncu launch list of bench.py into (a) a per-kernel table (time share, DRAM bytes per launch) and (b)
profiles/kernel_traffic.json, the file bench.py reads `roofline.traffic` from.

Capture (one GPU, on the box):
  TCOW_CUDA_GRAPH=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
      -c 1200 --csv --log-file gpurun_out/rNN_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
      --no-eager-baseline --no-train
usage: python tools/ncu_traffic.py gpurun_out/rNN_launches.csv profiles/rNN_launches_summary.txt [profiles/kernel_traffic.json]
Per-launch times under ncu are cold-cache and serialised: compare SHARES with the bench.py breakdown, not absolutes."""
import collections
import csv
import json
import re
import sys

src, out_txt = sys.argv[1], sys.argv[2]
out_json = sys.argv[3] if len(sys.argv) > 3 else None
rows = [r for r in csv.reader(open(src, errors='replace')) if len(r) >= 15 and r[0].isdigit()]
per = collections.OrderedDict()          # launch id -> {name, metric: value}
for r in rows:
    d = per.setdefault(r[0], {'name': r[4]})
    try:
        d[r[12]] = float(r[14].replace(',', ''))
    except ValueError:
        pass
units = {r[12]: r[13] for r in rows}


def to_bytes(v, unit):
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def to_ns(v, unit):
    return v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(unit, 1)


agg = collections.defaultdict(lambda: [0.0, 0, 0.0, 0.0])
for d in per.values():
    name = re.sub(r'\((?:int|bool|unsigned int)\)', '', d['name'])
    key = re.sub(r'\(.*', '', name).replace('void ', '').replace('tcow::', '')
    a = agg[key]
    a[0] += to_ns(d.get('gpu__time_duration.sum', 0.0), units.get('gpu__time_duration.sum', 'ns'))
    a[1] += 1
    a[2] += to_bytes(d.get('dram__bytes_read.sum', 0.0), units.get('dram__bytes_read.sum', 'byte'))
    a[3] += to_bytes(d.get('dram__bytes_write.sum', 0.0), units.get('dram__bytes_write.sum', 'byte'))
tot = sum(a[0] for a in agg.values())
lines = [f'# {src}: per-kernel totals of the ncu launch list (see tools/ncu_traffic.py for the capture command)',
         '# per-launch times are cold-cache and serialised: compare SHARES with the bench.py breakdown, not absolutes',
         f'{"time ms":>10} {"share":>6} {"n":>5} {"avg us":>9} {"rd MB/launch":>13} {"wr MB/launch":>13}  kernel']
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    lines.append(f'{a[0] / 1e6:10.3f} {100 * a[0] / tot:5.1f}% {a[1]:5d} {a[0] / a[1] / 1e3:9.1f} {a[2] / a[1] / 1e6:13.1f} '
                 f'{a[3] / a[1] / 1e6:13.1f}  {k[:90]}')
open(out_txt, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:24]))
if out_json:
    classes = {}
    fam = collections.defaultdict(lambda: [0, 0.0])
    for k, a in agg.items():
        if a[2] + a[3] > 0:
            classes[k] = {'launches': a[1], 'dram_bytes_per_launch': (a[2] + a[3]) / a[1], 'avg_us_under_ncu': a[0] / a[1] / 1e3}
            base = re.sub(r'<.*', '', k)
            fam[base][0] += a[1]
            fam[base][1] += a[2] + a[3]
    for base, (n, b) in fam.items():     # launch-weighted mean over all template instances of a kernel family
        classes.setdefault(base, {'launches': n, 'dram_bytes_per_launch': b / n})
    json.dump({'source': f'{src} via tools/ncu_traffic.py (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, B=8 bench)',
               'classes': classes}, open(out_json, 'w'), indent=1)
