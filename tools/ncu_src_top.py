"""Print the instructions with the most stall samples from `ncu --page source --csv` output.
usage: ncu_src_top.py file.csv [section_index] [n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
print('sections:', [(k, rows[s][1][:60]) for k, s in enumerate(starts)])
s0 = starts[sec]
s1 = starts[sec + 1] if sec + 1 < len(starts) else len(rows)
hdr = rows[s0 + 1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[s0 + 2:s1] if len(r) == len(hdr)]
tot = sum(int(r[ix['# Samples']]) for r in data)
print('kernel:', rows[s0][1][:100], ' total samples', tot, ' instructions', len(data))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix['# Samples']]))[:n]
for i in order:
    r = data[i]
    st = {h[6:]: int(r[ix[h]]) for h in stall_cols if int(r[ix[h]]) > 0}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{i:5d} {int(r[ix['# Samples']]):6d} {100 * int(r[ix['# Samples']]) / tot:5.1f}%  {r[ix['Source']].strip()[:64]:64s} {st}")
