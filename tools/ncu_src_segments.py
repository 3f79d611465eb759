"""Stall samples of one kernel from `ncu --page source --csv`, summed between its synchronisation instructions
(barrier waits, TMEM loads / stores, arrivals) — where a warp-specialised kernel's time goes, role by role.
usage: ncu_src_segments.py file.csv [min_samples]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 60
s0 = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'][0]
hdr = rows[s0 + 1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[s0 + 2:] if len(r) == len(hdr)]
samp = [int(r[ix['# Samples']]) for r in data]
src = [r[ix['Source']].strip() for r in data]
tot = sum(samp)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
marks = ('USETMAXREG', 'LDTM', 'SYNCS.ARRIVE', 'NANOSLEEP', 'UTMASTG', 'STTM.x16', 'EXIT', 'BAR.SYNC')
keys = [i for i, sx in enumerate(src) if any(k in sx for k in marks)]
print('kernel:', rows[s0][1][:80], ' total samples', tot)
prev = 0
for i in keys:
    seg = sum(samp[prev:i])
    if seg > thr or samp[i] > thr:
        st = {h[6:]: sum(int(r[ix[h]]) for r in data[prev:i]) for h in stall_cols}
        st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f'[{prev:5d},{i:5d}) {seg:6d} {100 * seg / tot:5.1f}%  then {src[i][:48]:48s} {samp[i]:5d}  {st}')
    prev = i + 1
