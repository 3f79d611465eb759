"""Per-kernel counts of the Blackwell-specific SASS instructions in libtcow_b200.so (cuobjdump -sass; no GPU needed).
usage: python tools/sass_inventory.py > profiles/sass_inventory.txt
UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG/UTMAREDG = TMA load/store/reduce-add,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = mma.sync, MUFU = special-function unit, FFMA2/FADD2/FMUL2 = packed fp32x2."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'tcow_b200', 'libtcow_b200.so')
OPS = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTCBAR', 'SYNCS', 'HMMA', 'MUFU', 'FFMA2', 'FADD2', 'FMUL2',
       'FMNMX3', 'LDL', 'STL']
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
counts, total, name = {}, {}, None
for line in sass.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        total[name] = 0
        continue
    if name is None:
        continue
    m = re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
    if m:
        total[name] += 1
        op = m.group(1)
        if op in OPS:
            counts[name][op] += 1
print(f'# {os.path.relpath(lib, ROOT)}: SASS instruction inventory per kernel (cuobjdump -sass, sm_100a); tools/sass_inventory.py')
print('kernel'.ljust(72) + ' insts ' + ' '.join(o.rjust(8) for o in OPS))
tot = collections.Counter()
for n in sorted(counts, key=lambda k: -total[k]):
    d = re.sub(r'\(.*', '', re.sub(r'\((?:int|bool|unsigned int)\)', '', demangle(n)).replace('tcow::', '').replace('void ', ''))[:70]
    print(d.ljust(72) + f'{total[n]:6d} ' + ' '.join((str(counts[n][o]) if counts[n][o] else '.').rjust(8) for o in OPS))
    tot.update(counts[n])
print('TOTAL'.ljust(72) + f'{sum(total.values()):6d} ' + ' '.join(str(tot[o]).rjust(8) for o in OPS))
