"""One line per captured launch of an `ncu --set full` report: duration, DRAM bytes, pipe activity, registers, grid.
usage: python tools/ncu_full_summary.py gpurun_out/rNN_full.ncu-rep > profiles/rNN_ncu_full.txt   (ncu on PATH; no GPU needed)"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale=None):
    i = ix.get(name)
    if i is None or r[i] in ('', 'n/a'):
        return float('nan')
    v = float(r[i].replace(',', ''))
    u = units[i]
    if scale == 'MB':
        v *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
    if scale == 'us':
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1.0)
    return v


print('kernel'.ljust(52) + '      us dram_rd_MB dram_wr_MB traffic_MB tensor%   xu% issue%  regs   grid')
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[ix['Kernel Name']].replace('tcow::', '').replace('void ', '')[:50]
    rd, wr = val(r, 'dram__bytes_read.sum', 'MB'), val(r, 'dram__bytes_write.sum', 'MB')
    print(name.ljust(52) + f"{val(r, 'gpu__time_duration.sum', 'us'):8.1f} {rd:10.1f} {wr:10.1f} {rd + wr:10.1f} "
          f"{val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):7.1f} "
          f"{val(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):5.1f} "
          f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{val(r, 'launch__registers_per_thread'):5.0f} {r[ix['Grid Size']] if 'Grid Size' in ix else ''}")
