#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export: contiguous SASS regions with
similar executed counts, their share of all executed instructions and of the stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
hdr = rows[1]; data = []
for r in rows[2:]:
    if len(r) < 10: continue
    if r[0] == 'Address' or r[0].startswith('Kernel'): break
    data.append(r)
iS = hdr.index('Source'); iE = hdr.index('Instructions Executed'); iN = hdr.index('# Samples')
tot = sum(int(r[iE]) for r in data); ts = sum(int(r[iN]) for r in data)
print('total warp instr', tot, 'sass lines', len(data), 'samples', ts)
segs = []; cur = None
for idx, r in enumerate(data):
    e = int(r[iE]); s = int(r[iN])
    if cur and abs(e - cur['e']) <= 0.03 * max(e, cur['e'], 1):
        cur['n'] += 1; cur['sum'] += e; cur['samp'] += s; cur['end'] = idx
    else:
        cur = {'e': e, 'n': 1, 'sum': e, 'samp': s, 'start': idx, 'end': idx}; segs.append(cur)
for s in segs:
    if s['sum'] > tot * thr or s['samp'] > ts * thr:
        print(f"{s['start']:5d}-{s['end']:5d} n={s['n']:4d} exec/instr={s['e']:>10d} instr={s['sum']/tot*100:5.1f}% samples={s['samp']/ts*100:5.1f}%  {data[s['start']][iS].strip()[:60]}")
