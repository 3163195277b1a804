#!/usr/bin/env python
"""Stall-sample breakdown of one kernel from an `ncu --page source --csv` dump (first matching launch):
samples by opcode, totals by stall reason, and the hottest SASS lines.
usage: python tools/ncu_src_stalls.py <src.csv> <kernel substr> [top N]"""
import collections, csv, re, sys

path, sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
fn, hdr, rows = None, None, []
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] in ('Function Name', 'Kernel Name'):
        if fn and sub in fn and rows:
            break
        fn = row[1]
        continue
    if row[0] in ('Line No', 'Address'):
        hdr = {}
        for i, h in enumerate(row):
            hdr.setdefault(h, i)
        src_cols = [i for i, h in enumerate(row) if h == 'Source']
        continue
    if fn and sub in fn and hdr and len(row) > hdr['Address'] and row[hdr['Address']].strip():
        rows.append(row)


def num(r, h):
    try:
        return int(r[hdr[h]].replace(',', '') or 0)
    except (ValueError, IndexError):
        return 0


tot = sum(num(r, '# Samples') for r in rows)
print(fn, '| sass rows', len(rows), '| samples', tot)
c, ci = collections.Counter(), collections.Counter()
for r in rows:
    s = re.sub(r'^\s*@!?U?P\w+\s+', '', r[src_cols[-1]].strip())
    op = s.split()[0].split('.')[0] if s else '?'
    c[op] += num(r, '# Samples')
    ci[op] += num(r, 'Instructions Executed')
for op, n in c.most_common(14):
    print(f'  {op:10s} samples {n:7d} {100 * n / max(tot, 1):5.1f}%   inst {ci[op]}')
st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tt = {h: sum(num(r, h) for r in rows) for h in st}
print('  by reason:', ', '.join(f'{h[6:]} {v}' for h, v in sorted(tt.items(), key=lambda kv: -kv[1])[:10]))
rows.sort(key=lambda r: -num(r, '# Samples'))
for r in rows[:top]:
    why = {h[6:]: num(r, h) for h in st if num(r, h) > max(20, num(r, '# Samples') // 8)}
    print(f"  {r[hdr['Address']][-5:]} {r[src_cols[-1]][:64]:64s} {num(r, '# Samples'):6d} {why}")
